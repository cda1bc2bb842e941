"""FEA-Net stencil modules on the GPU (reference: FEANet/model.py).

``KNet`` keeps the reference's parameters (``net1``: identity "split" conv 1->C, ``net2``: C->1 conv holding one 3x3
kernel per material pattern) so that state_dicts and user edits behave as in the reference, but ``forward`` is one
sm_100a kernel: the split conv and the one-hot mask multiply are replaced by a uint8 pattern-key lookup
(``(Ku)[i,j] = sum_d W[key(i+d)][d] * u[i+d]``, weights indexed by the SOURCE node's key, zero padding).
"""
import numpy as np
import torch
import torch.nn as nn

import mgfea
from mgfea import Field, Grid, as_field, check, lib, stream_ptr

"""Note that input of KNet and FNet model is a batch of images with dimension of (N, Cin, H, W), N is batch size, Cin is number of channels."""


def _like_input(x, fld):
    """hand the result back on the caller's device: CUDA in -> strided CUDA view, CPU in -> contiguous CPU tensor"""
    if x.is_cuda:
        return fld.view
    return mgfea.to_contiguous(fld).cpu()


class KNet(nn.Module):
    def __init__(self, mesh):
        super(KNet, self).__init__()
        self.nnode_edge = mesh.nnode_edge
        self.kernel_dict = mesh.kernel_dict
        self.n_channel = len(mesh.kernel_dict)
        self._mesh = mesh
        self.net1 = nn.Conv2d(in_channels=1, out_channels=self.n_channel, kernel_size=3, padding=1, bias=False)
        self.net2 = nn.Conv2d(in_channels=self.n_channel, out_channels=1, kernel_size=3, padding=1, bias=False)
        with torch.no_grad():
            ident = torch.zeros(3, 3)
            ident[1, 1] = 1.0
            for pkey in self.kernel_dict:
                self.net1.weight[pkey, 0] = ident
                self.net2.weight[0, pkey] = torch.from_numpy(np.asarray(self.kernel_dict[pkey]))
        self._keys_np_cache = None
        self._keys_dev = None
        self._ktab = mgfea.DeviceTable()
        self._global_pattern = None

    @property
    def _keys_np(self):
        """host uint8 key map (None for single-pattern meshes), built on first use"""
        if self.n_channel == 1:
            return None
        if self._keys_np_cache is None:
            mesh = self._mesh
            k = getattr(mesh, "pattern_keys", None)
            if k is None:
                # duck-typed reference-style mesh: rebuild the key map from its one-hot maps
                k = np.zeros(self.nnode_edge * self.nnode_edge, np.uint8)
                for pkey in self.kernel_dict:
                    k[np.asarray(mesh.global_pattern_center[pkey]).reshape(-1) != 0] = pkey
                k = k.reshape(self.nnode_edge, self.nnode_edge)
            self._keys_np_cache = k
        return self._keys_np_cache

    # -- reference attribute: (1, C, N, N) fp32 one-hot pattern masks (model.py:32-35); built only when asked for
    @property
    def global_pattern(self):
        if self._global_pattern is None:
            N = self.nnode_edge
            gp = torch.zeros((1, self.n_channel, N, N))
            if self._keys_np is None:
                gp[0, 0] = 1.0
            else:
                kt = torch.from_numpy(self._keys_np.astype(np.int64))
                gp.scatter_(1, kt[None, None], 1.0)
            self._global_pattern = gp
        return self._global_pattern

    def convert_global_pattern(self, global_pattern_center):
        gp = torch.zeros((1, self.n_channel, self.nnode_edge, self.nnode_edge))
        for pkey in self.kernel_dict:
            gp[0, pkey, :, :] = torch.from_numpy(np.asarray(global_pattern_center[pkey])).reshape(
                self.nnode_edge, self.nnode_edge)
        self._global_pattern = gp

    # -- device-side description shared with JacobiBlock and the V-cycle drivers
    def keys_dev(self):
        if self.n_channel == 1:
            return None
        if self._keys_dev is None:
            gen = getattr(self._mesh, "device_pattern_keys", None)  # closed-form meshes: setup kernel, no host pass
            self._keys_dev = gen() if gen is not None else mgfea.pack_keys(self._keys_np)
        return self._keys_dev

    def ktab_dev(self):
        """live net2 weights as a [C][9] device table"""
        return self._ktab.get(self.net2.weight)

    def grid_struct(self, fld: Field, invd=None, bc=None) -> Grid:
        g = Grid()
        g.N, g.pitch, g.plane = fld.N, fld.pitch, fld.plane
        g.npat = self.n_channel
        kd = self.keys_dev()
        g.keys = kd.data_ptr() if kd is not None else None
        g.key_pitch = kd.shape[1] if kd is not None else 0
        g.ktab = self.ktab_dev().data_ptr()
        g.invd = invd.data_ptr() if invd is not None else None
        if bc is not None:
            g.bc_idx, g.bc_val, g.bc_plane = bc[0].ptr, bc[1].ptr, (bc[0].plane if bc[0].B > 1 else 0)
        return g

    def forward(self, u):
        _, _, H, _ = u.shape
        uf = as_field(u)
        out = Field(uf.B, uf.N, uf.store.device)
        if H == self.nnode_edge:
            g = self.grid_struct(uf)
            check(lib().mgfea_stiffness_apply(g, uf.ptr, out.ptr, uf.B, stream_ptr()))
        elif self.n_channel == 1:
            # reference pads the mask with ones (model.py:27-28): plain conv with the single kernel on the larger array
            check(lib().mgfea_load_vector(self.ktab_dev().data_ptr(), uf.ptr, out.ptr, uf.N, uf.pitch, uf.plane, uf.B,
                                          stream_ptr()))
        else:
            raise mgfea.MgfeaError("KNet.forward on an array larger than the mesh is only defined for one pattern")
        return _like_input(u, out)

    def split_x(self, x):
        '''Split the field x based on the material phase'''
        xf = as_field(x)
        if xf.N != self.nnode_edge:
            raise mgfea.MgfeaError("split_x: field size differs from the mesh")
        out = torch.empty((xf.B, self.n_channel, xf.N, xf.N), dtype=torch.float32, device=xf.store.device)
        g = self.grid_struct(xf)
        check(lib().mgfea_split_x(g, xf.ptr, out.data_ptr(), xf.B, stream_ptr()))
        return out if x.is_cuda else out.cpu()


class FNet(nn.Module):
    def __init__(self, h):
        super(FNet, self).__init__()
        self.h = h
        self.net = nn.Conv2d(in_channels=1, out_channels=1, kernel_size=3, padding=1, bias=False)
        f_weights_np = np.array([[h * h / 36., h * h / 9., h * h / 36.],
                                 [h * h / 9., 4. * h * h / 9., h * h / 9.],
                                 [h * h / 36., h * h / 9., h * h / 36.]], dtype=np.float32).reshape(1, 1, 3, 3)
        self.net.weight = nn.Parameter(torch.from_numpy(f_weights_np))
        self._w = mgfea.DeviceTable()

    def forward(self, x):
        xf = as_field(x)
        out = Field(xf.B, xf.N, xf.store.device)
        check(lib().mgfea_load_vector(self._w.get(self.net.weight).data_ptr(), xf.ptr, out.ptr, xf.N, xf.pitch,
                                      xf.plane, xf.B, stream_ptr()))
        return _like_input(x, out)
