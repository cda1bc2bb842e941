"""Weighted-Jacobi smoother with Dirichlet reset on the GPU (reference: FEANet/jacobi.py)."""
import numpy as np
import torch

import mgfea
from mgfea import Field, as_field, check, lib, stream_ptr

from .model import _like_input


def _is_default_ring(geometry_idx, boundary_value):
    """True when the masks are the square ring with zero boundary values (geo.py:13-30): the kernels then use index
    arithmetic instead of reading mask fields.  Tagged tensors from FEANet.geo skip the O(N^2) comparison."""
    tg, tb = getattr(geometry_idx, "_mgfea_default_ring", None), getattr(boundary_value, "_mgfea_zero", None)
    if tg is not None and tb is not None and tg == geometry_idx._version and tb == boundary_value._version:
        return True  # untouched since FEANet.geo built them (an in-place edit bumps the version -> full comparison below)
    if geometry_idx.shape[0] != 1 or geometry_idx.shape != boundary_value.shape:
        return False
    g = geometry_idx.detach().cpu()
    ring = torch.ones_like(g)
    ring[..., 0, :] = 0
    ring[..., -1, :] = 0
    ring[..., :, 0] = 0
    ring[..., :, -1] = 0
    return bool(torch.equal(g, ring)) and not bool(boundary_value.detach().cpu().any())


class JacobiBlock():
    """ Define all the methods necessary for a CNN-based Jacobi iteration (Dirichlet boundary condition)

        Knet: neural network model for stiffness terms
        mesh: an object that define the mesh
        geometry_idx : tensor-like, shape = [*, *, n, n]; 1.0 for inner points 0.0 elsewhere.
        boundary_value: tensor-like, shape = [*, *, n, n]; desired values for boundary points 0.0 elsewhere.
    """

    def __init__(self, Knet, mesh, omega, geometry_idx, boundary_value):
        if geometry_idx is None:
            # internal light-weight form (FEANet.distributed at 8193^2 / 16385^2): default square ring with zero boundary
            # values, mask tensors materialised only if somebody reads the attributes
            self.nnode_edge = mesh.nnode_edge
            self._geometry_idx = self._boundary_value = None
            self._default_bc = True
        else:
            self.nnode_edge = geometry_idx.shape[2]
            self._geometry_idx = geometry_idx
            self._boundary_value = boundary_value
            self._default_bc = _is_default_ring(geometry_idx, boundary_value)
        self.omega = omega
        self.mesh = mesh
        self.Knet = Knet
        self._d_mat = None
        self._bc_fields = None
        # per-pattern omega/d exactly as torch evaluates `self.omega/self.d_mat` (jacobi.py:46): reciprocal, then
        # multiply by fl32(omega)
        diag = np.array([np.asarray(mesh.kernel_dict[k], dtype=np.float32)[1, 1] for k in sorted(mesh.kernel_dict)],
                        dtype=np.float32)
        self._diag = diag
        self._invd_np = ((np.float32(1.0) / diag).astype(np.float32) * np.float32(omega)).astype(np.float32)
        self._invd_dev = None

    @property
    def geometry_idx(self):
        if self._geometry_idx is None:
            from .geo import Geometry

            g = Geometry(self.nnode_edge)
            self._geometry_idx, self._boundary_value = g.geometry_idx, g.boundary_value
        return self._geometry_idx

    @geometry_idx.setter
    def geometry_idx(self, v):
        self._geometry_idx = v

    @property
    def boundary_value(self):
        if self._boundary_value is None:
            _ = self.geometry_idx
        return self._boundary_value

    @boundary_value.setter
    def boundary_value(self, v):
        self._boundary_value = v

    # -- reference attribute: (B or 1,1,N,N) Jacobi diagonal (jacobi.py:31-37); lazy, not used by the kernels
    @property
    def d_mat(self):
        if self._d_mat is None:
            d = torch.zeros_like(self.geometry_idx)
            keys = getattr(self.Knet, "_keys_np", None)
            if keys is None:
                d += float(self._diag[0])
            else:
                d += torch.from_numpy(self._diag[keys.astype(np.int64)]).to(d.device)
            self._d_mat = d
        return self._d_mat

    def compute_diagonal_matrix(self):
        self._d_mat = None
        return self.d_mat

    def invd_dev(self):
        if self._invd_dev is None:
            self._invd_dev = torch.from_numpy(self._invd_np).to(mgfea.require_cuda())
        return self._invd_dev

    def bc_fields(self):
        if self._default_bc:
            return None
        if self._bc_fields is None:
            self._bc_fields = (as_field(self.geometry_idx), as_field(self.boundary_value))
        return self._bc_fields

    def grid_struct(self, fld):
        return self.Knet.grid_struct(fld, self.invd_dev(), self.bc_fields())

    def reset_boundary(self, u):
        """ Reset values at the boundary of the domain """
        uf = as_field(u)
        out = Field(uf.B, uf.N, uf.store.device)
        check(lib().mgfea_reset_boundary(self.grid_struct(uf), uf.ptr, out.ptr, uf.B, stream_ptr()))
        return _like_input(u, out)

    def smooth_fields(self, uf, ff, n_iter=1, smoother=mgfea.SMOOTH_JACOBI, hw=None, nlayers=0):
        """n_iter sweeps on padded fields, temporally blocked inside one tile pass where the halo allows"""
        g = self.grid_struct(uf)
        dev = uf.store.device
        depth = (1 + nlayers) if smoother == mgfea.SMOOTH_HJACOBI else 1
        kmax = max(1, 8 // depth)
        cur, remaining = uf, n_iter
        while remaining > 0:
            k = min(kmax, remaining)
            out = Field(uf.B, uf.N, dev)
            check(lib().mgfea_smooth(g, cur.ptr, out.ptr, ff.ptr, k, smoother, hw, nlayers, uf.B, stream_ptr()))
            cur, remaining = out, remaining - k
        return cur

    def jacobi_convolution(self, initial_u, forcing_term, n_iter=1):
        """ Jacobi method iteration step defined as a convolution:
        u_new = omega/d_mat*residual + u, where residual = f - K*u (* is convolution operator here)
        note that the forcing_term should be already convoluted, i.e., forcing_term = fnet(f), when source term is f.
        `n_iter` accepts the older API that FEANet/multigrid.py:46 still calls."""
        uf, ff = as_field(initial_u), as_field(forcing_term)
        if ff.B != uf.B:
            if ff.B == 1:  # the reference broadcasts a single right-hand side over the batch (elementwise ops)
                ff = as_field(torch.as_tensor(forcing_term).reshape(1, 1, ff.N, ff.N).expand(uf.B, -1, -1, -1).contiguous())
            else:
                raise mgfea.MgfeaError("u and f batch sizes differ")
        out = self.smooth_fields(uf, ff, n_iter)
        return _like_input(initial_u, out)


class JacobiBlockPBC():
    """ Define all the methods necessary for a CNN-based Jacobi iteration (periodic boundary condition); currently only
        for homogeneous problems (reference: FEANet/jacobi.py:50-97)

        Knet: neural network model for stiffness terms
    """

    def __init__(self, mesh, Knet=None, omega=2. / 3.):
        self.nnode_edge = mesh.nnode_edge
        self.omega = omega
        self.mesh = mesh
        self.Knet = Knet
        if len(mesh.kernel_dict) != 1:
            raise mgfea.MgfeaError("JacobiBlockPBC is for homogeneous (single-pattern) meshes, like the reference's")
        self._w9 = np.asarray(mesh.kernel_dict[0], dtype=np.float32).reshape(9)
        self._invd = ((np.float32(1.0) / self._w9[4:5]).astype(np.float32) * np.float32(omega)).astype(np.float32)
        self._dev = None
        self._wtab = mgfea.DeviceTable()
        self.d_mat = torch.zeros((1, 1, self.nnode_edge, self.nnode_edge))
        self.compute_diagonal_matrix()

    def compute_diagonal_matrix(self):
        """ Compute diagonal matrix for Jacobi iteration """
        for pkey in self.mesh.kernel_dict:
            K_weights = torch.from_numpy(np.asarray(self.mesh.kernel_dict[pkey], dtype=np.float32))
            global_pattern = torch.from_numpy(np.asarray(self.mesh.global_pattern_center[pkey])).reshape(
                self.nnode_edge, self.nnode_edge)
            self.d_mat[0, 0, :, :] += global_pattern * K_weights[1, 1]

    def pbc_boundary(self, u):
        """ Expand the domain boundary to be periodic.  Input size [n+1, n+1]; output size [n+3, n+3] """
        u_central = u[:, :, :-1, :-1]
        return torch.nn.functional.pad(u_central, (1, 2, 1, 2), 'circular')

    def reset_boundary(self, u):
        """ Copy the boundary values to be the same.  Input size [n+1, n+1]; output size [n+1, n+1] """
        u_central = u[:, :, :-1, :-1]
        return torch.nn.functional.pad(u_central, (0, 1, 0, 1), 'circular')

    def _weights(self):
        """live kernel: the Knet's net2 weight when a Knet was given (user edits / load_state_dict), else the mesh table"""
        if self.Knet is not None and hasattr(self.Knet, "net2"):
            return self._wtab.get(self.Knet.net2.weight.reshape(-1, 9)[:1])
        return self._wtab.get(torch.from_numpy(self._w9.reshape(1, 9)))

    def jacobi_convolution(self, u, forcing_term, n_iter=1):
        """ Jacobi method iteration step defined as a convolution:
        u_new = omega/d_mat*residual + u, where residual = f - K*u on the periodically padded field.
        As in the reference the forcing_term must come padded to [n+3, n+3] (its Knet runs on the padded array). """
        uf = as_field(u)
        N = uf.N
        ff = as_field(forcing_term)
        if ff.N != N + 2 or ff.B != uf.B:
            raise mgfea.MgfeaError(f"forcing_term must be (B,1,{N + 2},{N + 2}): the reference's caller pads it")
        if self._dev is None:
            self._dev = torch.from_numpy(self._invd).to(uf.store.device)
        w = self._weights()
        cur = uf
        for _ in range(n_iter):
            out = Field(uf.B, N, uf.store.device)
            check(lib().mgfea_smooth_pbc(w.data_ptr(), self._dev.data_ptr(), cur.ptr, out.ptr, ff.ptr, N, uf.pitch,
                                         uf.plane, ff.pitch, ff.plane, uf.B, stream_ptr()))
            cur = out
        return _like_input(u, cur)
