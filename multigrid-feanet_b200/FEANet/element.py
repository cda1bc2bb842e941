"""General per-element conductivity: -div(a grad u) = f with one conductivity per ELEMENT.

The reference's data model already carries such a field (`material`, (n, n) per sample: Data/dataset.py:71-104,
Data/TestPoisson/poisson2d_33x33.h5) but its operator only knows the 16 two-phase patterns of MeshCenterInterface
(FEANet/mesh.py:103-117).  The classes below keep the reference's module surface (KNet.forward, JacobiBlock.
jacobi_convolution / reset_boundary / d_mat, the Multigrid driver of MM_Model_convergence.ipynb cell 3) with the pattern
lookup replaced by the element values themselves; on a two-phase map they reproduce the pattern operator bit for bit on
every interior node.  Kernels: mgfea_elem_* (csrc/mgfea_elem.cuh).  Coarse levels: mean of the four child elements
(`mgfea_elem_coarsen`), or an explicit list of per-level fields.
"""
import math

import numpy as np
import torch
import torch.nn as nn

import mgfea
from mgfea import Field, as_field, check, lib, stream_ptr

from .geo import Geometry
from .jacobi import _is_default_ring
from .model import FNet, KNet, _like_input
from .mesh import MeshSquare
from .solver import FULL_WEIGHTING_16


def _element_field(material, N, dev):
    """(n, n) conductivities -> padded device field [1][N][pitch] (element (r, c) at [r][c], zero elsewhere)"""
    m = torch.as_tensor(material, dtype=torch.float32)
    if m.dim() == 4:
        m = m[0, 0]
    elif m.dim() == 3:
        m = m[0]
    if tuple(m.shape) != (N - 1, N - 1):
        raise mgfea.MgfeaError(f"material must be ({N - 1}, {N - 1}) (one value per element), got {tuple(m.shape)}")
    fld = Field(1, N, dev)
    fld.store[0, : N - 1, : N - 1].copy_(m.to(dev))
    return fld


class ElementKNet(nn.Module):
    """KNet for a per-element conductivity field: forward(u) = K u on all nodes, zero padding (FEANet/model.py:22-30)"""

    def __init__(self, material, nnode_edge=None, _field=None):
        super().__init__()
        self.dev = mgfea.require_cuda()
        if _field is not None:
            self.nnode_edge, self.a = _field.N, _field
        else:
            m = torch.as_tensor(material)
            self.nnode_edge = int(nnode_edge or (m.shape[-1] + 1))
            self.a = _element_field(m, self.nnode_edge, self.dev)
        self.n_channel = 1

    @property
    def material(self):
        n = self.nnode_edge - 1
        return self.a.store[0, :n, :n]

    def forward(self, u):
        uf = as_field(u)
        if uf.N != self.nnode_edge:
            raise mgfea.MgfeaError(f"field is {uf.N}^2, the element map is for {self.nnode_edge}^2 nodes")
        out = Field(uf.B, uf.N, uf.store.device)
        check(lib().mgfea_elem_stiffness_apply(self.a.ptr, uf.ptr, out.ptr, uf.N, uf.pitch, uf.plane, uf.B, stream_ptr()))
        return _like_input(u, out)

    def residual(self, u, f):
        """f - K u on all nodes (one kernel)"""
        uf, ff = as_field(u), as_field(f)
        out = Field(uf.B, uf.N, uf.store.device)
        check(lib().mgfea_elem_residual(self.a.ptr, uf.ptr, ff.ptr, out.ptr, uf.N, uf.pitch, uf.plane, uf.B, stream_ptr()))
        return _like_input(u, out)

    def coarsen(self):
        """the next coarser level's operator: conductivity = mean of the four child elements"""
        N = self.nnode_edge
        Nc = (N - 1) // 2 + 1
        ac = Field(1, Nc, self.dev)
        check(lib().mgfea_elem_coarsen(self.a.ptr, ac.ptr, N, self.a.pitch, ac.pitch, stream_ptr()))
        return ElementKNet(None, _field=ac)


class ElementJacobiBlock:
    """JacobiBlock (FEANet/jacobi.py:5-47) for ElementKNet: d = centre entry of each node's own kernel"""

    def __init__(self, Knet, omega=2 / 3., geometry_idx=None, boundary_value=None):
        self.Knet, self.omega, self.nnode_edge = Knet, omega, Knet.nnode_edge
        geo = Geometry(self.nnode_edge) if geometry_idx is None else None
        self.geometry_idx = geo.geometry_idx if geo else geometry_idx
        self.boundary_value = geo.boundary_value if geo else boundary_value
        self._default_bc = _is_default_ring(self.geometry_idx, self.boundary_value)
        self._bc = None
        self._d_mat = None

    @property
    def d_mat(self):
        """(1, 1, N, N) Jacobi diagonal (FEANet/jacobi.py:31-37): the centre entry of every node's own kernel, read off the
        operator itself with four colourings of unit impulses (impulses two nodes apart do not reach each other's 3x3
        stencil, so (K e)[i,j] at an impulse node is exactly its centre weight).  Built on demand; the smoother computes
        the same value in registers."""
        if self._d_mat is None:
            N = self.nnode_edge
            d = torch.zeros(1, 1, N, N, device=self.Knet.dev)
            for ci in range(2):
                for cj in range(2):
                    e = torch.zeros(1, 1, N, N, device=self.Knet.dev)
                    e[:, :, ci::2, cj::2] = 1.0
                    k = self.Knet(e)
                    d[:, :, ci::2, cj::2] = k[:, :, ci::2, cj::2]
            self._d_mat = d.cpu()
        return self._d_mat

    def _bc_fields(self):
        if self._bc is None:
            self._bc = (as_field(self.geometry_idx), as_field(self.boundary_value))
        return self._bc

    def reset_boundary(self, u):
        uf = as_field(u)
        out = Field(uf.B, uf.N, uf.store.device)
        g = mgfea.Grid()
        g.N, g.pitch, g.plane, g.npat = uf.N, uf.pitch, uf.plane, 1
        if not self._default_bc:
            idx, val = self._bc_fields()
            g.bc_idx, g.bc_val, g.bc_plane = idx.ptr, val.ptr, (idx.plane if idx.B > 1 else 0)
        check(lib().mgfea_reset_boundary(g, uf.ptr, out.ptr, uf.B, stream_ptr()))
        return _like_input(u, out)

    def smooth_fields(self, uf, ff, n_iter=1):
        """n_iter sweeps on padded fields; returns the Field holding the result"""
        if not self._default_bc:
            raise mgfea.MgfeaError("the element smoother implements the default Dirichlet ring (zero boundary values)")
        cur = uf
        for _ in range(n_iter):
            out = Field(uf.B, uf.N, uf.store.device)
            check(lib().mgfea_elem_smooth(self.Knet.a.ptr, cur.ptr, out.ptr, ff.ptr, float(np.float32(self.omega)), uf.N,
                                          uf.pitch, uf.plane, uf.B, stream_ptr()))
            cur = out
        return cur

    def jacobi_convolution(self, initial_u, forcing_term, n_iter=1):
        uf, ff = as_field(initial_u), as_field(forcing_term)
        if ff.B == 1 and uf.B > 1:
            ff = as_field(forcing_term.expand(uf.B, -1, -1, -1).contiguous())
        return _like_input(initial_u, self.smooth_fields(uf, ff, n_iter))


class ElementSingleGrid:
    """SingleGrid (MM_Model_convergence.ipynb cell 2) over an element conductivity map"""

    def __init__(self, size, n, knet):
        self.size, self.n, self.omega = size, n, 2 / 3.
        self.plate = Geometry(nnode_edge=n + 1)
        self.Knet = knet
        self.fnet = FNet(size / n)
        self.v = torch.zeros((1, 1, n + 1, n + 1), dtype=torch.float32)
        self.f = torch.zeros((1, 1, n + 1, n + 1), dtype=torch.float32)
        self.jac = ElementJacobiBlock(knet, self.omega, self.plate.geometry_idx, self.plate.boundary_value)

    def Relax(self, v, f, k):
        return self.jac.jacobi_convolution(v, f, n_iter=k)


class ElementMultigrid:
    """V(nu1,nu2) multigrid of MM_Model_convergence.ipynb cell 3 (full weighting x 4, bilinear prolongation, Jacobi
    omega = 2/3) with the per-element operator on every level.  `material`: (n, n) for the finest level (coarser levels by
    4-child averaging) or a list of per-level maps."""

    def __init__(self, n, material, final_level=None, batch=1):
        self.size, self.n = 2, n
        self.L = int(np.log2(n)) if final_level is None else final_level
        self.n_arr = np.array([int(n / (2. ** i)) for i in range(self.L)])
        self.dev = mgfea.require_cuda()
        if isinstance(material, (list, tuple)):
            knets = [ElementKNet(m) for m in material[: self.L]]
        else:
            knets = [ElementKNet(material)]
            for _ in range(1, self.L):
                knets.append(knets[-1].coarsen())
        self.grids = {i: ElementSingleGrid(self.size, int(self.n_arr[i]), knets[i]) for i in range(self.L)}
        self.initial_v = torch.zeros((1, 1, n + 1, n + 1), dtype=torch.float32)
        self.B = batch
        self._alloc(batch)
        self._rt = torch.from_numpy(FULL_WEIGHTING_16.reshape(1, 9).copy()).to(self.dev)
        iso = KNet(MeshSquare(2, 3))  # any single-pattern table: mgfea_restrict / the bilinear prolongation do not read it
        self._iso_tab = iso._ktab.get(iso.net2.weight.reshape(1, 9))
        self._graph = None
        self.sumsq = torch.zeros(batch, dtype=torch.float64, device=self.dev)

    def _alloc(self, B):
        N = [int(v) + 1 for v in self.n_arr]
        self.u = [Field(B, k, self.dev) for k in N]
        self.u_alt = [Field(B, k, self.dev) for k in N]
        self.f = [Field(B, k, self.dev) for k in N]
        self.r = [Field(B, k, self.dev) for k in N]

    def _grid(self, l):
        g = mgfea.Grid()
        fld = self.u[l]
        g.N, g.pitch, g.plane, g.npat = fld.N, fld.pitch, fld.plane, 1
        g.ktab = self._iso_tab.data_ptr()
        return g

    def _smooth(self, l, k):
        """k sweeps on level l: u[l] <-> u_alt[l] ping-pong, result back in u[l]"""
        a, B = self.grids[l].Knet.a, self.B
        om = float(np.float32(self.grids[l].omega))
        for _ in range(k):
            check(lib().mgfea_elem_smooth(a.ptr, self.u[l].ptr, self.u_alt[l].ptr, self.f[l].ptr, om, self.u[l].N,
                                          self.u[l].pitch, self.u[l].plane, B, stream_ptr()))
            self.u[l], self.u_alt[l] = self.u_alt[l], self.u[l]

    def cycle(self, v1=1, v2=1, first_level=0, want_norm=True):
        """one V(v1,v2) cycle on the device buffers (u[first_level], f[first_level] in place)"""
        L, B = self.L, self.B
        for l in range(first_level, L):
            if l > first_level:
                self.u[l].zero_()
            self._smooth(l, v1)
            if l < L - 1:
                a = self.grids[l].Knet.a
                check(lib().mgfea_elem_residual(a.ptr, self.u[l].ptr, self.f[l].ptr, self.r[l].ptr, self.u[l].N,
                                                self.u[l].pitch, self.u[l].plane, B, stream_ptr()))
                fc = self.f[l + 1]
                check(lib().mgfea_restrict(self._grid(l), self.r[l].ptr, fc.ptr, fc.pitch, fc.plane, self._rt.data_ptr(), 1,
                                           1, 4.0, None, B, stream_ptr()))
        for l in range(L - 1, first_level - 1, -1):
            if l < L - 1:
                check(lib().mgfea_prolong_correct_smooth(self._grid(l), self._grid(l + 1), self.u[l + 1].ptr,
                                                         self.u[l].ptr, self.u_alt[l].ptr, None, mgfea.PROLONG_BILINEAR,
                                                         None, 0, 0, 0.0, None, 0, 0, None, 0, B, stream_ptr()))
                self.u[l], self.u_alt[l] = self.u_alt[l], self.u[l]
            self._smooth(l, v2)
        if want_norm:
            return self.residual_sumsq(first_level)
        return None

    def residual_sumsq(self, l=0):
        a, B = self.grids[l].Knet.a, self.B
        check(lib().mgfea_elem_residual(a.ptr, self.u[l].ptr, self.f[l].ptr, self.r[l].ptr, self.u[l].N, self.u[l].pitch,
                                        self.u[l].plane, B, stream_ptr()))
        check(lib().mgfea_sumsq_interior(self.r[l].ptr, self.sumsq.data_ptr(), self.u[l].N, self.u[l].pitch,
                                         self.u[l].plane, B, stream_ptr()))
        return self.sumsq

    def _load(self, v, f):
        v, f = torch.as_tensor(v), torch.as_tensor(f)
        n1 = self.n + 1
        v = v.reshape(-1, 1, n1, n1) if v.dim() != 4 else v
        f = f.reshape(-1, 1, n1, n1) if f.dim() != 4 else f
        B = max(v.shape[0], f.shape[0])
        if B != self.B:
            self.B = B
            self._alloc(B)
            self.sumsq = torch.zeros(B, dtype=torch.float64, device=self.dev)
        self.u[0].view.copy_(v.to(dtype=torch.float32).expand(B, -1, -1, -1), non_blocking=True)
        self.f[0].view.copy_(f.to(dtype=torch.float32).expand(B, -1, -1, -1), non_blocking=True)

    def rec_V_cycle(self, l, v, f, v1v2=(1, 1)):
        if l != 0:
            raise mgfea.MgfeaError("ElementMultigrid.rec_V_cycle starts on level 0")
        self._load(v, f)
        self.cycle(v1v2[0], v1v2[1], 0, want_norm=False)
        self.grids[0].v = self.u[0].view.clone() if torch.as_tensor(v).is_cuda else self.u[0].view.cpu().contiguous()

    def Solve(self, v1v2=[1, 1], n_iter=None, EPS=None, max_cycles=256):
        """repeat V-cycles while (res > EPS or n < n_iter); EPS = absolute interior 2-norm over the whole batch.  Returns
        the residual list; the solution is left in self.grids[0].v"""
        if n_iter is None:
            if EPS is None:
                print("At least one of EPS and n_iter have to be assigned")
                return None
            n_iter = 0
        elif EPS is None:
            EPS = math.inf
        self._load(self.initial_v, self.grids[0].f)
        res, hist = 1.0, []
        while (res > EPS or len(hist) < n_iter) and len(hist) < max_cycles:
            ss = self.cycle(v1v2[0], v1v2[1])
            res = float(torch.sqrt(ss.sum()).item())
            hist.append(res)
        on_gpu = torch.is_tensor(self.initial_v) and self.initial_v.is_cuda
        self.grids[0].v = self.u[0].view.clone() if on_gpu else self.u[0].view.cpu().contiguous()
        return hist
