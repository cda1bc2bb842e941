"""V-cycle driver module with (16-channel, learnable) conv restriction / transposed-conv prolongation and scalar ratios
``w`` (reference: FEANet/multigrid.py), on the sm_100a engine."""
import numpy as np
import torch
import torch.nn as nn

import mgfea
from mgfea import Field, as_field, check, lib, stream_ptr

from .geo import Geometry
from .jacobi import JacobiBlock
from .mesh import MeshCenterInterface
from .model import FNet, KNet, _like_input
from .solver import VCycleEngine


class SingleGrid():
    '''Weighted Jacobi relaxation for a single grid of the two-phase plate (multigrid.py:12-47).
    n = number of intervals; f is already convoluted (f = fnet(ff)).'''

    def __init__(self, size, n):
        self.size = size
        self.n = n
        self.omega = 2 / 3.
        self.property = [1, 20]
        self.plate = Geometry(nnode_edge=n + 1)
        self.grid = MeshCenterInterface(size, prop=self.property, nnode_edge=n + 1)
        self.v = torch.zeros((1, 1, n + 1, n + 1), requires_grad=False, dtype=torch.float32)
        self.f = torch.zeros((1, 1, n + 1, n + 1), requires_grad=False, dtype=torch.float32)
        self.InstantiateFEANet()
        self.jac = JacobiBlock(self.Knet, self.grid, self.omega, self.plate.geometry_idx, self.plate.boundary_value)

    def IsCoarsest(self):
        return self.n == 2

    def InstantiateFEANet(self):
        self.Knet = KNet(self.grid)
        self.fnet = FNet(self.size / self.n)
        for param in self.Knet.parameters():
            param.requires_grad = False
        for param in self.fnet.parameters():
            param.requires_grad = False

    def Relax(self, v, f, num_sweeps_down):
        '''a fixed number of weighted Jacobi sweeps (the reference passes n_iter=..., multigrid.py:46)'''
        return self.jac.jacobi_convolution(v, f, n_iter=num_sweeps_down)


class RestrictionNet(nn.Module):
    '''16-channel restriction conv (one 3x3 kernel per material pattern), initialised with one kernel'''

    def __init__(self, linear_tensor_R):
        super(RestrictionNet, self).__init__()
        self.n_channel = 16
        self.net = nn.Conv2d(in_channels=self.n_channel, out_channels=1, kernel_size=3, stride=2, bias=False)
        with torch.no_grad():
            for i in range(self.n_channel):
                self.net.weight[0, i] = linear_tensor_R
        self._w = mgfea.DeviceTable()

    def forward(self, x_split):
        '''input (B,16,M,M) already split and sliced [1:-1,1:-1]; output (B,1,(M-1)/2,(M-1)/2) as nn.Conv2d(stride 2)'''
        x = x_split.detach()
        dev = mgfea.require_cuda()
        B, C, M, _ = x.shape
        xp = torch.zeros((B, C, M + 2, M + 2), dtype=torch.float32, device=dev)
        xp[:, :, 1:-1, 1:-1] = x.to(dev)
        Nc = (M + 1) // 2 + 1
        out = torch.empty((B, 1, Nc, Nc), dtype=torch.float32, device=dev)
        check(lib().mgfea_restrict_channels(xp.data_ptr(), out.data_ptr(), self._w.get(self.net.weight).data_ptr(), C,
                                            M + 2, B, stream_ptr()))
        out = out[:, :, 1:-1, 1:-1]
        return out if x_split.is_cuda else out.cpu()


class ProlongationNet(nn.Module):
    '''16-channel transposed conv (kernel 3, stride 2, padding 1)'''

    def __init__(self, linear_tensor_P):
        super(ProlongationNet, self).__init__()
        self.n_channel = 16
        self.net = nn.ConvTranspose2d(in_channels=self.n_channel, out_channels=1, kernel_size=3, stride=2, padding=1,
                                      bias=False)
        with torch.no_grad():
            for i in range(self.n_channel):
                self.net.weight[i, 0] = linear_tensor_P
        self._w = mgfea.DeviceTable()

    def forward(self, x_split):
        x = x_split.detach()
        dev = mgfea.require_cuda()
        B, C, Nc, _ = x.shape
        xc = x.to(device=dev, dtype=torch.float32).contiguous()
        out = torch.empty((B, 1, 2 * Nc - 1, 2 * Nc - 1), dtype=torch.float32, device=dev)
        check(lib().mgfea_prolong_channels(xc.data_ptr(), out.data_ptr(), self._w.get(self.net.weight).data_ptr(), C,
                                           Nc, B, stream_ptr()))
        return out if x_split.is_cuda else out.cpu()


class MultiGrid(nn.Module):
    '''Multigrid for the two-phase plate, n = finest grid size (multigrid.py:75-185); inference path only.'''

    def __init__(self, n, linear_tensor_R, linear_tensor_P, linear_ratio):
        super(MultiGrid, self).__init__()
        self.m0 = 2
        self.m = 6
        self.size = 2
        self.n = n
        self.L = int(np.log2(n))
        self.solution = []
        self.n_arr = self.SizeArray()
        self.grids = self.GridDict()
        self.conv = RestrictionNet(linear_tensor_R)
        self.deconv = ProlongationNet(linear_tensor_P)
        self.w = nn.Parameter(linear_ratio)
        self.conv.requires_grad_(True)
        self.deconv.requires_grad_(True)
        self.w.requires_grad_(False)
        self._eng = None
        self._eng_B = None

    def GridDict(self):
        return {i: SingleGrid(self.size, int(self.n_arr[i])) for i in range(self.L)}

    def SizeArray(self):
        return np.array([int(self.n / (2. ** i)) for i in range(self.L)])

    def Restrict(self, rF):
        '''restriction of an already split residual (B,16,N,N) to the next level, zero ring (no w[0] factor)'''
        rFC = self.conv(rF[:, :, 1:-1, 1:-1])
        return torch.nn.functional.pad(rFC, (1, 1, 1, 1), "constant", 0)

    def Interpolate(self, eFC):
        '''prolongation of an already split coarse correction (B,16,Nc,Nc) (no w[1] factor)'''
        return self.deconv(eFC)

    def _engine(self, B):
        if self._eng is None or self._eng_B != B:
            self._eng = VCycleEngine([self.grids[i].jac for i in range(self.L)], B=B, nu1=1, nu2=1, smoother="jac",
                                     prolong="table", rtab=self.conv.net.weight, ptab=self.deconv.net.weight,
                                     w_param=self.w, conv_rule=mgfea.CONV_MAX)
            self._eng_B = B
        return self._eng

    def _res_norms(self, x, f):
        xf, ff = as_field(x), as_field(f)
        ss = torch.zeros(xf.B, dtype=torch.float64, device=xf.store.device)
        check(lib().mgfea_residual_norm(self.grids[0].jac.grid_struct(xf), xf.ptr, ff.ptr, ss.data_ptr(), None, None,
                                        xf.B, stream_ptr()))
        return torch.sqrt(ss)

    def _training(self):
        """autograd is recording and R / P / w are trainable: the reference's training mode (multigrid.py:98-100)"""
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def qm(self, x):
        "Compute the convergence factor after m iterations"
        if self._training() or (torch.is_tensor(x) and x.requires_grad):
            # differentiable w.r.t. x (the last iterate): d||r||/dx = -K r / ||r|| on the interior, K symmetric
            r1 = _ResNormFn.apply(x, self.f, self)
            with torch.no_grad():
                r0 = self._res_norms(self.v_m0, self.f).to(r1.device)
            return torch.mean(torch.pow(r1 / r0, 1.0 / (self.m - self.m0 + 1))).to(torch.float32)
        r1, r0 = self._res_norms(x, self.f), self._res_norms(self.v_m0, self.f)
        return torch.mean(torch.pow(r1 / r0, 1.0 / (self.m - self.m0 + 1))).to(torch.float32).cpu()

    def random_sampling(self, v):
        d1, d2, d3, d4 = v.shape
        for i in range(d1):
            for j in range(d2):
                coef = 10 * np.random.rand(2) - 5
                v[i, j, :, :] = torch.from_numpy(coef[0] * np.random.random((d3, d4)) + coef[1])

    def forward(self, F):
        '''Input is RHS field F.  Under autograd with trainable R / P / w the LAST cycle is differentiable (the earlier
        ones are detached, as in the reference: multigrid.py:152-157), so `qm(forward(F)).backward()` trains the
        inter-grid kernels like the learn_intergrid notebooks do.'''
        self.f = self.grids[0].fnet(F)
        self.v = torch.zeros(tuple(F.shape), requires_grad=False, dtype=torch.float32)
        self.random_sampling(self.v)
        U = torch.clone(self.v).to(F.device)
        with torch.no_grad():
            for i in range(self.m - 1):
                U = self.iterate(U, self.f)
                if i == self.m0 - 1:
                    self.v_m0 = U.clone()
        if self._training():
            return self.iterate_grad(U, self.f)
        return self.iterate(U, self.f)

    def iterate(self, x, f):
        '''one V(1,1) cycle (fused kernels); x is the current solution on the finest grid.  The result is detached from
        autograd: use iterate_grad (or forward) to train R / P / w'''
        eng = self._engine(x.shape[0])
        eng.refresh()
        eng.set_u(x)
        eng.set_f(f)
        eng.cycle()
        out = eng.solution.clone() if x.is_cuda else eng.solution.cpu().contiguous()
        self.grids[0].v, self.grids[0].f = out, f
        return out

    def iterate_grad(self, x, f):
        '''the same cycle with a backward pass to conv.net.weight (R), deconv.net.weight (P) and w: run
        operator by operator so that the adjoint sweep finds every level's intermediate fields; bit-identical forward'''
        return _IterateFn.apply(x, f, self, self.conv.net.weight, self.deconv.net.weight, self.w)


def _zero_ring(t, N):
    t[:, 0, :] = 0
    t[:, N - 1, :] = 0
    t[:, :, 0] = 0
    t[:, :, N - 1:] = 0
    return t


class _ResNormFn(torch.autograd.Function):
    '''per-sample interior 2-norm of f - K x (multigrid.py:132-136), differentiable w.r.t. x'''

    @staticmethod
    def forward(ctx, x, f, mg):
        xf, ff = as_field(x.detach()), as_field(f.detach())
        jac = mg.grids[0].jac
        r = Field(xf.B, xf.N, xf.store.device)
        check(lib().mgfea_residual(jac.grid_struct(xf), xf.ptr, ff.ptr, r.ptr, xf.B, stream_ptr()))
        _zero_ring(r.store, xf.N)
        ss = torch.zeros(xf.B, dtype=torch.float64, device=r.store.device)
        check(lib().mgfea_sumsq_interior(r.ptr, ss.data_ptr(), xf.N, xf.pitch, xf.plane, xf.B, stream_ptr()))
        nrm = torch.sqrt(ss)
        ctx.r, ctx.nrm, ctx.jac, ctx.host = r, nrm, jac, not x.is_cuda
        return nrm.to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        r, jac = ctx.r, ctx.jac
        kr = Field(r.B, r.N, r.store.device)
        check(lib().mgfea_stiffness_apply(jac.grid_struct(r), r.ptr, kr.ptr, r.B, stream_ptr()))  # K^T r = K r
        coef = (-g.to(r.store.device, torch.float64) / ctx.nrm).to(torch.float32)
        gx = Field(r.B, r.N, r.store.device, store=kr.store * coef[:, None, None])
        out = gx.view.contiguous()
        return (out.cpu() if ctx.host else out), None, None


class _IterateFn(torch.autograd.Function):
    '''MultiGrid.iterate (multigrid.py:159-185) with reverse-mode differentiation through every level.
    Level j:  s_j = Jac(a_j, f_j)   (a_0 = x, a_j = 0)      r_j = f_j - K s_j       f_{j+1} = w0 R(r_j)
              coarsest: t = Jac(s, f)        up: c_j = s_j + w1 P(t_{j+1}),  t_j = Jac(c_j, f_j);   result t_0.
    Adjoints: Jac: g_w = m g, g_f = D^-1 g_w, g_u = m (g_w - K D^-1 g_w)   (K symmetric, D^-1 = omega / d per node);
    R / P: mgfea_restrict_adjoint / mgfea_prolong_adjoint; tables: mgfea_*_wgrad; w: <f_{j+1}, g> / w0, <c_j - s_j, g> / w1.'''

    @staticmethod
    def forward(ctx, x, f, mg, Rw, Pw, w):
        dev = mgfea.require_cuda()
        L, B = mg.L, x.shape[0]
        jacs = [mg.grids[j].jac for j in range(L)]
        rt = Rw.detach().to(dev, torch.float32).reshape(-1, 9).contiguous()
        pt = Pw.detach().to(dev, torch.float32).reshape(-1, 9).contiguous()
        w0, w1 = float(w[0]), float(w[1])
        a0, f0 = as_field(x.detach()), as_field(f.detach())
        fs, ss, rs, ts, cs = [f0], [], [], [None] * L, [None] * L

        def zeros(j):
            return Field(B, jacs[j].nnode_edge, dev)

        for j in range(L):
            s_j = jacs[j].smooth_fields(a0 if j == 0 else zeros(j), fs[j], 1)
            ss.append(s_j)
            if j < L - 1:
                g = jacs[j].grid_struct(s_j)
                r = zeros(j)
                check(lib().mgfea_residual(g, s_j.ptr, fs[j].ptr, r.ptr, B, stream_ptr()))
                fc = zeros(j + 1)
                check(lib().mgfea_restrict(g, r.ptr, fc.ptr, fc.pitch, fc.plane, rt.data_ptr(), rt.shape[0], 1, w0, None, B,
                                           stream_ptr()))
                rs.append(r)
                fs.append(fc)
        ts[L - 1] = jacs[L - 1].smooth_fields(ss[L - 1], fs[L - 1], 1)
        for j in range(L - 2, -1, -1):
            c = zeros(j)
            check(lib().mgfea_prolong_correct_smooth(jacs[j].grid_struct(c), jacs[j + 1].grid_struct(ts[j + 1]),
                                                     ts[j + 1].ptr, ss[j].ptr, c.ptr, None, mgfea.PROLONG_TABLE,
                                                     pt.data_ptr(), pt.shape[0], 1, w1, None, 0, 0, None, 0, B, stream_ptr()))
            cs[j] = c
            ts[j] = jacs[j].smooth_fields(c, fs[j], 1)
        ctx.mg, ctx.jacs, ctx.tabs, ctx.w01 = mg, jacs, (rt, pt), (w0, w1)
        ctx.fields = (fs, ss, rs, ts, cs)
        ctx.meta = (Rw.shape, Pw.shape, Rw.device, Pw.device, w.device, not x.is_cuda)
        out = ts[0].view.clone()
        return out.cpu() if not x.is_cuda else out

    @staticmethod
    def _invd_field(jac, fld):
        '''omega / d per node as a padded (1, N, pitch) device field'''
        cache = getattr(jac, "_invd_field", None)
        if cache is None:
            inv = jac.invd_dev()
            kd = jac.Knet.keys_dev()
            N, pitch = fld.N, fld.pitch
            out = torch.zeros((1, N, pitch), dtype=torch.float32, device=fld.store.device)
            out[0, :, :N] = inv[kd[:, :N].long()] if kd is not None else inv[0]
            cache = jac._invd_field = out
        return cache

    @staticmethod
    def _jac_adjoint(jac, g_out):
        '''g_out = dL/d Jac(u, f)  ->  (dL/du, dL/df), default Dirichlet ring with zero boundary values'''
        dev = g_out.store.device
        N = g_out.N
        gw = _zero_ring(g_out.store.clone(), N)
        gf = Field(g_out.B, N, dev, store=gw * _IterateFn._invd_field(jac, g_out))
        kg = Field(g_out.B, N, dev)
        check(lib().mgfea_stiffness_apply(jac.grid_struct(gf), gf.ptr, kg.ptr, g_out.B, stream_ptr()))
        gu = _zero_ring(gw - kg.store, N)
        return Field(g_out.B, N, dev, store=gu), gf

    @staticmethod
    def backward(ctx, g):
        mg, jacs = ctx.mg, ctx.jacs
        fs, ss, rs, ts, cs = ctx.fields
        rt, pt = ctx.tabs
        w0, w1 = ctx.w01
        L = mg.L
        dev = rt.device
        B = g.shape[0]
        if not jacs[0]._default_bc:
            raise mgfea.MgfeaError("iterate_grad: default Dirichlet ring only (the reference's MultiGrid)")
        accR = torch.zeros((rt.shape[0], 9), dtype=torch.float64, device=dev)
        accP = torch.zeros((pt.shape[0], 9), dtype=torch.float64, device=dev)
        gw0 = torch.zeros((), dtype=torch.float64, device=dev)
        gw1 = torch.zeros((), dtype=torch.float64, device=dev)
        g_t = as_field(g.detach().to(dev).contiguous())
        g_s, g_f = [None] * L, [None] * L
        # ---- up leg, fine -> coarse
        for j in range(L - 1):
            g_c, gf = _IterateFn._jac_adjoint(jacs[j], g_t)       # t_j = Jac(c_j, f_j)
            g_f[j] = gf.store
            g_s[j] = g_c.store                                     # c_j = s_j + w1 P(t_{j+1})
            gj, gcj = jacs[j].grid_struct(g_c), jacs[j + 1].grid_struct(ts[j + 1])
            check(lib().mgfea_prolong_wgrad(gj, gcj, pt.shape[0], w1, ts[j + 1].ptr, g_c.ptr, accP.data_ptr(), B, stream_ptr()))
            if w1 != 0.0:
                gw1 += ((cs[j].store - ss[j].store).double() * g_c.store.double()).sum() / w1
            g_tn = Field(B, jacs[j + 1].nnode_edge, dev)
            check(lib().mgfea_prolong_adjoint(gj, gcj, pt.data_ptr(), pt.shape[0], w1, g_c.ptr, g_tn.ptr, B, stream_ptr()))
            g_t = g_tn
        # ---- coarsest level: t = Jac(s, f), s = Jac(0, f)
        g_sL, gf = _IterateFn._jac_adjoint(jacs[L - 1], g_t)
        g_f[L - 1] = gf.store
        g_a, gf2 = _IterateFn._jac_adjoint(jacs[L - 1], g_sL)
        g_f[L - 1] = g_f[L - 1] + gf2.store
        # ---- down leg, coarse -> fine
        for j in range(L - 2, -1, -1):
            gfn = Field(B, jacs[j + 1].nnode_edge, dev, store=g_f[j + 1].contiguous())   # dL/d f_{j+1}
            gj, gcj = jacs[j].grid_struct(rs[j]), jacs[j + 1].grid_struct(gfn)
            check(lib().mgfea_restrict_wgrad(gj, gcj, rt.shape[0], w0, rs[j].ptr, gfn.ptr, accR.data_ptr(), B, stream_ptr()))
            if w0 != 0.0:
                gw0 += (fs[j + 1].store.double() * gfn.store.double()).sum() / w0
            g_r = Field(B, jacs[j].nnode_edge, dev)
            check(lib().mgfea_restrict_adjoint(gj, gcj, rt.data_ptr(), rt.shape[0], w0, gfn.ptr, g_r.ptr, B, stream_ptr()))
            kg = Field(B, jacs[j].nnode_edge, dev)                 # r_j = f_j - K s_j
            check(lib().mgfea_stiffness_apply(gj, g_r.ptr, kg.ptr, B, stream_ptr()))
            gs_tot = Field(B, jacs[j].nnode_edge, dev, store=g_s[j] - kg.store)
            g_a, gf3 = _IterateFn._jac_adjoint(jacs[j], gs_tot)    # s_j = Jac(a_j, f_j)
            g_f[j] = g_f[j] + g_r.store + gf3.store
        Rs, Ps, Rd, Pd, wd, host = ctx.meta
        gR = accR.to(torch.float32).reshape(Rs).to(Rd)
        gP = accP.to(torch.float32).reshape(Ps).to(Pd)
        gw = torch.stack([gw0, gw1]).to(torch.float32).to(wd)
        gx = g_a.view.contiguous()
        gf0 = Field(B, jacs[0].nnode_edge, dev, store=g_f[0].contiguous()).view.contiguous()
        return (gx.cpu() if host else gx), (gf0.cpu() if host else gf0), None, gR, gP, gw
