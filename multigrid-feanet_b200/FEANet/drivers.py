"""The reference's notebook drivers promoted to importable code, on top of the sm_100a V-cycle engine.

  Multigrid            MM_Model_convergence.ipynb cell 3  (iso Laplace, V(nu1,nu2), `Solve` == the solve() path)
  InterfaceMultigrid   MM_Interface_error.ipynb cell 2    (two-phase circle, pre-smooth always on level 0)
  HNet, HJacIterator   M-FEANet-mg_test.ipynb cells 4-5   (learned smoother, inference only)
  MGTestMultiGrid      M-FEANet-mg_test.ipynb cell 19     (`MultiGrid(n, hnet, P, mode)`: Step / forward)
  SingleGrid*          the per-notebook SingleGrid variants

Same constructor arguments, attributes and method semantics; tensors handed back follow the device of the inputs
(host tensors in -> H2D, kernels, D2H -> host tensors out).
"""
import math
import os
from functools import reduce

import numpy as np
import torch
import torch.nn as nn

import mgfea
from mgfea import Field, as_field, check, lib, stream_ptr

from .geo import Geometry
from .jacobi import JacobiBlock
from .mesh import MeshCenterInterface, MeshSquare
from .model import FNet, KNet, _like_input
from .solver import LINEAR_4, VCycleEngine


def compute_q(res_arr, m=None):
    if m is None:
        return res_arr[-1] / res_arr[-2]
    return res_arr[m] / res_arr[m - 1]


class SingleGrid():
    '''
    Weighted Jacobi relaxation for a single grid (MM_Model_convergence.ipynb cell 2; M-FEANet-mg_test.ipynb cell 3).
    n is the number of intervals: (n+1)*(n+1) nodes; f is already convoluted, f = fnet(ff).
    '''

    def __init__(self, size, n, mesh=None):
        self.size = size
        self.n = n
        self.omega = 2 / 3.
        self.plate = Geometry(nnode_edge=n + 1)
        self.grid = mesh if mesh is not None else MeshSquare(size, nnode_edge=n + 1)
        self.v = torch.zeros((1, 1, n + 1, n + 1), requires_grad=False, dtype=torch.float32)
        self.f = torch.zeros((1, 1, n + 1, n + 1), requires_grad=False, dtype=torch.float32)
        self.InstantiateFEANet()
        self.jac = JacobiBlock(self.Knet, self.grid, self.omega, self.plate.geometry_idx, self.plate.boundary_value)

    def IsCoarsest(self):
        return self.n == 2

    def ResetBoundary(self, bc_idx, bc_values):
        self.jac = JacobiBlock(self.Knet, self.grid, self.omega, bc_idx, bc_values)

    def InstantiateFEANet(self):
        self.Knet = KNet(self.grid)
        self.fnet = FNet(self.size / self.n)
        for param in self.Knet.parameters():
            param.requires_grad = False
        for param in self.fnet.parameters():
            param.requires_grad = False

    def Relax(self, *args):
        '''Relax(num_sweeps) updates self.v in place (MM_Model_convergence); Relax(v, f, num_sweeps) returns the
        smoothed field (MM_Interface_error / FEANet.multigrid signature).'''
        if len(args) == 1:
            self.v = self.jac.jacobi_convolution(self.v, self.f, n_iter=args[0])
            return None
        v, f, k = args
        return self.jac.jacobi_convolution(v, f, n_iter=k)


class _InterfaceSingleGrid(SingleGrid):
    def __init__(self, size, n, prop=(1, 20), shape=0):
        self.property = list(prop)
        super().__init__(size, n, mesh=MeshCenterInterface(size, prop=self.property, nnode_edge=n + 1, shape=shape))


def _solve_args(n_iter, EPS):
    if n_iter is None:
        if EPS is None:
            print("At least one of EPS and n_iter have to be assigned")
            return None
        return 0, EPS
    return n_iter, (np.inf if EPS is None else EPS)


class Multigrid():
    '''Geometric multigrid for the iso Laplace model problem, n = finest grid size
    (MM_Model_convergence.ipynb cell 3).'''
    _grid_cls = SingleGrid
    _quirk = False

    def __init__(self, n=64, final_level=None, batch=1, max_cycles=256):
        self.size = 2
        self.n = n
        self.L = int(np.log2(n)) if final_level is None else final_level
        self.n_arr = self.SizeArray()
        self.grids = self.GridDict()
        self.batch = batch
        self.max_cycles = max_cycles
        self.initial_v = torch.from_numpy(self.random_data())
        self.grids[0].f = torch.zeros((1, 1, n + 1, n + 1), requires_grad=False, dtype=torch.float32)
        self.v1, self.v2 = 1, 1
        self._engines = {}

    def _make_grid(self, n):
        return self._grid_cls(self.size, n)

    def GridDict(self):
        return {i: self._make_grid(int(self.n_arr[i])) for i in range(self.L)}

    def SizeArray(self):
        return np.array([int(self.n / (2. ** i)) for i in range(self.L)])

    # ---- engine plumbing
    def _engine(self, v1, v2, first_level=0, B=None):
        B = B or self.batch
        key = (v1, v2, first_level, B, tuple(id(self.grids[l].jac) for l in range(first_level, self.L)))
        eng = self._engines.get(key)
        if eng is None:
            eng = VCycleEngine([self.grids[l].jac for l in range(first_level, self.L)], B=B, nu1=v1, nu2=v2,
                               smoother="jac", prolong="bilinear", rtab=None, r_scale=4.0,
                               quirk_level0=self._quirk, max_cycles=self.max_cycles)
            self._engines = {k: e for k, e in self._engines.items() if k[2] != first_level or k[3] != B}
            self._engines[key] = eng
        return eng

    # ---- intergrid operators as standalone methods (API parity; the cycle uses the fused kernels)
    def Restrict(self, f):
        '''full weighting of f[1:-1,1:-1] to the next level, zero ring (no factor 4: the caller multiplies)'''
        ff = as_field(f)
        lvl = int(np.where(self.n_arr == ff.N - 1)[0][0])
        Nc = (ff.N - 1) // 2 + 1
        out = Field(ff.B, Nc, ff.store.device)
        rt = torch.from_numpy(np.ascontiguousarray(
            (np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], np.float32) / np.float32(16.0)).reshape(1, 9))).to(ff.store.device)
        g = self.grids[lvl].jac.grid_struct(ff)
        check(lib().mgfea_restrict(g, ff.ptr, out.ptr, out.pitch, out.plane, rt.data_ptr(), 1, 0, 0.0, None, ff.B,
                                   stream_ptr()))
        return _like_input(f, out)

    def Interpolate(self, v):
        '''bilinear x2 (align_corners) followed by the fine level's reset_boundary'''
        vf = as_field(v)
        N = 2 * vf.N - 1
        lvl = int(np.where(self.n_arr == N - 1)[0][0])
        zero = Field(vf.B, N, vf.store.device)
        out = Field(vf.B, N, vf.store.device)
        jf, jc = self.grids[lvl].jac, self.grids[lvl + 1].jac if lvl + 1 in self.grids else None
        g = jf.grid_struct(out)
        gc = jc.grid_struct(vf) if jc is not None else KNet(MeshSquare(self.size, vf.N)).grid_struct(vf)
        check(lib().mgfea_prolong_correct_smooth(g, gc, vf.ptr, zero.ptr, out.ptr, None, mgfea.PROLONG_BILINEAR, None,
                                                 0, 0, 0.0, None, 0, 0, None, 0, vf.B, stream_ptr()))
        return _like_input(v, out)

    # ---- cycles
    def _run_cycle(self, l, v, f):
        B = (v.shape[0] if v.dim() == 4 else 1) if torch.is_tensor(v) else None  # no H2D copy just to learn the batch size
        eng = self._engine(self.v1, self.v2, first_level=l, B=B)
        eng.refresh()
        eng.set_u(v)
        eng.set_f(f)
        eng.cycle()
        self.grids[l].v = eng.solution if (torch.is_tensor(v) and v.is_cuda) else eng.solution.cpu().contiguous()
        self.grids[l].f = f
        for k in range(l + 1, self.L):
            self.grids[k].v = torch.zeros_like(self.grids[k].v)

    def rec_V_cycle(self, l, v, f):
        '''one recursive V(v1,v2) cycle starting on level l; updates self.grids[l].v'''
        self._run_cycle(l, v, f)

    def V_cycle(self, x, f):
        '''iterative formulation of the same cycle (identical arithmetic)'''
        self._run_cycle(0, x, f)

    def random_data(self):
        coef = 100000 + 50000 * np.random.rand(2)
        return (coef[0] * np.random.random((self.n + 1, self.n + 1)).astype('f') + coef[1]).astype(np.float32)

    def Solve(self, v1v2=[1, 1], rec=True, n_iter=None, EPS=None, chunk=4, use_graph=True):
        """Repeat V-cycles while (res > EPS or n < n_iter); EPS is an ABSOLUTE interior 2-norm of f - K u over the
        whole batch.  Returns the per-cycle residual list; the solution is left in self.grids[0].v."""
        args = _solve_args(n_iter, EPS)
        if args is None:
            return None
        self.v1, self.v2 = v1v2
        n1 = self.n + 1
        v0 = torch.as_tensor(self.initial_v).reshape(-1, 1, n1, n1) if torch.as_tensor(self.initial_v).dim() != 4 \
            else self.initial_v
        eng = self._engine(self.v1, self.v2, 0, B=max(v0.shape[0], torch.as_tensor(self.grids[0].f).shape[0]))
        eng.set_u(v0)
        eng.set_f(self.grids[0].f)
        res = eng.run(n_iter=args[0], EPS=args[1], chunk=chunk, use_graph=use_graph)
        on_gpu = torch.is_tensor(self.initial_v) and self.initial_v.is_cuda
        self.grids[0].v = eng.solution if on_gpu else eng.solution_to_host()
        self.engine = eng
        return res

    def SolveMixed(self, v1v2=[1, 1], n_iter=None, EPS=None, chunk=4, use_graph=True):
        '''Solve with the iterate / right-hand side / residual in fp64 and the V-cycle itself in fp32 (defect correction,
        SURVEY 8f.1): reproduces the residual history the reference gives after `.double()` (MM_poisson.ipynb cell 5) and
        goes below the fp32 floor.  Result: self.grids[0].v as a float64 (1,1,N,N) tensor; returns the residual list.'''
        if n_iter is None and EPS is None:
            print("At least one of EPS and n_iter have to be assigned")
            return None
        self.v1, self.v2 = v1v2
        key = ("mixed", self.v1, self.v2, tuple(id(self.grids[l].jac) for l in range(self.L)))
        eng = self._engines.get(key)
        if eng is None:
            eng = VCycleEngine([self.grids[l].jac for l in range(self.L)], B=1, nu1=self.v1, nu2=self.v2, smoother="jac",
                               prolong="bilinear", rtab=None, r_scale=4.0, quirk_level0=self._quirk,
                               max_cycles=self.max_cycles, compute_norm=False, zero_guess=True)
            self._engines[key] = eng
        N = self.n + 1
        u0, f = torch.as_tensor(self.initial_v), torch.as_tensor(self.grids[0].f)
        u0 = u0 if u0.dim() == 4 else u0.reshape(1, 1, N, N)  # (4-D tensors are passed on as they are: our own padded
        f = f if f.dim() == 4 else f.reshape(1, 1, N, N)      # views are then used in place)
        on_host = not u0.is_cuda
        res = eng.run_mixed(u0, f, n_iter=n_iter, EPS=EPS, chunk=chunk, use_graph=use_graph)
        sol = eng.solution64
        self.grids[0].v = sol.cpu().contiguous() if on_host else sol
        self._mixed_engine = eng
        return res

    def solve_jacobi(self, n_iter=None, EPS=None):
        args = _solve_args(n_iter, EPS)
        if args is None:
            return None
        n1 = self.n + 1
        eng = VCycleEngine([self.grids[0].jac], B=1, nu1=1, nu2=0, max_cycles=max(self.max_cycles, args[0] + 1, 4096))
        eng.set_u(torch.as_tensor(self.initial_v).reshape(1, 1, n1, n1))
        eng.set_f(self.grids[0].f)
        res = eng.run(n_iter=args[0], EPS=args[1], chunk=16)
        self.grids[0].v = eng.solution.cpu().contiguous()
        return res


class InterfaceMultigrid(Multigrid):
    '''Two-phase circle a=[1,20], f = fnet(ones) (MM_Interface_error.ipynb cell 2).  Keeps that notebook's quirk:
    the pre-smoothing step of every level is applied to level 0 (`self.grids[0].Relax` at every depth).'''
    _quirk = True

    def __init__(self, n=64, final_level=None, prop=(1, 20), shape=0, quirk=True, max_cycles=256):
        self._prop, self._shape = prop, shape
        self._quirk = quirk
        super().__init__(n, final_level, max_cycles=max_cycles)
        self.initial_v = torch.zeros((1, 1, n + 1, n + 1), dtype=torch.float32)
        ff = torch.ones(1, 1, n + 1, n + 1)
        self.grids[0].f = self.grids[0].fnet(ff)

    def _make_grid(self, n):
        return _InterfaceSingleGrid(self.size, n, self._prop, self._shape)


# ---------------------------------------------------------------------------------------------------------
class HNet(nn.Module):
    '''learned correction H = (c_L * g) o ... o (c_1 * g), c_i 3x3 pad 1 no bias, g = geometry_idx (mg_test cell 4)'''

    def __init__(self, nb_layers):
        super(HNet, self).__init__()
        self.convLayers = nn.ModuleList([nn.Conv2d(1, 1, 3, padding=1, bias=False) for _ in range(nb_layers)])
        self._w = [mgfea.DeviceTable() for _ in range(nb_layers)]

    def forward(self, x, geo_idx):
        '''geo_idx: internal points 1; boundary points 0'''
        acc = as_field(x)
        geo = as_field(geo_idx)
        zero = Field(geo.B, geo.N, geo.store.device)
        g = mgfea.Grid()
        g.N, g.pitch, g.plane, g.npat = acc.N, acc.pitch, acc.plane, 1
        g.bc_idx, g.bc_val, g.bc_plane = geo.ptr, zero.ptr, (geo.plane if geo.B > 1 else 0)
        for i, layer in enumerate(self.convLayers):
            conv = Field(acc.B, acc.N, acc.store.device)
            check(lib().mgfea_load_vector(self._w[i].get(layer.weight).data_ptr(), acc.ptr, conv.ptr, acc.N, acc.pitch,
                                          acc.plane, acc.B, stream_ptr()))
            out = Field(acc.B, acc.N, acc.store.device)
            check(lib().mgfea_reset_boundary(g, conv.ptr, out.ptr, acc.B, stream_ptr()))  # conv * geo_idx (+ 0)
            acc = out
        return _like_input(x, acc)


class HJacIterator(nn.Module):
    '''Jacobi + learned correction: u <- J(u) + H(J(u) - u) (mg_test cell 5).  HRelax is the fused inference kernel;
    HRelaxGrad / TrainSingleEpoch / Train back-propagate through the sweeps to the HNet kernels.'''

    def __init__(self, n, size=2, hnet=None, grid=None, batch_size=5, max_epochs=1000, nb_layers=3,
                 model_name='iso_poisson_33x33', model_dir='Model/learn_iterator/iso_poisson'):
        super(HJacIterator, self).__init__()
        self.size, self.n = size, n
        self.batch_size, self.max_epochs = batch_size, max_epochs
        self.grid = SingleGrid(size, n) if grid is None else grid
        self.net = HNet(nb_layers) if hnet is None else hnet
        self.model_dir, self.model_name = model_dir, model_name
        self._hw = None
        self._hw_sid = None

    def _hw_dev(self):
        ws = [l.weight for l in self.net.convLayers]
        sid = tuple((w.data_ptr(), w._version) for w in ws)
        if sid != self._hw_sid:
            self._hw = torch.stack([w.detach().reshape(9).float().cpu() for w in ws]).to(mgfea.require_cuda())
            self._hw_sid = sid
        return self._hw

    def HRelax(self, v, f, num_sweeps_down):
        '''num_sweeps_down sweeps of the modified Jacobi iteration, fused in one tile pass per sweep'''
        uf, ff = as_field(v), as_field(f)
        hw = self._hw_dev()
        out = self.grid.jac.smooth_fields(uf, ff, num_sweeps_down, mgfea.SMOOTH_HJACOBI, hw.data_ptr(), hw.shape[0])
        return _like_input(v, out)

    def RandomSampling(self, x):
        return torch.randn_like(x)

    # ---- training (M-FEANet-learn_iterator.ipynb cell 8; mg_test cell 5): back-propagation through HRelax
    def HRelaxGrad(self, v, f, num_sweeps_down):
        """HRelax with a backward pass w.r.t. the HNet kernels (and v): the forward runs the sweeps operator by operator
        (Jacobi kernel, x = J(u) - u, three masked 3x3 correlations) keeping what the adjoint needs; the backward uses the
        same sm_100a kernels -- the adjoint of a zero-padded correlation is the correlation with the flipped kernel, the
        adjoint of K is K (symmetric), the weight gradients are `mgfea_corr9` reductions.  CUDA tensors in and out."""
        ws = [l.weight for l in self.net.convLayers]
        return _HRelaxFn.apply(v, f, self, int(num_sweeps_down), *ws)

    def TrainSingleEpoch(self, train_dataloader, k_range=(1, 20)):
        """one pass over `train_dataloader` (batches (u, f, bc_value, bc_index), host or CUDA tensors: the reference's
        DataLoader or FEANet.dataset.DeviceBatchLoader): Adadelta on MSELoss(reduction='sum') of HRelax(random u0, fnet(f), k)
        against the FEM solution, k drawn from k_range like the notebook's random.randint(1, 20)"""
        import random

        if not hasattr(self, "optimizer"):
            self.loss = nn.MSELoss(reduction='sum')
            self.optimizer = torch.optim.Adadelta(self.net.parameters())
        dev = mgfea.require_cuda()
        running_loss, i = 0., -1
        for i, data in enumerate(train_dataloader):
            u_train, f_train, bc_value_train, bc_index_train = [t.to(dev, dtype=torch.float32) for t in data]
            self.optimizer.zero_grad()
            k = random.randint(*k_range)
            self.grid.ResetBoundary(bc_index_train, bc_value_train)
            ff = self.grid.fnet(f_train)
            uu = self.RandomSampling(f_train)
            u_out = self.HRelaxGrad(uu, ff, k)
            loss_i = self.loss(u_out, u_train)
            loss_i.backward()
            self.optimizer.step()
            running_loss += loss_i.item()
        return running_loss / (i + 1)

    def Train(self, training_set, save=True):
        from .dataset import DeviceBatchLoader

        train_dataloader = DeviceBatchLoader(training_set, batch_size=self.batch_size, shuffle=True)
        loss_train = torch.zeros((self.max_epochs, 1))
        avg_loss = self.TrainSingleEpoch(train_dataloader)
        print('Step-0 loss:', avg_loss)
        loss_train[0] = avg_loss
        for epoch in range(1, self.max_epochs):
            avg_loss = self.TrainSingleEpoch(train_dataloader)
            if epoch % 50 == 0:
                print('Step-' + str(epoch) + ' loss:', avg_loss)
            if save:  # save the model's state (mg_test cell 5)
                os.makedirs(self.model_dir, exist_ok=True)
                torch.save(self.net.state_dict(), os.path.join(self.model_dir, self.model_name + '.pth'))
            loss_train[epoch] = avg_loss
        return loss_train


def _flip9(w):
    return torch.flip(w.detach().reshape(3, 3), (0, 1)).contiguous().reshape(1, 9)


class _HRelaxFn(torch.autograd.Function):
    """u_{s+1} = J(u_s) + H(J(u_s) - u_s), H = (c3 . m) o (c2 . m) o (c1 . m)  (mg_test cells 4-5), `k` sweeps"""

    @staticmethod
    def _conv(fld, w9_dev):
        out = Field(fld.B, fld.N, fld.store.device)
        check(lib().mgfea_load_vector(w9_dev.data_ptr(), fld.ptr, out.ptr, fld.N, fld.pitch, fld.plane, fld.B, stream_ptr()))
        return out

    @staticmethod
    def forward(ctx, v, f, it, k, *ws):
        dev = mgfea.require_cuda()
        jac = it.grid.jac
        uf, ff = as_field(v.detach()), as_field(f.detach())
        m = as_field(jac.geometry_idx.to(dev)).store  # (B or 1, N, pitch): 1 inside, 0 on the Dirichlet nodes
        wd = [w.detach().to(dev, torch.float32).reshape(1, 9).contiguous() for w in ws]
        saved = []
        u = uf
        for _ in range(k):
            vj = jac.smooth_fields(u, ff, 1)                        # J(u): one Jacobi sweep (mgfea_smooth)
            x = Field(u.B, u.N, dev, store=vj.store - u.store)      # fields are zero in the padding columns
            h, acts = x, [x]
            for w9 in wd:
                z = _HRelaxFn._conv(h, w9)
                h = Field(u.B, u.N, dev, store=z.store * m)
                acts.append(h)
            saved.append(acts[:3])                                  # x, h1, h2: inputs of the three layers
            u = Field(u.B, u.N, dev, store=vj.store + acts[3].store)
        ctx.it, ctx.k, ctx.saved, ctx.m, ctx.wd, ctx.wdev = it, k, saved, m, wd, [w.device for w in ws]
        ctx.host_out = not v.is_cuda
        return _like_input(v, u).clone() if v.is_cuda else _like_input(v, u)

    @staticmethod
    def backward(ctx, g_out):
        it, k, m, wd = ctx.it, ctx.k, ctx.m, ctx.wd
        jac = it.grid.jac
        dev = m.device
        g = as_field(g_out.detach().to(dev).contiguous())
        acc = torch.zeros((len(wd), 9), dtype=torch.float64, device=dev)
        flipped = [_flip9(w) for w in wd]
        s = float(jac._invd_np[0])  # omega / d (single pattern: the HNet iterators are iso, like the reference's)
        if jac.Knet.n_channel != 1:
            raise mgfea.MgfeaError("HRelaxGrad: single-pattern meshes only (the reference trains on MeshSquare)")
        gs = g.store
        for sweep in range(k - 1, -1, -1):
            acts = ctx.saved[sweep]
            gh = gs                                                  # dL/dh3
            for l in range(len(wd) - 1, -1, -1):
                gz = Field(g.B, g.N, dev, store=gh * m)              # through the mask
                check(lib().mgfea_corr9(acts[l].ptr, gz.ptr, acc[l].data_ptr(), g.N, gz.pitch, gz.plane, g.B, stream_ptr()))
                gh = _HRelaxFn._conv(gz, flipped[l]).store           # adjoint of the correlation: flipped kernel
            gx = gh
            gv = gs + gx                                             # u' = v + H(x), x = v - u
            gw = Field(g.B, g.N, dev, store=gv * m)                  # v = reset(w): mask (the boundary values are constants)
            kg = Field(g.B, g.N, dev)
            check(lib().mgfea_stiffness_apply(jac.grid_struct(gw), gw.ptr, kg.ptr, g.B, stream_ptr()))  # K^T = K
            gs = (gw.store - s * kg.store) * m - gx                  # w = u~ + s (f - K u~), u~ = reset(u)
            gs[:, :, g.N:] = 0                                       # K leaves nothing there, keep the padding clean
        gu = Field(g.B, g.N, dev, store=gs)
        grad_v = _like_input(g_out, gu)
        gws = [acc[l].to(torch.float32).reshape(1, 1, 3, 3).to(ctx.wdev[l]) for l in range(len(wd))]
        return (grad_v, None, None, None) + tuple(gws)


class RestrictionNet1(nn.Module):
    '''1-channel restriction conv of mg_test cell 18'''

    def __init__(self, tensor_R):
        super().__init__()
        self.n_channel = 1
        self.net = nn.Conv2d(in_channels=1, out_channels=1, kernel_size=3, stride=2, bias=False)
        with torch.no_grad():
            self.net.weight[0, 0] = tensor_R


class ProlongationNet1(nn.Module):
    '''1-channel transposed conv of mg_test cell 18'''

    def __init__(self, tensor_P):
        super().__init__()
        self.n_channel = 1
        self.net = nn.ConvTranspose2d(in_channels=1, out_channels=1, kernel_size=3, stride=2, padding=1, bias=False)
        with torch.no_grad():
            self.net.weight[0, 0] = tensor_P


class MGTestMultiGrid(nn.Module):
    '''`MultiGrid(n, hnet, P, mode)` of M-FEANet-mg_test.ipynb cell 19: V(1,1) with conv / transposed-conv intergrid
    operators (R = P = [1 2 1;2 4 2;1 2 1]/4, factor 4 folded in), Jacobi ('jac') or learned ('hjac') smoother, data
    Dirichlet BCs on level 0.'''

    def __init__(self, n, hnet, P, mode='jac'):
        super().__init__()
        self.size = 2
        self.n = n
        self.L = int(np.log2(n))
        self.hnet = hnet
        self.mode = mode
        self.n_arr = self.SizeArray()
        self.iterators = self.IteratorDict()
        self.conv = RestrictionNet1(P)
        self.deconv = ProlongationNet1(P)
        self.conv.requires_grad_(False)
        self.deconv.requires_grad_(False)
        self._eng = None
        self._eng_key = None

    def IteratorDict(self):
        return {i: HJacIterator(size=self.size, hnet=self.hnet, n=int(self.n_arr[i])) for i in range(self.L)}

    def SizeArray(self):
        return np.array([int(self.n / (2. ** i)) for i in range(self.L)])

    def _engine(self, B):
        key = (B, self.mode, tuple(id(self.iterators[i].grid.jac) for i in range(self.L)))
        if key != self._eng_key:
            self._eng = VCycleEngine([self.iterators[i].grid.jac for i in range(self.L)], B=B, nu1=1, nu2=1,
                                     smoother=self.mode, hnet=self.hnet, prolong="table", rtab=self.conv.net.weight,
                                     r_scale=None, ptab=self.deconv.net.weight, p_scale=None,
                                     conv_rule=mgfea.CONV_MAX)
            self._eng_key = key
        return self._eng

    def Relax(self, iter, u, f, n_iter):
        if self.mode == 'jac':
            return iter.grid.jac.jacobi_convolution(u, f)
        return iter.HRelax(u, f, n_iter)

    def Step(self, v, f):
        '''one V(1,1) cycle; v, f (B,1,N,N) on the finest grid'''
        B = v.shape[0]
        eng = self._engine(B)
        eng.refresh()
        eng.set_u(v)
        eng.set_f(f)
        eng.cycle()
        out = eng.solution.clone() if v.is_cuda else eng.solution.cpu().contiguous()
        self.iterators[0].grid.v = out
        self.iterators[0].grid.f = f
        return out

    def forward(self, u0, F, bc_idx, bc_value, k):
        '''initial solution u0, RHS field F, Dirichlet masks, k cycles'''
        self.f = self.iterators[0].grid.fnet(F)
        self.iterators[0].grid.ResetBoundary(bc_idx, bc_value)
        self.u0 = self.iterators[0].grid.jac.reset_boundary(u0)
        U = self.u0.clone()
        for _ in range(k - 1):
            U = self.Step(U, self.f)
        self.last_v = U.clone()
        return self.Step(U, self.f)

    def residual_norms(self, u):
        '''per-sample interior 2-norm of f - K u (torch.norm(res[:, :, 1:-1, 1:-1], dim=(2,3)) of the notebook)'''
        uf, ff = as_field(u), as_field(self.f)
        ss = torch.zeros(uf.B, dtype=torch.float64, device=uf.store.device)
        g = self.iterators[0].grid.jac.grid_struct(uf)
        check(lib().mgfea_residual_norm(g, uf.ptr, ff.ptr, ss.data_ptr(), None, None, uf.B, stream_ptr()))
        return torch.sqrt(ss).to(torch.float32).reshape(-1, 1)

    def loss(self, uk, k=None):
        return torch.mean(self.residual_norms(uk) / self.residual_norms(self.last_v))

    def solve(self, u, f=None, EPS=5e-5, max_cycles=100):
        '''the notebook's convergence loop (cells 21-22) on the device: cycles until every sample's |r| <= EPS'''
        f = self.f if f is None else f
        eng = self._engine(u.shape[0])
        eng.set_u(u)
        eng.set_f(f)
        hist = eng.run(EPS=EPS, max_cycles=max_cycles)
        return (eng.solution.clone() if u.is_cuda else eng.solution.cpu().contiguous()), hist
