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

    def _inference_only(self, what):
        """The reference trains R / P / w by back-propagating through this cycle (multigrid.py:98-100,145-157); the fused
        sm_100a cycle has no backward pass (not built).  Returning a silently detached tensor would make a
        training loop fail far away ("does not require grad") or, worse, train nothing -- fail HERE instead."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise mgfea.MgfeaError(
                f"MultiGrid.{what}: inference only -- the CUDA V-cycle is not differentiable, but conv / deconv / w "
                "require grad and autograd is recording.  Wrap the call in torch.no_grad() (or "
                "requires_grad_(False)) to solve; training the inter-grid operators is not supported by this package.")

    def qm(self, x):
        "Compute the convergence factor after m iterations"
        self._inference_only("qm")
        r1, r0 = self._res_norms(x, self.f), self._res_norms(self.v_m0, self.f)
        return torch.mean(torch.pow(r1 / r0, 1.0 / (self.m - self.m0 + 1))).to(torch.float32).cpu()

    def random_sampling(self, v):
        d1, d2, d3, d4 = v.shape
        for i in range(d1):
            for j in range(d2):
                coef = 10 * np.random.rand(2) - 5
                v[i, j, :, :] = torch.from_numpy(coef[0] * np.random.random((d3, d4)) + coef[1])

    def forward(self, F):
        '''Input is RHS field F'''
        self._inference_only("forward")
        self.f = self.grids[0].fnet(F)
        self.v = torch.zeros(tuple(F.shape), requires_grad=False, dtype=torch.float32)
        self.random_sampling(self.v)
        U = torch.clone(self.v).to(F.device)
        for i in range(self.m - 1):
            U = self.iterate(U, self.f)
            if i == self.m0 - 1:
                self.v_m0 = U.clone()
        return self.iterate(U, self.f)

    def iterate(self, x, f):
        '''one V(1,1) cycle; x is the current solution on the finest grid.  The result is DETACHED from autograd (see
        _inference_only): iterate is the solve() building block, forward / qm are the training entry points and refuse'''
        eng = self._engine(x.shape[0])
        eng.refresh()
        eng.set_u(x)
        eng.set_f(f)
        eng.cycle()
        out = eng.solution.clone() if x.is_cuda else eng.solution.cpu().contiguous()
        self.grids[0].v, self.grids[0].f = out, f
        return out
