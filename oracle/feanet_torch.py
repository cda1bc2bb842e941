"""CPU torch restatement of the reference V-cycle that issues the SAME ATen calls as the reference
(conv2d 1->C identity "split", mask multiply, conv2d C->1, elementwise Jacobi update with two reset_boundary passes,
strided conv2d restriction, F.interpolate bilinear prolongation, torch.sum norm + .item()).

TEST INFRASTRUCTURE ONLY (cpu_baseline / `bench.py --impl reference` legs and tests): this is the reference's CPU torch
path timed on the box's host cores, because /root/reference itself does not travel to the GPU box.
Parity status: PINNED -- tests/test_oracle_golden.py::test_feanet_torch_matches_reference_histories compares it with the
golden histories produced by the unmodified reference (bit-identical op sequence => identical numbers on one torch build).

Reference lines: FEANet/model.py:22-30, FEANet/jacobi.py:27-47, MM_Model_convergence.ipynb cell 3
(Restrict, Interpolate, rec_V_cycle, Solve).
"""
import numpy as np
import torch
import torch.nn.functional as F


class Level:
    def __init__(self, N, ktab, keys=None, omega=2.0 / 3.0):
        C = ktab.shape[0]
        self.N, self.C = N, C
        self.w1 = torch.zeros(C, 1, 3, 3)
        self.w1[:, 0, 1, 1] = 1.0
        self.w2 = torch.from_numpy(np.ascontiguousarray(ktab, dtype=np.float32)).reshape(1, C, 3, 3)
        if keys is None:
            self.gp = torch.ones(1, 1, N, N)
            d = torch.full((1, 1, N, N), float(ktab[0, 4]))
        else:
            kt = torch.from_numpy(keys.astype(np.int64))
            self.gp = torch.zeros(1, C, N, N).scatter_(1, kt[None, None], 1.0)
            d = torch.from_numpy(np.ascontiguousarray(ktab[:, 4], dtype=np.float32))[kt][None, None]
        self.d_mat = d.contiguous()
        self.omega = omega
        g = torch.ones(1, 1, N, N)
        g[0, 0, 0, :] = 0
        g[0, 0, -1, :] = 0
        g[0, 0, :, 0] = 0
        g[0, 0, :, -1] = 0
        self.idx, self.bval = g, torch.zeros_like(g)

    def K(self, u):
        return F.conv2d(F.conv2d(u, self.w1, padding=1) * self.gp, self.w2, padding=1)

    def reset(self, u):
        return u * self.idx + self.bval

    def jacobi(self, u, f):
        u = self.reset(u)
        residual = f - self.K(u)
        u_new = self.omega / self.d_mat * residual + u
        return self.reset(u_new)


_RK = torch.tensor([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=torch.float32) / 16.0


def restrict(r):
    return F.pad(F.conv2d(r[:, :, 1:-1, 1:-1], _RK.view(1, 1, 3, 3), stride=2), (1, 1, 1, 1), "constant", 0)


def vcycle(levels, v, f, nu1=1, nu2=1, l=0):
    lv = levels[l]
    for _ in range(nu1):
        v = lv.jacobi(v, f)
    if l < len(levels) - 1:
        r = f - lv.K(v)
        fc = 4 * restrict(r)
        vc = torch.zeros(v.shape[0], 1, levels[l + 1].N, levels[l + 1].N)
        vc = vcycle(levels, vc, fc, nu1, nu2, l + 1)
        e = F.interpolate(vc, size=lv.N, mode="bilinear", align_corners=True)
        v = v + lv.reset(e)
    for _ in range(nu2):
        v = lv.jacobi(v, f)
    return v


def make_levels(n, L=None, prop=None, keys_fn=None, ktab_fn=None):
    from . import oracle as O

    L = int(np.log2(n)) if L is None else L
    out = []
    for l in range(L):
        N = int(n / 2.0 ** l) + 1
        if prop is None:
            out.append(Level(N, O.kernel_table([1.0], 1).reshape(1, 9)))
        else:
            out.append(Level(N, O.kernel_table(prop, 16).reshape(16, 9), O.pattern_keys(N, 0)))
    return out


def solve(levels, u0, f, n_iter=None, EPS=None, nu1=1, nu2=1):
    if n_iter is None:
        n_iter = 0
    elif EPS is None:
        EPS = np.inf
    v, res, hist = u0, 1.0, []
    with torch.no_grad():
        while res > EPS or len(hist) < n_iter:
            v = vcycle(levels, v, f, nu1, nu2)
            r = f - levels[0].K(v)
            res = torch.sqrt(torch.sum(r[:, :, 1:-1, 1:-1] ** 2)).item()
            hist.append(res)
    return v, hist
