"""Debug probe: run ONE tile program in its own process and compare with the oracle. usage: probe.py <case> <tma|cpasync> [N]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multigrid-feanet_b200")]
import numpy as np
import torch

import mgfea
from oracle import oracle as O

case, loader = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 65
mgfea.set_loader(loader == "tma")
from FEANet.drivers import SingleGrid, Multigrid
from FEANet.mesh import MeshCenterInterface, MeshSquare
from FEANet.model import KNet, FNet
from FEANet.solver import VCycleEngine

rs = np.random.RandomState(1)
B = 2
u = rs.standard_normal((B, 1, N, N)).astype(np.float32)
f = rs.standard_normal((B, 1, N, N)).astype(np.float32)
cu = lambda x: torch.from_numpy(x).cuda()


def cmp(got, ref, name):
    got = got.detach().cpu().numpy()
    if got.ndim == 4:
        got = got[:, 0]
    bad = (got != ref).sum()
    print(f"{case}/{loader}/N={N} {name}: mismatches {bad}/{ref.size} maxabs {np.abs(got - ref).max():.3e}", flush=True)


if case == "ku_iso":
    cmp(KNet(MeshSquare(2, N))(cu(u)), O.stiffness_apply(u, None, O.kernel_table([1.0], 1).reshape(1, 9)), "Ku")
elif case == "ku_keys":
    kt = O.kernel_table([1, 20], 16).reshape(16, 9)
    cmp(KNet(MeshCenterInterface(2, [1, 20], N))(cu(u)), O.stiffness_apply(u, O.pattern_keys(N, 0), kt), "Ku")
elif case in ("jac_iso", "jac_keys"):
    keyed = case.endswith("keys")
    g = SingleGrid(2, N - 1, mesh=MeshCenterInterface(2, [1, 20], N) if keyed else None)
    kt = O.kernel_table([1, 20], 16).reshape(16, 9) if keyed else O.kernel_table([1.0], 1).reshape(1, 9)
    keys = O.pattern_keys(N, 0) if keyed else None
    for k in (1, 2):
        cmp(g.jac.jacobi_convolution(cu(u), cu(f), n_iter=k), O.jacobi(u, f, keys, kt, O.inv_diag(2 / 3., kt[:, 4]), nsweeps=k), f"jac x{k}")
elif case == "restrict":
    mg = Multigrid(N - 1)
    cmp(mg.Restrict(cu(f)), O.restrict(f, None, O.FW16, None), "Restrict")
elif case == "prolong":
    mg = Multigrid(N - 1)
    Nc = (N - 1) // 2 + 1
    vc = rs.standard_normal((B, 1, Nc, Nc)).astype(np.float32)
    cmp(mg.Interpolate(cu(vc)), O.prolong_bilinear(vc, np.zeros((B, N, N), np.float32)), "Interpolate")
elif case == "norm":
    g = SingleGrid(2, N - 1)
    eng = VCycleEngine([g.jac], B=B)
    eng.set_u(cu(u)); eng.set_f(cu(f))
    ss = eng.residual_sumsq().cpu().numpy()
    ref = O.sumsq_interior(O.residual(u, f, None, O.kernel_table([1.0], 1).reshape(1, 9)))
    print(case, loader, "sumsq rel err", np.abs(ss - ref) / ref, flush=True)
elif case == "vcycle":
    n = N - 1
    L = int(np.log2(n))
    jacs = [SingleGrid(2, n // 2 ** l).jac for l in range(L)]
    eng = VCycleEngine(jacs, B=B)
    eng.set_u(cu(u)); eng.set_f(cu(f * 0.01))
    eng.cycle()
    torch.cuda.synchronize()
    ref = O.vcycle(O.make_levels(n), O.CycleCfg(), u[:, 0], f * 0.01)
    cmp(eng.solution, ref, "vcycle")
torch.cuda.synchronize()
print(case, loader, "done", flush=True)
