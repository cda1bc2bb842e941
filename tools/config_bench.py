"""Secondary BASELINE.json configs at full size: ms/cycle, algorithmic roofline fraction, residual history.
usage: config_bench.py [cfg2|cfg3|cfg3jac|all]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multigrid-feanet_b200")]
import numpy as np
import torch

import mgfea
from bench import algorithmic_bytes_per_cycle, hbm_peak
from FEANet.drivers import HNet, Multigrid, SingleGrid, _InterfaceSingleGrid
from FEANet.solver import LINEAR_4, VCycleEngine

which = sys.argv[1] if len(sys.argv) > 1 else "all"
peak, _ = hbm_peak()
G = os.path.join(ROOT, "tests", "golden")
OPS = np.load(os.path.join(G, "ops.npz"))


def time_cycles(eng, n=40):
    eng.refresh()
    eng._ctl_reset(0, -1.0, eng.max_cycles)
    eng._ensure_graph()
    zero = eng.ctl.clone()
    for _ in range(5):
        eng._graph.replay()
    eng.ctl.copy_(zero)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        eng._graph.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def report(name, eng, n, L, B, key_bytes, hist):
    ms = time_cycles(eng)
    balg = algorithmic_bytes_per_cycle(n, L, B=B, key_bytes=key_bytes)
    dof = (n + 1) ** 2 * B
    print(json.dumps({"config": name, "ms_per_cycle": ms, "cycles_per_s": 1e3 / ms, "gdof_per_s": dof / ms / 1e6,
                      "algorithmic_GB": balg / 1e9, "cycle_roofline_frac": balg / (ms * 1e-3) / 1e9 / peak,
                      "residual_history": hist}), flush=True)


if which in ("cfg2", "all"):
    # config 2: isotropic Poisson 1025^2, 8-level V-cycle, batch of 64 random RHS
    n, L, B = 1024, 8, 64
    jacs = [SingleGrid(2, n // 2 ** l).jac for l in range(L)]
    eng = VCycleEngine(jacs, B=B, conv_rule=mgfea.CONV_MAX)
    g = torch.Generator(device="cuda").manual_seed(0)
    F = torch.randn(B, 1, n + 1, n + 1, generator=g, device="cuda")
    f = SingleGrid(2, n).fnet(F)
    eng.set_u(torch.zeros(B, 1, n + 1, n + 1, device="cuda"))
    eng.set_f(f)
    r0 = torch.sqrt(eng.residual_sumsq()).cpu().numpy()
    h = eng.run(n_iter=8)
    hist = [float(np.max(x / r0)) for x in h]
    report("cfg2: iso 1025^2 x 64 RHS, 8 levels, V(1,1)", eng, n, L, B, 0, hist)

for tag, smoother in (("cfg3", "hjac"), ("cfg3jac", "jac")):
    if which in (tag, "all"):
        # config 3: two-material circle 1:100, 4097^2, 12 levels, learned (HNet) or Jacobi smoother, 16-ch linear R/P, w=[4,1]
        n, L, B = 4096, 12, 1
        grids = [_InterfaceSingleGrid(2, n // 2 ** l, prop=(1, 100), shape=0) for l in range(L)]
        hnet = HNet(3)
        hnet.load_state_dict({f"convLayers.{i}.weight": torch.from_numpy(OPS["hnet_w"][i]).reshape(1, 1, 3, 3) for i in range(3)})
        R16 = np.repeat((LINEAR_4 / np.float32(4.0)).reshape(1, 9), 16, 0)
        P4 = np.repeat(LINEAR_4.reshape(1, 9), 16, 0)
        eng = VCycleEngine([g.jac for g in grids], B=B, smoother=smoother, hnet=hnet, prolong="table", rtab=R16,
                           r_scale=4.0, ptab=P4, p_scale=1.0)
        f = grids[0].fnet(torch.ones(1, 1, n + 1, n + 1, device="cuda"))
        eng.set_u(torch.zeros(1, 1, n + 1, n + 1, device="cuda"))
        eng.set_f(f)
        r0 = float(torch.sqrt(eng.residual_sumsq().sum()).item())
        h = eng.run(n_iter=8)
        report(f"{tag}: two-phase circle 1:100, 4097^2, 12 levels, V(1,1) {smoother}, 16-ch R/P", eng, n, L, B, 1,
               [x / r0 for x in h])

if which in ("cfg3bil", "all"):
    # two-material circle 1:20, 4097^2, 12 levels, Jacobi, full weighting + bilinear prolongation (MM_Interface_error
    # operators without the level-0 quirk): the configuration the keyed streaming kernels serve
    n, L, B = 4096, 12, 1
    grids = [_InterfaceSingleGrid(2, n // 2 ** l, prop=(1, 20), shape=0) for l in range(L)]
    eng = VCycleEngine([g.jac for g in grids], B=B, smoother="jac")
    f = grids[0].fnet(torch.ones(1, 1, n + 1, n + 1, device="cuda"))
    eng.set_u(torch.zeros(1, 1, n + 1, n + 1, device="cuda"))
    eng.set_f(f)
    r0 = float(torch.sqrt(eng.residual_sumsq().sum()).item())
    h = eng.run(n_iter=8)
    report("cfg3bil: two-phase circle 1:20, 4097^2, 12 levels, V(1,1) jac, FW restriction + bilinear prolongation",
           eng, n, L, B, 1, [x / r0 for x in h])

for tag, n in (("cfg4_1gpu", 8192), ("cfg5_1gpu", 16384)):
    if which in (tag, "all", "big"):
        # configs 4 / 5 on ONE GPU (the row-slab numbers at 2/4/8 GPUs come from bench.py --gpus N): f = 0 model problem
        L, B = int(np.log2(n)), 1
        prob = Multigrid(n)
        eng = prob._engine(1, 1, 0, B=1)
        g = torch.Generator(device="cuda").manual_seed(123)
        eng.set_u(1.2e5 * torch.rand((1, 1, n + 1, n + 1), generator=g, device="cuda") + 1.3e5)
        eng.set_f(torch.zeros(1, 1, n + 1, n + 1, device="cuda"))
        r0 = float(torch.sqrt(eng.residual_sumsq().sum()).item())
        h = eng.run(n_iter=8)
        report(f"{tag}: iso Poisson {n + 1}^2, {L} levels, V(1,1), single RHS, 1 GPU", eng, n, L, B, 0, [x / r0 for x in h])
        del eng, prob
        torch.cuda.empty_cache()
