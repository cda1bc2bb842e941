"""Timeline of one replayed V-cycle: %globaltimer stamps before/after every fused-leg launch (mgfea_trace), captured
in a CUDA graph so that the stamps are tight.  Stamp kernels serialise the launches (no PDL overlap) and add ~1 us each.
usage: cycle_trace.py [n] [reps] [iso|jac|hjac]   (jac / hjac: two-phase 1:100 circle, 16-channel R/P, BASELINE config 3)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multigrid-feanet_b200")]
import numpy as np
import torch

import mgfea
from FEANet.drivers import Multigrid

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
mode = sys.argv[3] if len(sys.argv) > 3 else "iso"
np.random.seed(123)
if mode == "iso":
    prob = Multigrid(n)
    eng = prob._engine(1, 1, 0, B=1)
    eng.set_u(prob.initial_v.reshape(1, 1, n + 1, n + 1))
else:
    from FEANet.drivers import HNet, _InterfaceSingleGrid
    from FEANet.solver import LINEAR_4, VCycleEngine

    OPS = np.load(os.path.join(ROOT, "tests", "golden", "ops.npz"))
    Lv = int(np.log2(n))
    grids = [_InterfaceSingleGrid(2, n // 2 ** l, prop=(1, 100), shape=0) for l in range(Lv)]
    hnet = HNet(3)
    hnet.load_state_dict({f"convLayers.{i}.weight": torch.from_numpy(OPS["hnet_w"][i]).reshape(1, 1, 3, 3) for i in range(3)})
    R16 = np.repeat((LINEAR_4 / np.float32(4.0)).reshape(1, 9), 16, 0)
    P4 = np.repeat(LINEAR_4.reshape(1, 9), 16, 0)
    if mode == "jacbil":  # full weighting + bilinear prolongation: the keyed streaming kernels
        grids = [_InterfaceSingleGrid(2, n // 2 ** l, prop=(1, 20), shape=0) for l in range(Lv)]
        eng = VCycleEngine([g.jac for g in grids], B=1, smoother="jac")
    else:
        eng = VCycleEngine([g.jac for g in grids], B=1, smoother=mode, hnet=hnet, prolong="table", rtab=R16, r_scale=4.0,
                           ptab=P4, p_scale=1.0)
    eng.set_u(torch.zeros(1, 1, n + 1, n + 1, device="cuda"))
    eng.set_f(grids[0].fnet(torch.ones(1, 1, n + 1, n + 1, device="cuda")))
eng.refresh()
eng._ctl_reset(0, -1.0, eng.max_cycles)
eng.cycle(use_ctl=True)  # lazy init outside capture
torch.cuda.synchronize()
buf = torch.zeros(256, dtype=torch.int64, device="cuda")
mgfea.lib().mgfea_trace(buf.data_ptr(), 256)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    eng.cycle(use_ctl=True)
mgfea.lib().mgfea_trace(None, 0)
z = eng.ctl.clone()
acc = None
for r in range(reps + 3):
    eng.ctl.copy_(z)
    g.replay()
    torch.cuda.synchronize()
    t = buf.cpu().numpy().astype(np.int64)
    tail = t[128:192]
    t = t[:128]
    k = int((t > 0).sum())
    t = t[:k]
    if r >= 3:
        acc = t - t[0] if acc is None else acc + (t - t[0])
acc = acc / reps / 1e3
names = []
L = eng.L
for i in range(0, k, 2):
    print(f"launch {i // 2:2d}: start {acc[i]:8.2f} us  dur {acc[i + 1] - acc[i]:7.2f} us  gap after {(acc[i + 2] - acc[i + 1]) if i + 2 < k else 0:5.2f}")
print(f"total {acc[k - 1]:8.2f} us over {k // 2} launches")
tail = tail[tail > 0]
if len(tail) < 2:
    sys.exit(0)
print("tail stages (clock64 deltas, us at 1.965 GHz):", " ".join(f"{d / 1965.0:.2f}" for d in np.diff(tail)))
print(f"tail total {(tail[-1] - tail[0]) / 1965.0:.2f} us, {len(tail) - 1} stages")
