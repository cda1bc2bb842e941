"""three eager config-3 V-cycles (two-phase circle 1:100, HNet smoother, 16-channel table R/P; no graph) for ncu"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multigrid-feanet_b200")]
import numpy as np
import torch

from FEANet.drivers import HNet, _InterfaceSingleGrid
from FEANet.solver import LINEAR_4, VCycleEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
smoother = sys.argv[2] if len(sys.argv) > 2 else "hjac"
L = int(np.log2(n))
hw = np.load(os.path.join(ROOT, "tests", "golden", "ops.npz"))["hnet_w"]
iso = len(sys.argv) > 3 and sys.argv[3] == "iso"  # single-pattern mesh: the chain without any per-node lookups
if iso:
    from FEANet.drivers import SingleGrid

    grids = [SingleGrid(2, n // 2 ** l) for l in range(L)]
else:
    grids = [_InterfaceSingleGrid(2, n // 2 ** l, prop=(1, 100), shape=0) for l in range(L)]
hnet = HNet(3)
hnet.load_state_dict({f"convLayers.{i}.weight": torch.from_numpy(hw[i]).reshape(1, 1, 3, 3) for i in range(3)})
R16 = np.repeat((LINEAR_4 / np.float32(4.0)).reshape(1, 9), 16, 0)
P4 = np.repeat(LINEAR_4.reshape(1, 9), 16, 0)
eng = VCycleEngine([g.jac for g in grids], B=1, smoother=smoother, hnet=hnet, prolong="table", rtab=R16, r_scale=4.0,
                   ptab=P4, p_scale=1.0)
eng.set_u(torch.zeros(1, 1, n + 1, n + 1, device="cuda"))
eng.set_f(grids[0].fnet(torch.ones(1, 1, n + 1, n + 1, device="cuda")))
for _ in range(3):
    eng.cycle()
torch.cuda.synchronize()
print("ok", float(eng.sumsq.sum().item()))
