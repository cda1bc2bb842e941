"""ms per V-cycle of the learned-smoother cycle (HNet, 16-channel table R/P) at 4097^2, graph replay, for several
thresholds of the streaming HNet kernels (hstream_min_n; 0 = tile programs only).  argv[1] = iso | keys"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multigrid-feanet_b200")]
import numpy as np
import torch

import mgfea
from bench import time_engine
from FEANet.drivers import HNet, SingleGrid, _InterfaceSingleGrid
from FEANet.solver import LINEAR_4, VCycleEngine

kind = sys.argv[1] if len(sys.argv) > 1 else "iso"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
L = int(np.log2(n))
hw = np.load(os.path.join(ROOT, "tests", "golden", "ops.npz"))["hnet_w"]
out = {"kind": kind, "n": n}
for thr in [int(a) for a in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["0", "4097", "2049", "129"])]:
    mgfea.set_option("hstream_min_n", thr)
    mgfea.set_option("hstream_keys", 1)
    if kind == "iso":
        grids = [SingleGrid(2, n // 2 ** l) for l in range(L)]
    else:
        grids = [_InterfaceSingleGrid(2, n // 2 ** l, prop=(1, 100), shape=0) for l in range(L)]
    hnet = HNet(3)
    hnet.load_state_dict({f"convLayers.{i}.weight": torch.from_numpy(hw[i]).reshape(1, 1, 3, 3) for i in range(3)})
    R16 = np.repeat((LINEAR_4 / np.float32(4.0)).reshape(1, 9), 16, 0)
    P4 = np.repeat(LINEAR_4.reshape(1, 9), 16, 0)
    eng = VCycleEngine([g.jac for g in grids], B=1, smoother="hjac", hnet=hnet, prolong="table", rtab=R16, r_scale=4.0,
                       ptab=P4, p_scale=1.0)
    eng.set_u(torch.zeros(1, 1, n + 1, n + 1, device="cuda"))
    eng.set_f(grids[0].fnet(torch.ones(1, 1, n + 1, n + 1, device="cuda")))
    ms, runs = time_engine(eng, 20, reset=lambda: eng.u[0].zero_())
    out[f"min_n={thr}"] = round(ms, 4)
    del eng, grids
    torch.cuda.empty_cache()
print(json.dumps(out))
