"""level-0 legs of the learned-smoother cycle in isolation (us per launch, CUDA events, 30 reps): HNet down leg
(sweep + residual + restriction) and up leg (table prolongation + sweep + norm).  argv: iso|keys [n] [hstream_min_n]
[down|up|both] [strips per resident warp]"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multigrid-feanet_b200")]
import numpy as np
import torch

import mgfea
from FEANet.drivers import HNet, SingleGrid, _InterfaceSingleGrid
from FEANet.solver import LINEAR_4, VCycleEngine

kind = sys.argv[1] if len(sys.argv) > 1 else "iso"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
thr = int(sys.argv[3]) if len(sys.argv) > 3 else 129
which = sys.argv[4] if len(sys.argv) > 4 else "both"
if len(sys.argv) > 5:
    mgfea.set_option("hstream_over", int(sys.argv[5]))
mgfea.set_option("hstream_min_n", thr)
mgfea.set_option("hstream_keys", 1)
L = int(np.log2(n))
hw = np.load(os.path.join(ROOT, "tests", "golden", "ops.npz"))["hnet_w"]
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
eng.set_u(torch.randn(1, 1, n + 1, n + 1, device="cuda"))
eng.set_f(grids[0].fnet(torch.ones(1, 1, n + 1, n + 1, device="cuda")))
eng.refresh()
g0, g1, cfg = eng._grids[0], eng._grids[1], eng._cfg
lib, st = mgfea.lib(), mgfea.stream_ptr
ntab = 1 if kind == "iso" else 16  # single-pattern levels use table 0 (as mgfea_vcycle does)


def tk(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def down():
    mgfea.check(lib.mgfea_smooth_residual_restrict(
        ctypes.byref(g0), eng.u[0].ptr, eng.u_alt[0].ptr, eng.f[0].ptr, 1, mgfea.SMOOTH_HJACOBI, cfg.hw, 3, eng.f[1].ptr,
        eng.f[1].pitch, eng.f[1].plane, cfg.rtab, ntab, 1, 4.0, None, 1, st()))


def up_norm():
    mgfea.check(lib.mgfea_prolong_correct_smooth_norm(
        ctypes.byref(g0), ctypes.byref(g1), eng.u[1].ptr, eng.u_alt[0].ptr, eng.u[0].ptr, eng.f[0].ptr,
        mgfea.PROLONG_TABLE, cfg.ptab, ntab, 1, 1.0, None, 1, mgfea.SMOOTH_HJACOBI, cfg.hw, 3, eng.sumsq.data_ptr(), 1,
        st()))


out = {"kind": kind, "n": n, "hstream_min_n": thr, "over": sys.argv[5] if len(sys.argv) > 5 else "default"}
if which in ("down", "both"):
    out["down_us"] = round(tk(down), 1)
if which in ("up", "both"):
    out["up_norm_us"] = round(tk(up_norm), 1)
print(json.dumps(out))
