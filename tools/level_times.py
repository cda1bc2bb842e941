"""Live (warm, no profiler) time of every launch of the iso V(1,1) cycle, level by level: CUDA events around repeated
launches of the same fused leg through the C ABI.  usage: level_times.py [n] [B]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multigrid-feanet_b200")]
import numpy as np
import torch

import mgfea
from FEANet.drivers import Multigrid

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
np.random.seed(123)
prob = Multigrid(n, batch=B) if B > 1 else Multigrid(n)
eng = prob._engine(1, 1, 0, B=B)
eng.refresh()
rt = eng._keep[0]
L = eng.L


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


tot = 0.0
for l in range(L - 1):
    N = eng.u[l].N
    if N <= 65:
        break
    g0, g1 = eng._grids[l], eng._grids[l + 1]

    def down(l=l, g0=g0):
        mgfea.check(mgfea.lib().mgfea_smooth_residual_restrict(
            ctypes.byref(g0), eng.u[l].ptr if l == 0 else None, eng.u_alt[l].ptr, eng.f[l].ptr, 1, 0, None, 0,
            eng.f[l + 1].ptr, eng.f[l + 1].pitch, eng.f[l + 1].plane, rt.data_ptr(), 1, 1, 4.0, None, B, mgfea.stream_ptr()))

    def up(l=l, g0=g0, g1=g1):
        mgfea.check(mgfea.lib().mgfea_prolong_correct_smooth(
            ctypes.byref(g0), ctypes.byref(g1), eng.u[l + 1].ptr, eng.u_alt[l].ptr, eng.u[l].ptr, eng.f[l].ptr,
            mgfea.PROLONG_BILINEAR, None, 0, 0, 0.0, None, 1, 0, None, 0, B, mgfea.stream_ptr()))

    td, tu = timeit(down), timeit(up)
    tot += td + tu
    print(f"level {l} N={N}: down {td:7.2f} us  up {tu:7.2f} us")
    lt = l + 1
# the tail: V-cycle on the sub-hierarchy starting at the last streamed level minus its two legs
sub_g = (mgfea.Grid * (L - lt + 1))(*[eng._grids[i] for i in range(lt - 1, L)])
sub_b = (mgfea.LevelBufs * (L - lt + 1))(*[eng._bufs[i] for i in range(lt - 1, L)])
cfg = eng._cfg


def sub():
    mgfea.check(mgfea.lib().mgfea_vcycle(sub_g, sub_b, L - lt + 1, ctypes.byref(cfg), eng.sumsq.data_ptr(), None, None, B,
                                         mgfea.stream_ptr()))


ts = timeit(sub)
print(f"sub-cycle from level {lt - 1} (2 legs + tail): {ts:7.2f} us -> tail ~ {ts - td - tu:7.2f} us")
print(f"sum of legs {tot:7.2f} us + tail {ts - td - tu:7.2f} = {tot + ts - td - tu:7.2f} us")
eng._ctl_reset(0, -1.0, eng.max_cycles)
eng._ensure_graph()
z = eng.ctl.clone()
print(f"graph cycle: {timeit(lambda: (eng.ctl.copy_(z, non_blocking=True), eng._graph.replay()), 100):7.2f} us")
