"""A/B timing of one build: 4097^2 cycle (graph replay, median of 5 x 100) + the level-0 legs in isolation."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multigrid-feanet_b200")]
import numpy as np
import torch

import mgfea
from bench import model_u0, time_engine
from FEANet.drivers import Multigrid

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for kv in sys.argv[2:]:  # name=value kernel-selection options (mgfea_set_option)
    k, v = kv.split("=")
    mgfea.set_option(k, int(v))
N = n + 1
prob = Multigrid(n)
eng = prob._engine(1, 1, 0, B=1)
eng.set_u(torch.from_numpy(model_u0(n)).reshape(1, 1, N, N))
eng.set_f(torch.zeros(1, 1, N, N))
ms, runs = time_engine(eng, 100)
g0, g1, rt = eng._grids[0], eng._grids[1], eng._keep[0]


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
    mgfea.check(mgfea.lib().mgfea_smooth_residual_restrict(
        ctypes.byref(g0), eng.u[0].ptr, eng.u_alt[0].ptr, eng.f[0].ptr, 1, 0, None, 0, eng.f[1].ptr,
        eng.f[1].pitch, eng.f[1].plane, rt.data_ptr(), 1, 1, 4.0, None, 1, mgfea.stream_ptr()))


def up_norm():
    mgfea.check(mgfea.lib().mgfea_prolong_correct_smooth_norm(
        ctypes.byref(g0), ctypes.byref(g1), eng.u[1].ptr, eng.u_alt[0].ptr, eng.u[0].ptr, eng.f[0].ptr,
        mgfea.PROLONG_BILINEAR, None, 0, 0, 0.0, None, 1, 0, None, 0, eng.sumsq.data_ptr(), 1, mgfea.stream_ptr()))


print(json.dumps({"lib": os.path.basename(mgfea.LIB_PATH), "n": n, "options": sys.argv[2:], "ms_per_cycle": ms, "runs": runs,
                  "down_us": tk(down), "up_norm_us": tk(up_norm)}))
