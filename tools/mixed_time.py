"""time of one fp64 defect-correction step (graph replay) and of its parts at n^2.  usage: mixed_time.py [n]"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multigrid-feanet_b200")]
import numpy as np, torch
import mgfea
from FEANet.drivers import Multigrid
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = n + 1
np.random.seed(123)
prob = Multigrid(n)
g = torch.Generator(device="cuda").manual_seed(0)
prob.grids[0].f = prob.grids[0].fnet(torch.randn(1, 1, N, N, generator=g, device="cuda"))
prob.initial_v = torch.zeros(1, 1, N, N, device="cuda")
prob.SolveMixed([1, 1], n_iter=2)
eng = prob._mixed_engine


def timeit(fn, reps=20):
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


z = eng.ctl.clone()
z[0] = 0
z[1] = 0
g0 = ctypes.byref(eng._grids[0])
print("graph step      %8.1f us" % timeit(lambda: (eng.ctl.copy_(z, non_blocking=True), eng._graph64.replay())))
print("fp32 cycle      %8.1f us" % timeit(lambda: eng.cycle()))
print("correct_f64     %8.1f us" % timeit(lambda: mgfea.check(mgfea.lib().mgfea_correct_f64(g0, eng.u64.data_ptr(), eng.u[0].ptr, None, 1, mgfea.stream_ptr()))))
print("defect_f64      %8.1f us" % timeit(lambda: mgfea.check(mgfea.lib().mgfea_defect_f64(g0, eng.u64.data_ptr(), eng.f64.data_ptr(), eng.f[0].ptr, eng.sumsq.data_ptr(), None, None, 1, mgfea.stream_ptr()))))
torch.cuda.synchronize()
t0 = time.perf_counter()
eng._load64(eng.u64, prob.initial_v, True)
eng._load64(eng.f64, prob.grids[0].f, False)
torch.cuda.synchronize()
print("load64 u,f      %8.1f us" % ((time.perf_counter() - t0) * 1e6))
t0 = time.perf_counter()
r = prob.SolveMixed([1, 1], n_iter=13)
torch.cuda.synchronize()
print("SolveMixed(13)  %8.1f us" % ((time.perf_counter() - t0) * 1e6))
r0 = float(torch.sqrt(eng.r0_sumsq.sum()).item())
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = prob.SolveMixed([1, 1], EPS=1e-8 * r0, chunk=4)
    torch.cuda.synchronize()
    print("SolveMixed(EPS) %8.1f us, %d cycles" % ((time.perf_counter() - t0) * 1e6, len(r)))
