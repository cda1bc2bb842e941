"""H2D / D2H rates from pinned host memory for the shapes Multigrid.Solve moves (contiguous vs strided-into-padded)."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multigrid-feanet_b200")]
import torch
N = 4097
pitch = (N + 31) // 32 * 32
h = torch.randn(1, 1, N, N).pin_memory()
d_cont = torch.empty(1, 1, N, N, device="cuda")
d_pad = torch.zeros(1, N, pitch, device="cuda")
hp = torch.empty(1, 1, N, N).pin_memory()


def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


mb = N * N * 4 / 1e6
print("H2D contiguous        %.2f ms  %.1f GB/s" % (1e3 * t(lambda: d_cont.copy_(h, non_blocking=True)), mb / 1e3 / t(lambda: d_cont.copy_(h, non_blocking=True))))
v = d_pad[:, None, :, :N]
print("H2D strided (padded)  %.2f ms  %.1f GB/s" % (1e3 * t(lambda: v.copy_(h, non_blocking=True)), mb / 1e3 / t(lambda: v.copy_(h, non_blocking=True))))
print("D2H contiguous        %.2f ms  %.1f GB/s" % (1e3 * t(lambda: hp.copy_(d_cont, non_blocking=True)), mb / 1e3 / t(lambda: hp.copy_(d_cont, non_blocking=True))))
print("D2H strided (padded)  %.2f ms  %.1f GB/s" % (1e3 * t(lambda: hp.copy_(v, non_blocking=True)), mb / 1e3 / t(lambda: hp.copy_(v, non_blocking=True))))
s2 = torch.cuda.Stream()
h2 = torch.randn(1, 1, N, N).pin_memory()
d2 = torch.empty(1, 1, N, N, device="cuda")


def both():
    d_cont.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        hp.copy_(d2, non_blocking=True)


print("H2D + D2H concurrent  %.2f ms (2 x %.0f MB)" % (1e3 * t(both), mb))
