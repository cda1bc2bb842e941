"""keyed streaming kernels vs the oracle, one fused leg at a time (debug aid).  usage: keyed_check.py [n]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multigrid-feanet_b200")]
import numpy as np, torch
import mgfea
from FEANet.drivers import _InterfaceSingleGrid
from FEANet.solver import VCycleEngine, FULL_WEIGHTING_16
from oracle import oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
prop = tuple(float(x) for x in sys.argv[2].split(",")) if len(sys.argv) > 2 else (1, 20)
N = n + 1
grids = [_InterfaceSingleGrid(2, n // 2 ** l, prop=prop, shape=0) for l in range(2)]
eng = VCycleEngine([g.jac for g in grids], B=1, smoother="jac")
rs = np.random.RandomState(3)
u0 = rs.standard_normal((1, N, N)).astype(np.float32)
f = 0.01 * rs.standard_normal((1, N, N)).astype(np.float32)
eng.set_u(torch.from_numpy(u0)); eng.set_f(torch.from_numpy(f))
lv = O.make_levels(n, 2, prop=list(prop), shape=0)
g0, g1 = eng._grids[0], eng._grids[1]
rt = eng._keep[0]


def report(name, got, ref):
    bad = got != ref
    print(name, "mismatches", int(bad.sum()), "of", bad.size)
    if bad.any():
        ys, xs = np.nonzero(bad)
        print("   rows", ys.min(), ys.max(), "cols", xs.min(), xs.max(), "max abs", np.abs(got - ref).max())
        print("   per 120-col strip", np.bincount(xs // 120)[:8], " per 32-row band", np.bincount(ys // 32)[:20])
        print("   first", ys[0], xs[0], got[ys[0], xs[0]], ref[ys[0], xs[0]], "key", lv[0].keys[ys[0], xs[0]])
        if got.shape[0] == N:
            print("   by key", np.bincount(lv[0].keys[ys, xs], minlength=16), " row%6", np.bincount(ys % 6, minlength=6),
                  " col%4", np.bincount(xs % 4, minlength=4))


for zero in (False, True):
    mgfea.check(mgfea.lib().mgfea_smooth_residual_restrict(
        ctypes.byref(g0), None if zero else eng.u[0].ptr, eng.u_alt[0].ptr, eng.f[0].ptr, 1, 0, None, 0, eng.f[1].ptr,
        eng.f[1].pitch, eng.f[1].plane, rt.data_ptr(), 1, 1, 4.0, None, 1, mgfea.stream_ptr()))
    uin = np.zeros_like(u0) if zero else u0
    u1 = O.jacobi(uin, f, lv[0].keys, lv[0].ktab, lv[0].invd)
    fc = O.restrict(O.residual(u1, f, lv[0].keys, lv[0].ktab), None, O.FW16, 4.0)
    report(f"down zero={zero} u1", eng.u_alt[0].view.cpu().numpy()[0, 0], u1[0])
    report(f"down zero={zero} fc", eng.f[1].view.cpu().numpy()[0, 0], fc[0])
vc = rs.standard_normal((1, n // 2 + 1, n // 2 + 1)).astype(np.float32)
vc[:, 0, :] = vc[:, -1, :] = 0; vc[:, :, 0] = vc[:, :, -1] = 0
eng.u[1].view.copy_(torch.from_numpy(vc)[:, None])
eng.u_alt[0].view.copy_(torch.from_numpy(u0)[:, None])
ss = torch.zeros(1, dtype=torch.float64, device="cuda")
mgfea.check(mgfea.lib().mgfea_prolong_correct_smooth_norm(
    ctypes.byref(g0), ctypes.byref(g1), eng.u[1].ptr, eng.u_alt[0].ptr, eng.u[0].ptr, eng.f[0].ptr,
    mgfea.PROLONG_BILINEAR, None, 0, 0, 0.0, None, 1, 0, None, 0, ss.data_ptr(), 1, mgfea.stream_ptr()))
uc = O.prolong_bilinear(vc, u0)
u2 = O.jacobi(uc, f, lv[0].keys, lv[0].ktab, lv[0].invd)
report("up u2", eng.u[0].view.cpu().numpy()[0, 0], u2[0])
print("norm", float(ss.item()), O.sumsq_interior(O.residual(u2, f, lv[0].keys, lv[0].ktab)))

# ---- which rounding does the kernel use at a mismatching inclusion node?  (Jacobi update from u0; this is how the
# FFMA2 contraction of the packed update was found: the kernel matched `fused`, the oracle `unfused`)
eng.set_u(torch.from_numpy(u0))
mgfea.check(mgfea.lib().mgfea_smooth_residual_restrict(
    ctypes.byref(g0), eng.u[0].ptr, eng.u_alt[0].ptr, eng.f[0].ptr, 1, 0, None, 0, eng.f[1].ptr,
    eng.f[1].pitch, eng.f[1].plane, rt.data_ptr(), 1, 1, 4.0, None, 1, mgfea.stream_ptr()))
got = eng.u_alt[0].view.cpu().numpy()[0, 0]
ref = O.jacobi(u0, f, lv[0].keys, lv[0].ktab, lv[0].invd)[0]
ys, xs = np.nonzero((got != ref) & (lv[0].keys == 1))
f32 = np.float32
tab = np.asarray(lv[0].ktab, np.float32).reshape(-1, 9)
invd = np.asarray(lv[0].invd, np.float32)
for y, x in list(zip(ys, xs))[:4]:
    s = None
    for a in range(3):
        for c in range(3):
            w, v = tab[lv[0].keys[y - 1 + a, x - 1 + c], 3 * a + c], u0[0, y - 1 + a, x - 1 + c]
            s = f32(np.float64(w) * np.float64(v)) if s is None else f32(np.float64(w) * np.float64(v) + np.float64(s))
    inv = invd[1]
    d = f32(f[0, y, x] - s)
    unf = f32(f32(inv * d) + u0[0, y, x])
    fus = f32(np.float64(inv) * np.float64(d) + np.float64(u0[0, y, x]))
    print(f"node ({y},{x}): kernel {got[y, x]!r} oracle {ref[y, x]!r} unfused {unf!r} fused {fus!r}  Ku {s!r} inv {inv!r}")
