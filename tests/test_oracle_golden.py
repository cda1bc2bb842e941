"""Pins the oracle (oracle/mgfea_oracle.c + oracle/oracle.py) against golden vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tolerance import check_band

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
OPS = np.load(os.path.join(G, "ops.npz"))
MESH = np.load(os.path.join(G, "mesh.npz"))
ARR = np.load(os.path.join(G, "solve_arrays.npz"))
HIST = json.load(open(os.path.join(G, "solve_histories.json")))
BANDS = json.load(open(os.path.join(G, "bands.json")))  # fp64 runs of the reference (noise bands) + config 3

SIZES = (9, 17, 33)
TAGS = {"iso": (None, None), "c20": ([1, 20], 0), "s100": ([1, 100], 1)}


def close(a, b, rtol=2e-6, name=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    scale = max(np.abs(b).max(), 1e-30)
    err = np.abs(a - b).max() / scale
    assert err <= rtol, f"{name}: rel-to-max err {err:.3e} > {rtol}"


def setup(tag, N):
    prop, shape = TAGS[tag]
    if prop is None:
        return None, O.kernel_table([1.0], 1).reshape(1, 9)
    return O.pattern_keys(N, shape), O.kernel_table(prop, 16).reshape(16, 9)


# ---------------------------------------------------------------- mesh / tables (bit exact)
@pytest.mark.parametrize("n", [4, 8, 16, 32, 64])
@pytest.mark.parametrize("shape", [0, 1])
def test_pattern_keys_bit_exact(n, shape):
    assert np.array_equal(O.pattern_keys(n + 1, shape), MESH[f"keys_n{n}_s{shape}"])


@pytest.mark.parametrize("prop", [[1, 20], [1, 100], [3, 0.5]])
def test_kernel_table_bit_exact(prop):
    ref = MESH[f"ktab_{prop[0]}_{prop[1]}"]
    assert np.array_equal(O.kernel_table(prop, 16).view(np.uint32), ref.view(np.uint32))
    assert np.array_equal(O.kernel_table([1.0], 1)[0].view(np.uint32), MESH["ktab_iso"].view(np.uint32))


# ---------------------------------------------------------------- operators
@pytest.mark.parametrize("N", SIZES)
@pytest.mark.parametrize("tag", list(TAGS))
def test_stiffness_split_residual(tag, N):
    keys, ktab = setup(tag, N)
    u, f = OPS[f"u_{N}"], OPS[f"f_{N}"]
    close(O.stiffness_apply(u, keys, ktab), OPS[f"{tag}_Ku_{N}"][:, 0], name="Ku")
    assert np.array_equal(O.split_x(u, keys, ktab.shape[0]), OPS[f"{tag}_split_{N}"])
    close(O.residual(u, f, keys, ktab), OPS[f"{tag}_res_{N}"][:, 0], name="res")
    # d_mat = per-key diagonal gathered by key (jacobi.py:31-37), bit exact
    d = ktab[:, 4][keys] if keys is not None else np.full((N, N), ktab[0, 4], np.float32)
    assert np.array_equal(d, OPS[f"{tag}_dmat_{N}"][0, 0])


@pytest.mark.parametrize("N", SIZES)
def test_load_vector(N):
    w = O.load_vector_weights(2.0 / (N - 1))
    assert np.array_equal(w.reshape(9), OPS[f"fnet_w_{N}"])
    close(O.conv3x3(OPS[f"u_{N}"], w), OPS[f"fnet_{N}"][:, 0], name="fnet")


@pytest.mark.parametrize("N", SIZES)
@pytest.mark.parametrize("tag", list(TAGS))
def test_jacobi_and_hjacobi(tag, N):
    keys, ktab = setup(tag, N)
    invd = O.inv_diag(2 / 3., ktab[:, 4])
    u, f = OPS[f"u_{N}"], OPS[f"f_{N}"]
    bi, bv = OPS[f"bidx_{N}"], OPS[f"bval_{N}"]
    hw = OPS["hnet_w"]
    close(O.jacobi(u, f, keys, ktab, invd), OPS[f"{tag}_jac1_{N}"][:, 0], name="jac1")
    close(O.jacobi(u, f, keys, ktab, invd, nsweeps=3), OPS[f"{tag}_jac3_{N}"][:, 0], rtol=4e-6, name="jac3")
    close(O.jacobi(u, f, keys, ktab, invd, bi, bv), OPS[f"{tag}_jacbc1_{N}"][:, 0], name="jacbc1")
    close(O.jacobi(u, f, keys, ktab, invd, bi, bv, nsweeps=2), OPS[f"{tag}_jacbc2_{N}"][:, 0], rtol=4e-6, name="jacbc2")
    close(O.hjacobi(u, f, keys, ktab, invd, hw), OPS[f"{tag}_hjac1_{N}"][:, 0], rtol=4e-6, name="hjac1")
    close(O.hjacobi(u, f, keys, ktab, invd, hw, nsweeps=2), OPS[f"{tag}_hjac2_{N}"][:, 0], rtol=8e-6, name="hjac2")
    close(O.hjacobi(u, f, keys, ktab, invd, hw, bi, bv), OPS[f"{tag}_hjacbc1_{N}"][:, 0], rtol=4e-6, name="hjacbc1")
    close(O.hjacobi(u, f, keys, ktab, invd, hw, bi, bv, nsweeps=2), OPS[f"{tag}_hjacbc2_{N}"][:, 0], rtol=8e-6,
          name="hjacbc2")
    # ring of the default-BC result is exactly the boundary value (bit exact masks)
    out = O.jacobi(u, f, keys, ktab, invd)
    assert (out[:, 0, :] == 0).all() and (out[:, -1, :] == 0).all() and (out[:, :, 0] == 0).all()


@pytest.mark.parametrize("N", SIZES)
def test_intergrid_variant_a(N):
    close(O.restrict(OPS[f"f_{N}"], None, O.FW16, 4.0), OPS[f"restrictA_{N}"][:, 0], name="restrictA")
    got = O.prolong_bilinear(OPS[f"vcA_{N}"], OPS[f"u_{N}"])
    # bilinear weights are exact in fp32 -> bit-exact agreement with ATen upsample + add
    assert np.array_equal(got, OPS[f"prolongA_{N}"][:, 0])


@pytest.mark.parametrize("N", SIZES)
@pytest.mark.parametrize("tag", ["c20", "s100"])
def test_intergrid_variant_b_16ch(tag, N):
    prop, shape = TAGS[tag]
    keys, ktab = setup(tag, N)
    r = OPS[f"{tag}_res_{N}"][:, 0]
    close(O.restrict(r, keys, OPS[f"{tag}_Rw_{N}"], 1.7), OPS[f"{tag}_restrictB_{N}"][:, 0], rtol=3e-6, name="restrictB")
    Nc = (N - 1) // 2 + 1
    keys_c = O.pattern_keys(Nc, shape)
    got = O.prolong_table(OPS[f"{tag}_vc_{N}"], OPS[f"u_{N}"], keys_c, OPS[f"{tag}_Pw_{N}"], 0.9)
    close(got, OPS[f"{tag}_prolongB_{N}"][:, 0], rtol=3e-6, name="prolongB")


# ---------------------------------------------------------------- known-answer fixture K u = fnet(rhs)
def test_iso33_known_answer_fixture():
    """Data/IsoPoisson/poisson2d_33x33.h5: dense-matrix Q1 FEM solutions satisfy K u = FNet(rhs) on the interior
    (1.4e-8 in fp64); in fp32 the residual is at rounding level relative to |f|."""
    u, rhs = ARR["iso33_u"], ARR["iso33_rhs"]
    ktab = O.kernel_table([1.0], 1).reshape(1, 9)
    f = O.conv3x3(rhs, O.load_vector_weights(2.0 / 32))
    r = O.residual(u, f, None, ktab)
    rel = np.sqrt(O.sumsq_interior(r)) / np.sqrt(O.sumsq_interior(f))
    assert (rel < 5e-5).all(), rel  # fp32 rounding of K u (|u|~1) relative to |f|~h^2
    assert np.array_equal(u * (1 - ARR["iso33_bidx"]), ARR["iso33_bval"])


# ---------------------------------------------------------------- V-cycle drivers
def model_u0(n, seed=123):
    np.random.seed(seed)
    coef = 100000 + 50000 * np.random.rand(2)
    return (coef[0] * np.random.random((n + 1, n + 1)).astype("f") + coef[1]).astype(np.float32)


def rhs_field(n, seed):
    rs = np.random.RandomState(seed)
    F = rs.standard_normal((1, 1, n + 1, n + 1)).astype(np.float32)
    return O.conv3x3(F, O.load_vector_weights(2.0 / n))


def check_hist(got, ref, ref64=None, rtol=1e-5, name=""):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    tol = np.full(ref.shape, rtol)
    if ref64 is not None:  # fp32-vs-fp64 noise band of the reference itself
        r64 = np.asarray(ref64)
        tol = np.maximum(tol, 10 * np.abs(ref - r64) / r64)
    rel = np.abs(got - ref) / ref
    assert (rel <= tol).all(), f"{name}: rel {rel} tol {tol}"


SMALL = [k for k in HIST if k.startswith("modelA") and HIST[k]["n"] <= 256]


@pytest.mark.parametrize("tag", SMALL)
def test_model_problem_histories(tag):
    h = HIST[tag]
    n = h["n"]
    levels = O.make_levels(n, h["L"])
    cfg = O.CycleCfg(nu1=h["v1v2"][0], nu2=h["v1v2"][1])
    if h["rhs_seed"] is None:
        u0, f = model_u0(n), np.zeros((1, n + 1, n + 1), np.float32)
    else:
        u0, f = np.zeros((n + 1, n + 1), np.float32), rhs_field(n, h["rhs_seed"])
    u, res = O.solve(levels, cfg, u0, f, n_iter=h["n_iter"], EPS=h["EPS"])
    assert len(res) == len(h["res"]), "V-cycle count differs"
    # Tolerance policy (DESIGN.md "parity"): 1e-5 relative per cycle, widened only by the reference's own fp32
    # noise: (a) its fp32-vs-fp64 drift when recorded, (b) ~1e-6 per cycle of accumulated summation-order noise on
    # the levels where ATen uses im2col+MKL sgemm instead of oneDNN (single sample, N<=129), which only shows
    # once the history is > 12 cycles deep or the residual has dropped below 1e-8 relative.
    ref = np.array(h["res"])
    tol = np.where((np.arange(len(ref)) < 12) & (ref / ref[0] >= 1e-8), 1e-5, 5e-5)
    if "res64" in h:
        r64 = np.array(h["res64"])
        tol = np.maximum(tol, 10 * np.abs(ref - r64) / r64)
    rel = np.abs(np.array(res) - ref) / ref
    assert (rel <= tol).all(), f"{tag}: rel {rel} tol {tol}"
    if tag + "_u" in ARR.files and h["rhs_seed"] is None and h["n_iter"] and h["n_iter"] <= 12:
        close(u, ARR[tag + "_u"][:, 0], rtol=2e-5, name=tag + "_u")


def test_model_problem_1024():
    for tag in ("modelA_n1024_L10", "modelA_n1024_L8"):
        h = HIST[tag]
        n = h["n"]
        levels = O.make_levels(n, h["L"])
        _, res = O.solve(levels, O.CycleCfg(), model_u0(n), np.zeros((1, n + 1, n + 1), np.float32), n_iter=6)
        check_hist(res, h["res"][:6], h.get("res64", [None] * 6)[:6] if "res64" in h else None, name=tag)


def test_interface_quirk_history():
    h = HIST["interface_quirk_n64"]
    n = 64
    levels = O.make_levels(n, None, prop=[1, 20], shape=0)
    cfg = O.CycleCfg(quirk_level0=True)
    f = O.conv3x3(np.ones((1, n + 1, n + 1), np.float32), O.load_vector_weights(2.0 / n))
    u, res = O.solve(levels, cfg, np.zeros((n + 1, n + 1), np.float32), f, EPS=5e-5)
    assert len(res) == len(h["res"]) == 14
    # this 1:20 problem amplifies fp32 noise (the reference run here differs from the outputs recorded in the notebook
    # by 4.5e-6 at cycle 1 .. 1.5e-2 at cycle 14): tolerance = the reference's own recorded fp32-vs-fp64 drift
    check_band(res, h["res"], BANDS["interface_quirk_n64"]["res64"], "interface quirk")
    nb = np.array(h["notebook_recorded"])
    assert (np.abs(np.array(res) - nb) / nb <= 2 * np.minimum(2e-5 * 2.0 ** np.arange(len(res)), 2e-2)).all()


@pytest.mark.parametrize("mode", ["jac", "hjac"])
@pytest.mark.parametrize("k", [0, 1, 2])
def test_mgtest_histories(mode, k):
    h = HIST[f"mgtest_{mode}_s{k}"]
    n = 32
    levels = O.make_levels(n)
    levels[0].idx, levels[0].bval = ARR["iso33_bidx"][k][None], ARR["iso33_bval"][k][None]
    cfg = O.CycleCfg(smoother=mode, hw=OPS["hnet_w"], prolong="table", rtab=O.LIN4, r_scale=None, ptab=O.LIN4)
    f = O.conv3x3(ARR["iso33_rhs"][k], O.load_vector_weights(2.0 / n))
    u = O.reset_boundary(np.zeros((1, n + 1, n + 1), np.float32), levels[0].idx, levels[0].bval)
    res = [float(O.residual_norm(levels, u, f)[0])]
    # the notebook loop starts Step from u_mg = zeros (NOT the reset u0): SURVEY App. A.6
    u = np.zeros((1, n + 1, n + 1), np.float32)
    while abs(res[-1]) > 5e-5 and len(res) < 60:
        u = O.vcycle(levels, cfg, u, f)
        res.append(float(O.residual_norm(levels, u, f)[0]))
    assert len(res) == len(h["res"]), (len(res), len(h["res"]))
    check_band(res, h["res"], BANDS[f"mgtest_{mode}_s{k}"]["res64"], f"mgtest {mode} {k}")
    close(u, ARR[f"mgtest_{mode}_s{k}_u"][:, 0], rtol=2e-5)


@pytest.mark.parametrize("tag", ["linear", "learned"])
def test_committed_multigrid_iterate(tag):
    h = HIST[f"iterate_{tag}"]
    n = 32
    levels = O.make_levels(n, None, prop=[1, 20], shape=0)
    if tag == "linear":
        R, P, w = O.FW16, O.LIN4, np.array([4.0, 1.0], np.float32)
    else:
        R, P, w = ARR["learned_R"], ARR["learned_P"], ARR["learned_w"]
    cfg = O.CycleCfg(prolong="table", rtab=R, r_scale=float(w[0]), ptab=P, p_scale=float(w[1]))
    f = O.conv3x3(ARR["iterate_F"], O.load_vector_weights(2.0 / n))
    x = ARR["iterate_x0"][:, 0]
    for it in range(6):
        x = O.vcycle(levels, cfg, x, f)
        got = O.residual_norm(levels, x, f)
        ref = np.array(h["res"][it])
        assert (np.abs(got - ref) / ref < 3e-5 * 4 ** it).all(), (it, got, ref)
    close(x, ARR[f"iterate_{tag}_u"][:, 0], rtol=1e-4)


@pytest.mark.parametrize("mode", ["jac", "hjac"])
@pytest.mark.parametrize("n", [64, 256])
def test_config3_histories(mode, n):
    """BASELINE config 3 (two-phase circle 1:100, keys + 16-channel linear R/P, w = [4, 1], Jacobi / learned HNet
    smoother) against the UNMODIFIED FEANet/multigrid.py MultiGrid.iterate fed with a closed-form mesh
    (tests/golden/make_golden.py gen_bands).  Faithful includes the reference's failure: at n = 256 the Jacobi cycle
    STALLS near 3 r0 (rediscretised coarse operators + linear R/P across a 1:100 interface), and so must we."""
    b = BANDS[f"cfg3_{mode}_n{n}"]
    levels = O.make_levels(n, None, prop=b["prop"], shape=0)
    R, P = np.repeat(O.FW16, 16, 0), np.repeat(O.LIN4, 16, 0)
    cfg = O.CycleCfg(smoother=mode, hw=OPS["hnet_w"], prolong="table", rtab=R, r_scale=b["w"][0], ptab=P,
                     p_scale=b["w"][1])
    f = O.conv3x3(np.ones((1, n + 1, n + 1), np.float32), O.load_vector_weights(2.0 / n))
    u = np.zeros((1, n + 1, n + 1), np.float32)
    assert abs(float(O.residual_norm(levels, u, f)[0]) - b["r0"]) <= 1e-6 * b["r0"]
    res = []
    for _ in range(len(b["res"])):
        u = O.vcycle(levels, cfg, u, f)
        res.append(float(O.residual_norm(levels, u, f)[0]))
    check_band(res, b["res"], b["res64"], f"config 3 {mode} n={n}")
    if n == 256 and mode == "jac":
        assert min(res[3:]) > 2.0 * b["r0"], "the reference stalls here; a converging history is NOT parity"


def test_feanet_torch_matches_reference_histories():
    """the ATen-call-for-call restatement used as cpu_baseline reproduces the reference's golden residuals exactly
    (same op sequence on the same torch build => bit-identical; allow 1e-6 for a different build)"""
    import torch

    from oracle import feanet_torch as FT

    for tag in ("modelA_n64_v11", "modelA_n64_L4_v11", "modelA_n256_v11"):
        h = HIST[tag]
        n = h["n"]
        lv = FT.make_levels(n, h["L"])
        u0 = torch.from_numpy(model_u0(n)).reshape(1, 1, n + 1, n + 1)
        _, res = FT.solve(lv, u0, torch.zeros(1, 1, n + 1, n + 1), n_iter=8)
        ref = np.array(h["res"][:8])
        assert (np.abs(np.array(res) - ref) / ref < 1e-6).all(), (tag, res, ref)
    # two-phase: Level.K with 16 channels equals the oracle's key-indexed stencil bit for bit
    N = 33
    keys, ktab = setup("c20", N)
    lvl = FT.Level(N, ktab, keys)
    u = torch.from_numpy(OPS[f"u_{N}"])
    assert np.array_equal(lvl.K(u).numpy()[:, 0], O.stiffness_apply(OPS[f"u_{N}"], keys, ktab))


# ---------------------------------------------------------------- per-element conductivity
TP = np.load(os.path.join(G, "testpoisson.npz"))


@pytest.mark.parametrize("shape,prop", [(0, [1, 20]), (1, [1, 100]), (0, [3, 0.5])])
@pytest.mark.parametrize("N", [9, 17, 33, 65])
def test_element_operator_equals_pattern_operator(shape, prop, N):
    """the per-element operator fed with the two-phase conductivity map reproduces the reference's 16-pattern operator
    (pinned above against the reference) bit for bit: K u on the interior, Jacobi sweeps everywhere"""
    keys, tab = O.pattern_keys(N, shape), O.kernel_table(prop, 16).reshape(16, 9)
    a = np.array(prop, np.float32)[O.phase_map(N, shape)]
    # the phase map is consistent with all four pattern positions of the reference's keys, not only e4
    pat = np.array([O.REF_PATTERNS[k] for k in range(16)])[keys.astype(np.int64)]  # (N, N, 4): e1..e4
    ph = np.zeros((N + 1, N + 1), np.int64)
    ph[1:N, 1:N] = O.phase_map(N, shape)
    i, j = np.meshgrid(np.arange(1, N - 1), np.arange(1, N - 1), indexing="ij")
    assert np.array_equal(pat[i, j, 0], ph[i, j + 1]) and np.array_equal(pat[i, j, 1], ph[i, j])
    assert np.array_equal(pat[i, j, 2], ph[i + 1, j]) and np.array_equal(pat[i, j, 3], ph[i + 1, j + 1])
    rs = np.random.RandomState(N)
    u = rs.standard_normal((2, N, N)).astype(np.float32)
    f = rs.standard_normal((2, N, N)).astype(np.float32)
    k1, k2 = O.stiffness_apply(u, keys, tab), O.elem_stiffness_apply(u, a)
    assert np.array_equal(k1[:, 1:-1, 1:-1], k2[:, 1:-1, 1:-1])
    j1 = O.jacobi(u, f, keys, tab, O.inv_diag(2 / 3., tab[:, 4]), nsweeps=3)
    assert np.array_equal(j1, O.elem_jacobi(u, f, a, nsweeps=3))
    d = O.elem_diag(a)
    assert np.array_equal(d[1:-1, 1:-1], tab[:, 4][keys][1:-1, 1:-1])


def test_element_operator_testpoisson_fixture():
    """Data/TestPoisson/poisson2d_33x33.h5: `material` is one value per element (32 x 32); with it the per-element
    operator satisfies K solution = fnet(source) on the interior like the reference's own modules do in fp64
    (8.8e-10, recorded in the fixture), here in fp32"""
    N = 33
    for k in range(3):
        a = TP["material"][k].astype(np.float32)
        assert a.shape == (N - 1, N - 1)
        u, src = TP["solution"][k].astype(np.float32), TP["source"][k].astype(np.float32)
        f = O.conv3x3(src, O.load_vector_weights(2.0 / (N - 1)))
        r = O.elem_residual(u, f, a)[0, 1:-1, 1:-1]
        assert np.abs(r).max() <= 2e-5 * np.abs(O.elem_stiffness_apply(u, a)).max()
        iso = O.stiffness_apply(u, None, O.kernel_table([1.0], 1).reshape(1, 9))
        if (a == 1).all():  # this file's material is homogeneous: the element operator IS the MeshSquare operator
            assert np.array_equal(O.elem_stiffness_apply(u, a)[:, 1:-1, 1:-1], iso[:, 1:-1, 1:-1])
    assert float(TP["ref_residual_interior_max"][0]) < 1e-8


def test_element_vcycle_converges_and_coarsening():
    n = 64
    rs = np.random.RandomState(4)
    a = np.exp(rs.uniform(np.log(0.2), np.log(5.0), (n, n))).astype(np.float32)
    levels = [a]
    for _ in range(int(np.log2(n)) - 1):
        levels.append(O.coarsen_elements(levels[-1]))
    assert levels[-1].shape == (2, 2) and np.allclose(levels[1][0, 0], a[:2, :2].mean(), rtol=1e-6)
    f = 0.01 * rs.standard_normal((1, n + 1, n + 1)).astype(np.float32)
    u = np.zeros((1, n + 1, n + 1), np.float32)
    res = [float(np.sqrt(O.sumsq_interior(O.elem_residual(u, f, a))[0]))]
    for _ in range(8):
        u = O.elem_vcycle(levels, u, f)
        res.append(float(np.sqrt(O.sumsq_interior(O.elem_residual(u, f, a))[0])))
    assert res[-1] < 1e-2 * res[0] and all(res[i + 1] < res[i] for i in range(8))  # q ~ 0.46 at 25x random contrast


# ---------------------------------------------------------------- periodic-BC smoother (FEANet/jacobi.py:50-97)
PBC = np.load(os.path.join(G, "pbc.npz"))


@pytest.mark.parametrize("n", [8, 16, 32])
def test_jacobi_pbc_bit_exact(n):
    w = O.kernel_table([1.0], 1).reshape(9)
    invd = float(O.inv_diag(2 / 3., np.array([w[4]], np.float32))[0])
    for k, tag in ((1, "jac1"), (3, "jac3")):
        got = O.jacobi_pbc(PBC[f"u_{n}"], PBC[f"fpad_{n}"], w, invd, k)
        assert np.array_equal(got, PBC[f"{tag}_{n}"][:, 0]), (n, k)
