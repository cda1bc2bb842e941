#!/usr/bin/env python
"""Generate the committed golden vectors by running the UNMODIFIED reference in this container.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz, *.json

Needs /root/reference (read-only).  The outputs are small fixtures that travel to the GPU
box; the tests never read /root/reference themselves.  Inputs that are too large to commit
(>= 1025^2) are regenerated from the seeds recorded in the fixture (numpy MT19937 is
platform independent) and only the residual histories are stored.
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402

torch.set_num_threads(os.cpu_count())
OUT = HERE


def t2n(t):
    return t.detach().cpu().numpy().copy()


def keys_of(mesh):
    n = mesh.nnode_edge
    k = np.zeros(n * n, np.int64)
    tot = np.zeros(n * n, np.int64)
    for key, m in mesh.global_pattern_center.items():
        k += key * np.asarray(m)
        tot += np.asarray(m)
    assert (tot == 1).all()
    return k.reshape(n, n).astype(np.uint8)


def gen_mesh():
    H.load_reference()
    from FEANet.mesh import MeshCenterInterface, MeshSquare

    out = {}
    for n in (4, 8, 16, 32, 64):
        for shape in (0, 1):
            m = MeshCenterInterface(2, [1, 20], n + 1, shape=shape)
            out[f"keys_n{n}_s{shape}"] = keys_of(m)
    for prop in ([1, 20], [1, 100], [3, 0.5]):
        m = MeshCenterInterface(2, prop, 9)
        out[f"ktab_{prop[0]}_{prop[1]}"] = np.stack([m.kernel_dict[k] for k in range(16)])
    out["ktab_iso"] = MeshSquare(2, 9).kernel_dict[0]
    np.savez_compressed(os.path.join(OUT, "mesh.npz"), **out)
    print("mesh.npz", len(out))


def gen_ops():
    H.load_reference()
    from FEANet.geo import Geometry
    from FEANet.jacobi import JacobiBlock
    from FEANet.mesh import MeshCenterInterface, MeshSquare
    from FEANet.model import FNet, KNet

    ns = H.notebook_namespace("M-FEANet-mg_test.ipynb", [1, 2, 3, 4, 5])
    HNet = ns["HNet"]
    hnet = HNet(3)
    hnet.load_state_dict(torch.load(os.path.join(H.REF, "Model/learn_iterator/iso_poisson/iso_poisson_33x33.pth"),
                                    weights_only=True))
    hw = np.stack([t2n(l.weight).reshape(9) for l in hnet.convLayers])
    M = H.load_reference_multigrid_module()
    nsA = H.notebook_namespace("MM_Model_convergence.ipynb", [1, 2, 3, 4])

    g = torch.Generator().manual_seed(20260)
    out = {"hnet_w": hw}
    for N in (9, 17, 33):
        B = 2
        u = torch.randn(B, 1, N, N, generator=g)
        f = torch.randn(B, 1, N, N, generator=g)
        out[f"u_{N}"], out[f"f_{N}"] = t2n(u), t2n(f)
        # general (data) Dirichlet BC: ring values random, mask = interior ones (per-sample)
        geo = Geometry(N)
        bidx = geo.geometry_idx.repeat(B, 1, 1, 1).clone()
        bval = torch.randn(B, 1, N, N, generator=g) * (1 - bidx)
        out[f"bidx_{N}"], out[f"bval_{N}"] = t2n(bidx), t2n(bval)
        for tag, mesh in (("iso", MeshSquare(2, N)), ("c20", MeshCenterInterface(2, [1, 20], N, shape=0)),
                          ("s100", MeshCenterInterface(2, [1, 100], N, shape=1))):
            knet = KNet(mesh)
            fnet = FNet(2.0 / (N - 1))
            jac = JacobiBlock(knet, mesh, 2 / 3., geo.geometry_idx, geo.boundary_value)
            with torch.no_grad():
                out[f"{tag}_Ku_{N}"] = t2n(knet(u))
                out[f"{tag}_split_{N}"] = t2n(knet.split_x(u))
                out[f"{tag}_dmat_{N}"] = t2n(jac.d_mat)
                out[f"{tag}_jac1_{N}"] = t2n(jac.jacobi_convolution(u, f))
                v = u
                for _ in range(3):
                    v = jac.jacobi_convolution(v, f)
                out[f"{tag}_jac3_{N}"] = t2n(v)
                jac2 = JacobiBlock(knet, mesh, 2 / 3., bidx, bval)
                out[f"{tag}_dmatB_{N}"] = t2n(jac2.d_mat)
                out[f"{tag}_jacbc1_{N}"] = t2n(jac2.jacobi_convolution(u, f))
                v = u
                for _ in range(2):
                    v = jac2.jacobi_convolution(v, f)
                out[f"{tag}_jacbc2_{N}"] = t2n(v)
                if tag == "iso":
                    out[f"fnet_{N}"] = t2n(fnet(u))
                    out[f"fnet_w_{N}"] = t2n(fnet.net.weight).reshape(9)

                # learned smoother HRelax (mg_test cell 5) on top of each jac (default and data BC)
                class _G:  # duck-typed grid holder for HJacIterator
                    pass

                for bt, jj in (("", jac), ("bc", jac2)):
                    gh = _G()
                    gh.jac = jj
                    it = ns["HJacIterator"](n=N - 1, hnet=hnet, grid=gh)
                    out[f"{tag}_hjac{bt}1_{N}"] = t2n(it.HRelax(u, f, 1))
                    out[f"{tag}_hjac{bt}2_{N}"] = t2n(it.HRelax(u, f, 2))
                # residual
                r = f - knet(u)
                out[f"{tag}_res_{N}"] = t2n(r)
                # 16-channel / 1-channel R and P of FEANet/multigrid.py with per-channel perturbed weights
                C = knet.n_channel
                if C == 16:
                    P4 = torch.tensor([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=torch.float32) / 4.0
                    R16 = torch.tensor([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=torch.float32) / 16.0
                    conv = M.RestrictionNet(R16)
                    deconv = M.ProlongationNet(P4)
                    gp = torch.Generator().manual_seed(7 + N)
                    conv.net.weight.data += 0.05 * torch.randn(conv.net.weight.shape, generator=gp)
                    deconv.net.weight.data += 0.05 * torch.randn(deconv.net.weight.shape, generator=gp)
                    out[f"{tag}_Rw_{N}"] = t2n(conv.net.weight).reshape(16, 9)
                    out[f"{tag}_Pw_{N}"] = t2n(deconv.net.weight).reshape(16, 9)
                    rF = knet.split_x(r)
                    rFC = conv(rF[:, :, 1:-1, 1:-1].clone())
                    rFC = torch.nn.functional.pad(rFC, (1, 1, 1, 1), "constant", 0)
                    out[f"{tag}_restrictB_{N}"] = t2n(1.7 * rFC)
                    # prolongation uses the COARSE level's knet for the split
                    Nc = (N - 1) // 2 + 1
                    shape = 0 if tag == "c20" else 1
                    prop = [1, 20] if tag == "c20" else [1, 100]
                    mesh_c = MeshCenterInterface(2, prop, Nc, shape=shape)
                    knet_c = KNet(mesh_c)
                    vc = torch.randn(B, 1, Nc, Nc, generator=g)
                    out[f"{tag}_vc_{N}"] = t2n(vc)
                    eF = deconv(knet_c.split_x(vc).clone())
                    out[f"{tag}_prolongB_{N}"] = t2n(u + 0.9 * eF)
        # variant A restrict / interpolate (iso notebook driver); needs a Multigrid instance of size N-1
        np.random.seed(1)
        with contextlib.redirect_stdout(io.StringIO()):
            mgA = nsA["Multigrid"](N - 1)
        with torch.no_grad():
            out[f"restrictA_{N}"] = t2n(4 * mgA.Restrict(f))
            Nc = (N - 1) // 2 + 1
            vc = torch.randn(B, 1, Nc, Nc, generator=g)
            out[f"vcA_{N}"] = t2n(vc)
            out[f"prolongA_{N}"] = t2n(u + mgA.Interpolate(vc))
    np.savez_compressed(os.path.join(OUT, "ops.npz"), **out)
    print("ops.npz", len(out))


def model_problem_u0(n, seed=123):
    """MM_Model_convergence.ipynb cell 3 `random_data` with the (commented-out) seed enabled, cast to fp32."""
    np.random.seed(seed)
    coef = 100000 + 50000 * np.random.rand(2)
    return (coef[0] * np.random.random((n + 1, n + 1)).astype("f") + coef[1]).astype(np.float32)


def gen_solve():
    nsA = H.notebook_namespace("MM_Model_convergence.ipynb", [1, 2, 3, 4])
    MG = nsA["Multigrid"]
    hist = {}
    arrays = {}

    def run(tag, n, L, v1v2, rec=True, n_iter=None, EPS=None, rhs_seed=None, keep_u=False):
        t0 = time.time()
        np.random.seed(123)
        with contextlib.redirect_stdout(io.StringIO()):
            p = MG(n, L)
        p.initial_v = torch.from_numpy(model_problem_u0(n))
        if rhs_seed is not None:
            rs = np.random.RandomState(rhs_seed)
            F = torch.from_numpy(rs.standard_normal((1, 1, n + 1, n + 1)).astype(np.float32))
            with torch.no_grad():
                p.grids[0].f = p.grids[0].fnet(F)
            p.initial_v = torch.zeros(n + 1, n + 1)
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            res = p.Solve(list(v1v2), rec=rec, n_iter=n_iter, EPS=EPS)
        hist[tag] = dict(n=n, L=L, v1v2=list(v1v2), rec=rec, n_iter=n_iter, EPS=EPS, rhs_seed=rhs_seed,
                         res=[float(r) for r in res])
        if keep_u:
            arrays[tag + "_u"] = t2n(p.grids[0].v)
        print(tag, len(res), "cycles", "%.1fs" % (time.time() - t0), res[:2], res[-1])

    for n in (2, 4, 8, 16, 32):
        run(f"modelA_n{n}_v11", n, None, (1, 1), n_iter=12, keep_u=(n <= 16))
    run("modelA_n64_v11", 64, None, (1, 1), n_iter=20, keep_u=True)
    run("modelA_n64_v11_iter", 64, None, (1, 1), rec=False, n_iter=8, keep_u=True)
    for v in ((0, 1), (1, 0), (0, 2), (2, 0), (1, 2), (2, 1), (2, 2), (3, 3)):
        run(f"modelA_n64_v{v[0]}{v[1]}", 64, None, v, n_iter=10)
    run("modelA_n64_L4_v11", 64, 4, (1, 1), n_iter=20, keep_u=True)
    run("modelA_n64_eps1e-6", 64, None, (1, 1), EPS=1e-6)
    run("modelA_n64_rhs", 64, None, (1, 1), n_iter=12, rhs_seed=5, keep_u=True)
    run("modelA_n256_v11", 256, None, (1, 1), n_iter=15)
    run("modelA_n1024_L10", 1024, 10, (1, 1), n_iter=15)
    run("modelA_n1024_L8", 1024, 8, (1, 1), n_iter=10)
    run("modelA_n1024_rhs", 1024, 10, (1, 1), n_iter=8, rhs_seed=5)
    if os.environ.get("MGFEA_GOLDEN_BIG", "1") == "1":
        run("modelA_n4096_L12", 4096, 12, (1, 1), n_iter=13)

    # fp64 reference histories (noise band for the per-cycle tolerance, SURVEY section 7 "hard parts")
    def run64(tag, n, L, n_iter, rhs_seed=None):
        np.random.seed(123)
        with contextlib.redirect_stdout(io.StringIO()):
            p = MG(n, L)
        for lv in p.grids.values():
            lv.Knet.double()
            lv.Knet.global_pattern = lv.Knet.global_pattern.double()
            lv.jac.geometry_idx = lv.jac.geometry_idx.double()
            lv.jac.boundary_value = lv.jac.boundary_value.double()
            lv.jac.d_mat = lv.jac.d_mat.double()
            lv.v = lv.v.double()
            lv.f = lv.f.double()
            lv.fnet.double()
        p.initial_v = torch.from_numpy(model_problem_u0(n)).double()
        if rhs_seed is not None:
            rs = np.random.RandomState(rhs_seed)
            F = torch.from_numpy(rs.standard_normal((1, 1, n + 1, n + 1)).astype(np.float32)).double()
            with torch.no_grad():
                p.grids[0].f = p.grids[0].fnet.float()(F.float()).double()
            p.initial_v = torch.zeros(n + 1, n + 1).double()
        # Restrict builds an fp32 kernel: patch through a double-capable copy of the same code path
        import torch.nn.functional as F_

        def Restrict(f):
            k = torch.asarray([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=torch.float64) / 16.0
            return F_.pad(F_.conv2d(f[:, :, 1:-1, 1:-1], k.view(1, 1, 3, 3), stride=2), (1, 1, 1, 1), "constant", 0)

        p.Restrict = Restrict
        with torch.no_grad():
            res = p.Solve([1, 1], n_iter=n_iter)
        hist[tag]["res64"] = [float(r) for r in res]

    run64("modelA_n64_v11", 64, None, 20)
    run64("modelA_n64_rhs", 64, None, 12, rhs_seed=5)
    run64("modelA_n1024_L10", 1024, 10, 15)
    run64("modelA_n1024_L8", 1024, 8, 10)
    run64("modelA_n64_L4_v11", 64, 4, 20)

    # ---- MM_Interface_error.ipynb: two-phase circle a=[1,20], n=64, f=fnet(ones), u0=0, EPS=5e-5 (quirk variant)
    nsI = H.notebook_namespace("MM_Interface_error.ipynb", [0, 1, 2])
    t0 = time.time()
    prob = nsI["Multigrid"](64)
    prob.grids[0].v = torch.zeros((1, 1, 65, 65), dtype=torch.float32)
    res_list = []
    res = 1
    with torch.no_grad():
        while abs(res) > 5e-5 and len(res_list) < 40:
            prob.rec_V_cycle(0, prob.grids[0].v, prob.grids[0].f)
            r = prob.grids[0].f - prob.grids[0].Knet(prob.grids[0].v)
            res = torch.sqrt(torch.sum(r[:, :, 1:-1, 1:-1] ** 2)).item()
            res_list.append(res)
    hist["interface_quirk_n64"] = dict(n=64, prop=[1, 20], res=res_list,
                                       notebook_recorded=[0.04344373568892479, 0.025038596242666245,
                                                          0.016153400763869286, 0.0099326865747571,
                                                          0.005999982822686434, 0.0035448919516056776,
                                                          0.002057234989479184, 0.0011781713692471385,
                                                          0.000666382780764252, 0.0003720286185853183,
                                                          0.00020798530022148043, 0.00011407280544517562,
                                                          6.423080776585266e-05, 3.4823522582883015e-05])
    arrays["interface_quirk_n64_u"] = t2n(prob.grids[0].v)
    print("interface_quirk", len(res_list), "%.1fs" % (time.time() - t0), res_list[:3])

    # ---- M-FEANet-mg_test.ipynb MultiGrid (variant B, 1-channel R=P=[1 2 1;2 4 2;1 2 1]/4), jac and hjac,
    #      on Data/IsoPoisson/poisson2d_33x33.h5 samples 0..2 (data Dirichlet BCs on level 0), EPS=5e-5
    bi, bv, rhs, uex = H.read_h5_contiguous(os.path.join(H.REF, "Data/IsoPoisson/poisson2d_33x33.h5"), (100, 33, 33))
    arrays["iso33_bidx"] = bi[:3].astype(np.float32)
    arrays["iso33_bval"] = bv[:3].astype(np.float32)
    arrays["iso33_rhs"] = rhs[:3].astype(np.float32)
    arrays["iso33_u"] = uex[:3].astype(np.float32)
    # known-answer: K u = fnet(rhs) in fp64
    arrays["iso33_rhs64"] = rhs[:3].copy()
    arrays["iso33_u64"] = uex[:3].copy()
    nsT = H.notebook_namespace("M-FEANet-mg_test.ipynb", [1, 2, 3, 4, 5, 18, 19, 20])
    hnet = nsT["HNet"](3)
    hnet.load_state_dict(torch.load(os.path.join(H.REF, "Model/learn_iterator/iso_poisson/iso_poisson_33x33.pth"),
                                    weights_only=True))
    for mode in ("jac", "hjac"):
        for k in range(3):
            n = 32
            mg = nsT["MultiGrid"](n=n, hnet=hnet, P=nsT["linear_tensor_P"], mode=mode)
            f_mg = torch.from_numpy(arrays["iso33_rhs"][k]).reshape(1, 1, n + 1, n + 1)
            bidx = torch.from_numpy(arrays["iso33_bidx"][k]).reshape(1, 1, n + 1, n + 1)
            bval = torch.from_numpy(arrays["iso33_bval"][k]).reshape(1, 1, n + 1, n + 1)
            u_mg = torch.zeros((1, 1, n + 1, n + 1), dtype=torch.float32)
            with torch.no_grad():
                mg(u_mg, f_mg, bidx, bval, 1)
                r = mg.f - mg.iterators[0].grid.Knet(mg.u0)
                res = torch.norm(r[:, :, 1:-1, 1:-1].clone(), dim=(2, 3)).item()
                rl = [res]
                while abs(res) > 5e-5 and len(rl) < 60:
                    u_mg = mg.Step(u_mg, mg.f)
                    r = mg.f - mg.iterators[0].grid.Knet(u_mg)
                    res = torch.norm(r[:, :, 1:-1, 1:-1].clone(), dim=(2, 3)).item()
                    rl.append(res)
            hist[f"mgtest_{mode}_s{k}"] = dict(n=n, mode=mode, sample=k, res=rl)
            arrays[f"mgtest_{mode}_s{k}_u"] = t2n(u_mg)
            print("mgtest", mode, k, len(rl) - 1, "cycles", rl[:3])

    # ---- committed FEANet/multigrid.py MultiGrid.iterate (16-channel R/P, w) with the n_iter shim, n=32,
    #      two-phase circle [1,20]; (a) linear R/P, w=[4,1]; (b) learned weights from Model/
    M = H.load_reference_multigrid_module()
    P4 = torch.tensor([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=torch.float32) / 4.0
    R16 = torch.tensor([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=torch.float32) / 16.0
    sd = torch.load(os.path.join(H.REF, "Model/learn_intergrid_operator/multigrid_rhs_qm/"
                                 "model_multigrid_interface_ratio.pth"), weights_only=True)
    arrays["learned_w"] = t2n(sd["w"])
    arrays["learned_R"] = t2n(sd["conv.net.weight"]).reshape(16, 9)
    arrays["learned_P"] = t2n(sd["deconv.net.weight"]).reshape(16, 9)
    n = 32
    rs = np.random.RandomState(11)
    Fr = torch.from_numpy(rs.standard_normal((2, 1, n + 1, n + 1)).astype(np.float32))
    x0 = torch.from_numpy(rs.standard_normal((2, 1, n + 1, n + 1)).astype(np.float32))
    arrays["iterate_F"] = t2n(Fr)
    arrays["iterate_x0"] = t2n(x0)
    for tag, learned in (("linear", False), ("learned", True)):
        mg = M.MultiGrid(n, R16, P4, torch.tensor([4.0, 1.0]))
        if learned:
            mg.load_state_dict(sd)
        with torch.no_grad():
            f = mg.grids[0].fnet(Fr)
            x = x0
            rl = []
            for _ in range(6):
                x = mg.iterate(x, f)
                r = f - mg.grids[0].Knet(x)
                rl.append(torch.norm(r[:, :, 1:-1, 1:-1], dim=(2, 3)).reshape(-1).tolist())
        hist[f"iterate_{tag}"] = dict(n=n, res=rl)
        arrays[f"iterate_{tag}_u"] = t2n(x)
        print("iterate", tag, rl[0], rl[-1])

    json.dump(hist, open(os.path.join(OUT, "solve_histories.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(OUT, "solve_arrays.npz"), **arrays)
    print("solve fixtures written")


# ---------------------------------------------------------------------------------------------------------
# fp64 noise bands for the histories that had none, and BASELINE config 3 through the unmodified reference
# ---------------------------------------------------------------------------------------------------------
def closed_form_keys(N, shape=0):
    """pattern key per node of the two-phase plate in integer arithmetic; asserted below against the
    reference's own MeshCenterInterface before it is used to feed the reference classes at sizes its O(N^4) setup
    cannot reach"""
    n = N - 1
    c = 2 * np.arange(n) + 1 - n
    if shape == 0:
        ph = (4 * (c[None, :] ** 2 + c[:, None] ** 2) < n * n).astype(np.int64)
    else:
        ph = ((2 * np.abs(c[None, :]) < n) & (2 * np.abs(c[:, None]) < n)).astype(np.int64)
    pe = np.zeros((n + 2, n + 2), np.int64)
    pe[1:-1, 1:-1] = ph  # pe[r+1, c+1] = phase of element (r, c)
    i, j = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    e1, e2, e3, e4 = pe[i, j + 1], pe[i, j], pe[i + 1, j], pe[i + 1, j + 1]  # (i-1,j), (i-1,j-1), (i,j-1), (i,j)
    ref = {0: [0, 0, 0, 0], 1: [1, 1, 1, 1], 2: [0, 0, 0, 1], 3: [0, 0, 1, 0], 4: [1, 0, 0, 0], 5: [0, 1, 0, 0],
           6: [0, 0, 1, 1], 7: [1, 1, 0, 0], 8: [0, 1, 1, 0], 9: [1, 0, 0, 1], 10: [0, 1, 0, 1], 11: [1, 0, 1, 0],
           12: [1, 1, 1, 0], 13: [1, 1, 0, 1], 14: [0, 1, 1, 1], 15: [1, 0, 1, 1]}
    lut = np.zeros(16, np.int64)
    for k, p in ref.items():
        lut[p[0] * 8 + p[1] * 4 + p[2] * 2 + p[3]] = k
    keys = lut[e1 * 8 + e2 * 4 + e3 * 2 + e4]
    keys[0, :] = keys[-1, :] = 0
    keys[:, 0] = keys[:, -1] = 0
    return keys.astype(np.uint8)


def _to_double(grid):
    grid.Knet.double()
    grid.Knet.global_pattern = grid.Knet.global_pattern.double()
    grid.jac.geometry_idx = grid.jac.geometry_idx.double()
    grid.jac.boundary_value = grid.jac.boundary_value.double()
    grid.jac.d_mat = grid.jac.d_mat.double()
    grid.v = grid.v.double()
    grid.f = grid.f.double()
    grid.fnet.double()


def gen_bands():
    import torch.nn.functional as F_

    out = {}
    # ---- MM_Interface_error (quirk variant), fp64
    nsI = H.notebook_namespace("MM_Interface_error.ipynb", [0, 1, 2])
    prob = nsI["Multigrid"](64)
    for g in prob.grids.values():
        _to_double(g)
    prob.grids[0].f = prob.grids[0].fnet(torch.ones(1, 1, 65, 65, dtype=torch.float64))

    def Restrict(f):
        k = torch.asarray([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=torch.float64) / 16.0
        return F_.pad(F_.conv2d(f[:, :, 1:-1, 1:-1], k.view(1, 1, 3, 3), stride=2), (1, 1, 1, 1), "constant", 0)

    prob.Restrict = Restrict
    prob.grids[0].v = torch.zeros((1, 1, 65, 65), dtype=torch.float64)
    rl = []
    with torch.no_grad():
        for _ in range(14):
            prob.rec_V_cycle(0, prob.grids[0].v, prob.grids[0].f)
            r = prob.grids[0].f - prob.grids[0].Knet(prob.grids[0].v)
            rl.append(torch.sqrt(torch.sum(r[:, :, 1:-1, 1:-1] ** 2)).item())
    out["interface_quirk_n64"] = dict(res64=rl)
    print("interface_quirk fp64", rl[:2], rl[-1])

    # ---- M-FEANet-mg_test MultiGrid on the IsoPoisson 33^2 samples, fp64
    bi, bv, rhs, uex = H.read_h5_contiguous(os.path.join(H.REF, "Data/IsoPoisson/poisson2d_33x33.h5"), (100, 33, 33))
    nsT = H.notebook_namespace("M-FEANet-mg_test.ipynb", [1, 2, 3, 4, 5, 18, 19, 20])
    sd = torch.load(os.path.join(H.REF, "Model/learn_iterator/iso_poisson/iso_poisson_33x33.pth"), weights_only=True)
    for mode in ("jac", "hjac"):
        for k in range(3):
            n = 32
            hnet = nsT["HNet"](3)
            hnet.load_state_dict(sd)
            hnet.double()
            mg = nsT["MultiGrid"](n=n, hnet=hnet, P=nsT["linear_tensor_P"], mode=mode)
            mg.conv.double()
            mg.deconv.double()
            for it in mg.iterators.values():
                _to_double(it.grid)
            # the fp32 run casts the dataset to fp32 first: same inputs here, widened
            f_mg = torch.from_numpy(rhs[k].astype(np.float32)).double().reshape(1, 1, n + 1, n + 1)
            bidx = torch.from_numpy(bi[k].astype(np.float32)).double().reshape(1, 1, n + 1, n + 1)
            bval = torch.from_numpy(bv[k].astype(np.float32)).double().reshape(1, 1, n + 1, n + 1)
            u_mg = torch.zeros((1, 1, n + 1, n + 1), dtype=torch.float64)
            want = len(json.load(open(os.path.join(OUT, "solve_histories.json")))[f"mgtest_{mode}_s{k}"]["res"])
            with torch.no_grad():
                mg(u_mg, f_mg, bidx, bval, 1)
                r = mg.f - mg.iterators[0].grid.Knet(mg.u0)
                rl = [torch.norm(r[:, :, 1:-1, 1:-1].clone(), dim=(2, 3)).item()]
                while len(rl) < want:
                    u_mg = mg.Step(u_mg, mg.f)
                    r = mg.f - mg.iterators[0].grid.Knet(u_mg)
                    rl.append(torch.norm(r[:, :, 1:-1, 1:-1].clone(), dim=(2, 3)).item())
            out[f"mgtest_{mode}_s{k}"] = dict(res64=rl)
            print("mgtest fp64", mode, k, rl[:2], rl[-1])

    # ---- BASELINE config 3: two-phase circle 1:100, 16-channel linear R/P, w = [4, 1], F = ones, u0 = 0, through the
    #      UNMODIFIED FEANet/multigrid.py MultiGrid.iterate (n_iter shim) fed with a closed-form duck-typed mesh (the
    #      reference's MeshCenterInterface is O(N^4)); 'hjac': each level's Relax is the unmodified mg_test
    #      HJacIterator.HRelax with the shipped iso_poisson_33x33.pth weights
    M = H.load_reference_multigrid_module()
    from FEANet.mesh import MeshCenterInterface as RefMesh

    PROP = [1, 100]
    kd = RefMesh(2, PROP, 9).kernel_dict  # the 16 x (3x3) table is size independent: the reference builds it
    for n in (8, 16, 32):  # the closed form against the reference's own key maps
        m = RefMesh(2, PROP, n + 1)
        assert np.array_equal(keys_of(m), closed_form_keys(n + 1, 0)), n

    class FastMesh:
        def __init__(self, size, prop, nnode_edge, shape=0):
            self.nnode_edge = int(nnode_edge)
            self.kernel_dict = kd
            k = closed_form_keys(self.nnode_edge, shape).reshape(-1)
            self.global_pattern_center = {key: (k == key).astype(int) for key in kd}

    HJ = nsT["HJacIterator"]

    def make(n, mode, double):
        hnet = nsT["HNet"](3)
        hnet.load_state_dict(sd)

        class HGrid(M.SingleGrid):
            def Relax(self, v, f, k):
                if not hasattr(self, "_it"):
                    self._it = HJ(n=self.n, hnet=hnet, grid=self)
                return self._it.HRelax(v, f, k)

        saved = M.MeshCenterInterface, M.SingleGrid
        M.MeshCenterInterface = FastMesh
        if mode == "hjac":
            M.SingleGrid = HGrid
        try:
            P4 = torch.tensor([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=torch.float32) / 4.0
            R16 = torch.tensor([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=torch.float32) / 16.0
            mg = M.MultiGrid(n, R16, P4, torch.tensor([4.0, 1.0]))
        finally:
            M.MeshCenterInterface, M.SingleGrid = saved
        if double:
            hnet.double()
            mg.double()
            for g in mg.grids.values():
                _to_double(g)
        return mg

    for n, ncyc in ((64, 10), (256, 8)):
        for mode in ("jac", "hjac"):
            rec = {}
            for double in (False, True):
                t0 = time.time()
                mg = make(n, mode, double)
                dt = torch.float64 if double else torch.float32
                with torch.no_grad():
                    f = mg.grids[0].fnet(torch.ones(1, 1, n + 1, n + 1, dtype=dt))
                    x = torch.zeros(1, 1, n + 1, n + 1, dtype=dt)
                    r = f - mg.grids[0].Knet(x)
                    r0 = torch.norm(r[:, :, 1:-1, 1:-1], dim=(2, 3)).item()
                    rl = []
                    for _ in range(ncyc):
                        x = mg.iterate(x, f)
                        r = f - mg.grids[0].Knet(x)
                        rl.append(torch.norm(r[:, :, 1:-1, 1:-1], dim=(2, 3)).item())
                rec["res64" if double else "res"] = rl
                rec["r0_64" if double else "r0"] = r0
                print("cfg3", n, mode, "fp64" if double else "fp32", "%.1fs" % (time.time() - t0), r0, rl)
            out[f"cfg3_{mode}_n{n}"] = dict(n=n, prop=PROP, mode=mode, w=[4.0, 1.0], **rec)
    json.dump(out, open(os.path.join(OUT, "bands.json"), "w"), indent=1)
    print("bands.json written")


def gen_testpoisson():
    """Data/TestPoisson/poisson2d_33x33.h5 (Data/dataset.py:71-104 TestPoissonDataSet): per-ELEMENT `material` (32 x 32),
    Dirichlet masks, source, FEM solution -- first 3 samples, as the fixture of the per-element operator"""
    p = os.path.join(H.REF, "Data/TestPoisson/poisson2d_33x33.h5")
    b33 = H.read_h5_contiguous(p, (10, 33, 33))  # file order: dirich_idx, dirich_value, neumann_idx, neumann_value, source, solution
    b32 = H.read_h5_contiguous(p, (10, 32, 32))
    assert len(b33) == 6 and len(b32) == 1
    out = {"material": b32[0][:3].copy(), "dirich_idx": b33[0][:3].copy(), "dirich_value": b33[1][:3].copy(),
           "neumann_idx": b33[2][:3].copy(), "neumann_value": b33[3][:3].copy(), "source": b33[4][:3].copy(),
           "solution": b33[5][:3].copy()}
    # known answer with the reference's own modules in fp64 (MM_poisson.ipynb cell 2: KNet(MeshSquare).double())
    H.load_reference()
    from FEANet.mesh import MeshSquare
    from FEANet.model import FNet, KNet

    knet, fnet = KNet(MeshSquare(2, 33)).double(), FNet(2.0 / 32).double()
    knet.global_pattern = knet.global_pattern.double()
    with torch.no_grad():
        r = fnet(torch.from_numpy(out["source"])[:, None]) - knet(torch.from_numpy(out["solution"])[:, None])
    out["ref_residual_interior_max"] = np.array([float(r[:, :, 1:-1, 1:-1].abs().max())])
    print("testpoisson: max interior |fnet(source) - K solution| (fp64, reference modules) =", out["ref_residual_interior_max"])
    np.savez_compressed(os.path.join(OUT, "testpoisson.npz"), **out)


def gen_pbc():
    """JacobiBlockPBC (FEANet/jacobi.py:50-97) of the unmodified reference on random fields: 1 and 3 sweeps, the
    reset_boundary / pbc_boundary helpers, d_mat"""
    H.load_reference()
    from FEANet.jacobi import JacobiBlockPBC
    from FEANet.mesh import MeshSquare
    from FEANet.model import KNet

    g = torch.Generator().manual_seed(77)
    out = {}
    for n in (8, 16, 32):
        N = n + 1
        mesh = MeshSquare(2, N)
        jac = JacobiBlockPBC(mesh, KNet(mesh))
        u = torch.randn(2, 1, N, N, generator=g)
        fp = torch.randn(2, 1, N + 2, N + 2, generator=g)
        with torch.no_grad():
            out[f"u_{n}"], out[f"fpad_{n}"] = t2n(u), t2n(fp)
            out[f"pbc_{n}"], out[f"reset_{n}"] = t2n(jac.pbc_boundary(u)), t2n(jac.reset_boundary(u))
            v = jac.jacobi_convolution(u, fp)
            out[f"jac1_{n}"] = t2n(v)
            for _ in range(2):
                v = jac.jacobi_convolution(v, fp)
            out[f"jac3_{n}"] = t2n(v)
            out[f"dmat_{n}"] = t2n(jac.d_mat)
    np.savez_compressed(os.path.join(OUT, "pbc.npz"), **out)
    print("pbc.npz", len(out))


def gen_hgrad():
    """gradients of the reference's HNet training step (M-FEANet-learn_iterator.ipynb cell 8 / mg_test cell 5:
    loss = MSELoss(sum)(HRelax(uu, fnet(f), k), u) with per-sample Dirichlet masks) by the reference's own autograd"""
    nsT = H.notebook_namespace("M-FEANet-mg_test.ipynb", [1, 2, 3, 4, 5])
    sd = torch.load(os.path.join(H.REF, "Model/learn_iterator/iso_poisson/iso_poisson_33x33.pth"), weights_only=True)
    bi, bv, rhs, uex = H.read_h5_contiguous(os.path.join(H.REF, "Data/IsoPoisson/poisson2d_33x33.h5"), (100, 33, 33))
    out = {}
    g = torch.Generator().manual_seed(99)
    for k in (1, 3):
        hnet = nsT["HNet"](3)
        hnet.load_state_dict(sd)
        it = nsT["HJacIterator"](n=32, hnet=hnet)
        B = 2
        u_train = torch.from_numpy(uex[:B].astype(np.float32))[:, None]
        f_train = torch.from_numpy(rhs[:B].astype(np.float32))[:, None]
        bval = torch.from_numpy(bv[:B].astype(np.float32))[:, None]
        bidx = torch.from_numpy(bi[:B].astype(np.float32))[:, None]
        it.grid.ResetBoundary(bidx, bval)
        ff = it.grid.fnet(f_train)
        uu = torch.randn(B, 1, 33, 33, generator=g).requires_grad_(True)
        u_out = it.HRelax(uu, ff, k)
        loss = torch.nn.MSELoss(reduction="sum")(u_out, u_train)
        loss.backward()
        out[f"uu_k{k}"], out[f"uout_k{k}"], out[f"loss_k{k}"] = t2n(uu), t2n(u_out), np.array([loss.item()])
        out[f"guu_k{k}"] = t2n(uu.grad)
        for l, layer in enumerate(hnet.convLayers):
            out[f"gw{l}_k{k}"] = t2n(layer.weight.grad).reshape(9)
        print("hgrad k", k, loss.item(), [float(np.abs(out[f"gw{l}_k{k}"]).max()) for l in range(3)])
    np.savez_compressed(os.path.join(OUT, "hgrad.npz"), **out)


def gen_mggrad():
    """one training step of the committed FEANet/multigrid.py MultiGrid (n_iter shim): q = qm(forward(F)); q.backward()
    -> gradients of the 16-channel R / P kernels by the reference's own autograd (multigrid.py:98-100,132-157)"""
    M = H.load_reference_multigrid_module()
    out = {}
    for n in (16, 32):
        P4 = torch.tensor([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=torch.float32) / 4.0
        R16 = torch.tensor([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=torch.float32) / 16.0
        mg = M.MultiGrid(n, R16, P4, torch.tensor([4.0, 1.0]))
        mg.w.requires_grad_(True)  # also record the gradient of the ratios (some shipped models have learned w)
        gp = torch.Generator().manual_seed(3 + n)
        with torch.no_grad():  # perturb so that the 16 channels differ
            mg.conv.net.weight += 0.02 * torch.randn(mg.conv.net.weight.shape, generator=gp)
            mg.deconv.net.weight += 0.02 * torch.randn(mg.deconv.net.weight.shape, generator=gp)
        rs = np.random.RandomState(21 + n)
        F = torch.from_numpy(rs.standard_normal((2, 1, n + 1, n + 1)).astype(np.float32))
        np.random.seed(5)
        u = mg(F)
        q = mg.qm(u)
        q.backward()
        out[f"F_{n}"], out[f"R_{n}"], out[f"P_{n}"] = t2n(F), t2n(mg.conv.net.weight), t2n(mg.deconv.net.weight)
        out[f"u_{n}"], out[f"q_{n}"] = t2n(u), np.array([q.item()])
        out[f"gR_{n}"], out[f"gP_{n}"], out[f"gw_{n}"] = t2n(mg.conv.net.weight.grad), t2n(mg.deconv.net.weight.grad), t2n(mg.w.grad)
        print("mggrad", n, q.item(), float(mg.conv.net.weight.grad.abs().max()), float(mg.deconv.net.weight.grad.abs().max()),
              mg.w.grad.tolist())
    np.savez_compressed(os.path.join(OUT, "mggrad.npz"), **out)


def gen_h5manifest():
    """names / shapes / dtypes / data hashes of the reference's HDF5 files as read by the product's own reader
    (FEANet/h5lite.py), cross-checked against the byte-scanning reader this harness has used since round 1"""
    import hashlib

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "multigrid-feanet_b200"))
    for k in [k for k in sys.modules if k == "FEANet" or k.startswith("FEANet.")]:
        del sys.modules[k]
    if H.REF in sys.path:
        sys.path.remove(H.REF)
    from FEANet.h5lite import H5File

    man = {}
    for rel, shp in (("Data/IsoPoisson/poisson2d_33x33.h5", (100, 33, 33)), ("Data/TestPoisson/poisson2d_33x33.h5", None),
                     ("Data/RHS/poisson2d_rhs_17x17.h5", None)):
        f = H5File(os.path.join(H.REF, rel))
        man[rel] = {k: dict(shape=list(f.shape(k)), dtype=str(f[k].dtype),
                            sha256=hashlib.sha256(f[k].tobytes()).hexdigest()) for k in f.keys()}
        if shp:
            blocks = H.read_h5_contiguous(os.path.join(H.REF, rel), shp)
            assert all(np.array_equal(f[k], b) for k, b in zip(f.order, blocks))
    json.dump(man, open(os.path.join(OUT, "h5_manifest.json"), "w"), indent=1)
    print("h5_manifest.json", {k: list(v) for k, v in man.items()})


if __name__ == "__main__":
    which = sys.argv[1:] or ["mesh", "ops", "solve", "bands", "testpoisson", "pbc", "hgrad", "mggrad", "h5"]
    if "mesh" in which:
        gen_mesh()
    if "ops" in which:
        gen_ops()
    if "solve" in which:
        gen_solve()
    if "bands" in which:
        gen_bands()
    if "testpoisson" in which:
        gen_testpoisson()
    if "pbc" in which:
        gen_pbc()
    if "hgrad" in which:
        gen_hgrad()
    if "mggrad" in which:
        gen_mggrad()
    if "h5" in which:
        gen_h5manifest()
