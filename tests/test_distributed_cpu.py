"""Multi-rank logic of the row-slab V-cycle on CPU: world_size 2 (and 4) with the gloo backend, the ORACLE as the local
operator.  Checks that partition + halo exchange + coarse agglomeration reproduce the single-process oracle bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleSlabOps:
    """local slab operators emulated with the oracle: embed the local rows (owned + ghost) in a zero full-size array,
    apply the full-domain operator, keep the owned rows (valid because every operator has a 3-row dependency cone)"""

    def __init__(self, part):
        from oracle import oracle as O

        self.O, self.part = O, part
        self.levels = O.make_levels(part.n, part.L)
        Nc = part.levels[part.ld]["N"] if part.ld < part.L else 0
        self.fc_full = torch.zeros((1, Nc, Nc), dtype=torch.float32)
        self.uc_full = torch.zeros((1, Nc, Nc), dtype=torch.float32)

    def alloc(self, l):
        lev = self.part.levels[l]
        return torch.zeros((1, lev["nrows"], lev["N"]), dtype=torch.float32)

    def coarse_f(self):
        return self.fc_full

    def coarse_u(self):
        return self.uc_full

    def _full(self, l, arr):
        lev = self.part.levels[l]
        full = np.zeros((1, lev["N"], lev["N"]), np.float32)
        if arr is not None:
            full[0, lev["row0"]:lev["row0"] + lev["nrows"]] = arr[0].numpy()
        return full

    def down(self, l, u_in, u_out, f, fc):
        O, lv = self.O, self.levels[l]
        lev, levc = self.part.levels[l], self.part.levels[l + 1]
        ff = self._full(l, f)
        u1 = O.jacobi(self._full(l, u_in), ff, lv.keys, lv.ktab, lv.invd)
        fcg = O.restrict(O.residual(u1, ff, lv.keys, lv.ktab), None, O.FW16, 4.0)
        o0, o1, r0 = lev["dn0"], lev["dn1"], lev["row0"]  # owned rows + the deep-halo rows this rank computes itself
        u_out[0, o0 - r0:o1 - r0] = torch.from_numpy(u1[0, o0:o1])
        c0, c1 = o0 // 2, (o1 + 1) // 2 if o1 == lev["N"] else o1 // 2
        fc[0, c0 - levc["row0"]:c1 - levc["row0"]] = torch.from_numpy(fcg[0, c0:c1])

    def up(self, l, vc, u_in, u_out, f, want_norm):
        O, lv = self.O, self.levels[l]
        lev = self.part.levels[l]
        ff = self._full(l, f)
        ucorr = O.prolong_bilinear(self._full(l + 1, vc), self._full(l, u_in))
        u2 = O.jacobi(ucorr, ff, lv.keys, lv.ktab, lv.invd)
        o0, o1, r0 = lev["up0"], lev["up1"], lev["row0"]
        u_out[0, o0 - r0:o1 - r0] = torch.from_numpy(u2[0, o0:o1])
        if not want_norm:
            return None
        o0, o1 = lev["own0"], lev["own1"]
        r = O.residual(u2, ff, lv.keys, lv.ktab)[0]
        rows = r[max(o0, 1):min(o1, lev["N"] - 1), 1:-1].astype(np.float64)
        return torch.tensor([float((rows * rows).sum())], dtype=torch.float64)

    def coarse_cycle(self):
        O, ld = self.O, self.part.ld
        u = O.vcycle(self.levels[ld:], O.CycleCfg(), np.zeros_like(self.fc_full.numpy()), self.fc_full.numpy())
        self.uc_full.copy_(torch.from_numpy(u))


def _worker(rank, world, n, dist_min_n, port, ret):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "multigrid-feanet_b200"), os.path.join(ROOT, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from FEANet.distributed import SlabMultigrid
        from test_distributed_cpu import OracleSlabOps as Ops

        rs = np.random.RandomState(3)
        u0 = rs.standard_normal((n + 1, n + 1)).astype(np.float32)
        f = 0.01 * rs.standard_normal((n + 1, n + 1)).astype(np.float32)
        mg = SlabMultigrid(n, ops_factory=Ops, dist_min_n=dist_min_n)
        mg.set_problem(torch.from_numpy(u0), torch.from_numpy(f))
        hist = mg.Solve(n_iter=3)
        sol = mg.gather_solution()
        if rank == 0:
            ret["hist"], ret["sol"], ret["ld"] = hist, sol.numpy(), mg.part.ld
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,dist_min_n", [(2, 64, 17), (4, 128, 33), (2, 32, 33)])
def test_row_slab_vcycle_matches_single_process(world, n, dist_min_n):
    from oracle import oracle as O

    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, n, dist_min_n, port, ret), nprocs=world, join=True)
    rs = np.random.RandomState(3)
    u0 = rs.standard_normal((n + 1, n + 1)).astype(np.float32)
    f = 0.01 * rs.standard_normal((n + 1, n + 1)).astype(np.float32)
    levels = O.make_levels(n)
    u, hist = O.solve(levels, O.CycleCfg(), u0, f[None], n_iter=3)
    assert ret["ld"] >= 1, "test must exercise distributed levels"
    assert np.array_equal(ret["sol"], u[0]), "slab-partitioned V-cycle differs from the single-process cycle"
    assert np.allclose(ret["hist"], hist, rtol=1e-12)


def test_partition_properties():
    sys.path[:0] = [os.path.join(ROOT, "multigrid-feanet_b200")]
    from FEANet.distributed import GHOST, SlabPartition

    for world in (2, 4, 8):
        for n in (8192, 16384):
            L = int(np.log2(n))
            parts = [SlabPartition(n, L, world, r) for r in range(world)]
            ld = parts[0].ld
            assert 1 <= ld < L
            for l in range(ld):
                N = n // 2 ** l + 1
                rows = []
                for p in parts:
                    lev = p.levels[l]
                    assert lev["dist"] and lev["own0"] % 2 == 0
                    assert lev["row0"] <= max(0, lev["own0"] - 3) and lev["row0"] + lev["nrows"] >= min(N, lev["own1"] + 3)
                    rows += list(range(lev["own0"], lev["own1"]))
                    if l + 1 < ld:  # fine row 2I and coarse row I on the same rank
                        assert p.levels[l + 1]["own0"] * 2 == lev["own0"]
                assert rows == list(range(N)), "owned rows must tile the level exactly once"
            assert all(not p.levels[ld]["dist"] for p in parts)
    assert GHOST >= 3


# ---------------------------------------------------------------------------------------------------------------------
# peer-memory exchange: the per-rank plans (FEANet.distributed.SlabExchangePlan, pure host logic) checked against each
# other and executed on byte buffers -- what mgfea_p2p_exchange does with them on the GPUs
def _plans(n, world, dist_min_n):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "multigrid-feanet_b200")]
    import mgfea
    from FEANet.distributed import SlabExchangePlan, SlabPartition

    L = int(np.log2(n))
    return [SlabExchangePlan(SlabPartition(n, L, world, r, dist_min_n), mgfea.pitch_for) for r in range(world)]


@pytest.mark.parametrize("world,n,dist_min_n", [(2, 512, 129), (4, 1024, 257), (8, 16384, 2049), (8, 2048, 257)])
def test_exchange_plans_are_consistent(world, n, dist_min_n):
    from FEANet.distributed import GHOST

    plans = _plans(n, world, dist_min_n)
    ld = plans[0].part.ld
    assert 0 < ld
    assert all(p.off == plans[0].off and p.nbytes == plans[0].nbytes for p in plans)  # one layout for all ranks
    # the steps of the deep-halo cycle: initial exchange, gather, push of the new iterate + norm all-reduce, and the
    # fp64 steps of the mixed-precision solve
    steps = [((("u", 0), ("f", 0)), False, False), ((), True, False), ((("u", 0),), False, True), ((), False, True),
             ((("u64", 0),), False, False), ((("f64", 0),), False, False)]
    for halos, gather, reduce in steps:
        pl = [p.plan(halos, gather, reduce) for p in plans]
        assert len({q["grid"] for q in pl}) == 1 and 1 <= pl[0]["grid"] <= 512  # the flags count CTAs: same on every rank
        # every flag a rank waits on is raised by exactly one rank, and nobody raises a flag that is not awaited
        raised = {}
        for r, q in enumerate(pl):
            assert len(q["jobs"]) <= 16 and len(q["signals"]) <= 8 and len(q["waits"]) <= 8
            for tgt, fo in q["signals"]:
                assert (tgt, fo) not in raised
                raised[(tgt, fo)] = r
        awaited = {(r, fo) for r, q in enumerate(pl) for fo in q["waits"]}
        assert set(raised) == awaited
        for r, q in enumerate(pl):  # a rank that receives rows from another rank also waits for that rank
            for so, tgt, do, nb in q["jobs"]:
                assert nb % 16 == 0 and so % 16 == 0 and do % 16 == 0
                if tgt != r:
                    assert raised_by(raised, tgt, r)
        # halo jobs land exactly in the target's ghost rows of the same array
        for r, p in enumerate(plans):
            for name, l in halos:
                rowb = p.pitch[l] * p.esize(name)
                for so, tgt, do, nb in pl[r]["jobs"]:
                    if not (p.off[(name, l)] <= so < p.off[(name, l)] + p.part.levels[l]["nrows"] * rowb) or nb == 16:
                        continue
                    lt = plans[tgt].part.levels[l]
                    g0 = (so - p.off[(name, l)]) // rowb + p.part.levels[l]["row0"]  # first global row sent
                    G = p.part.levels[l]["G"]
                    assert G >= GHOST and nb == G * rowb
                    assert p.part.levels[l]["own0"] <= g0 and g0 + G <= p.part.levels[l]["own1"]  # owned by sender
                    t0 = (do - p.off[(name, l)]) // rowb + lt["row0"]
                    assert t0 == g0 and (g0 + G <= lt["own0"] or g0 >= lt["own1"])  # ghost rows of the target
                    assert lt["row0"] <= g0 and g0 + G <= lt["row0"] + lt["nrows"]
    # gather: the replicated right-hand side is covered exactly once by the ranks' owned rows (last ring row stays zero)
    pl = [p.plan((), True, False) for p in plans]
    rowb, off = plans[0].pitch[ld] * 4, plans[0].off[("f", ld)]
    nl = n // 2 ** ld
    rows = np.zeros(nl + 1, int)
    for r, q in enumerate(pl):
        mine = [(so, nb) for so, tgt, do, nb in q["jobs"] if so >= off and so == do]
        assert len(mine) == world - 1 and len({m for m in mine}) == 1
        rows[(mine[0][0] - off) // rowb:(mine[0][0] - off + mine[0][1]) // rowb] += 1
    assert (rows[:nl] == 1).all() and rows[nl] == 0


def raised_by(raised, receiver, sender):
    return any(tgt == receiver and src == sender for (tgt, _), src in raised.items())


def test_exchange_plan_moves_the_right_rows():
    """execute the halo jobs of every rank on byte buffers filled with (array id, global row) codes"""
    from FEANet.distributed import GHOST

    world, n, dmin = 4, 1024, 257
    plans = _plans(n, world, dmin)
    mem = [np.zeros(p.nbytes, np.uint8) for p in plans]
    l, name = 0, "u"
    rowb = plans[0].pitch[l] * 4

    def rows_view(r):
        lev = plans[r].part.levels[l]
        o = plans[r].off[(name, l)]
        return mem[r][o:o + lev["nrows"] * rowb].view(np.float32).reshape(lev["nrows"], -1), lev

    for r in range(world):  # owned rows carry their global row index, ghost rows -1
        v, lev = rows_view(r)
        v[:] = -1
        for g in range(lev["own0"], lev["own1"]):
            v[g - lev["row0"]] = g
    for r in range(world):
        for so, tgt, do, nb in plans[r].plan(((name, l),))["jobs"]:
            mem[tgt][do:do + nb] = mem[r][so:so + nb]
    for r in range(world):
        v, lev = rows_view(r)
        for i in range(lev["nrows"]):
            g = lev["row0"] + i
            assert (v[i] == g).all(), (r, g)


@pytest.mark.parametrize("world,n,dist_min_n", [(2, 512, 129), (4, 1024, 257), (8, 16384, 2049), (8, 2048, 257), (1, 256, 129)])
def test_slab_partition_invariants(world, n, dist_min_n):
    """every row of every distributed level has exactly one owner, boundaries are even (fine row 2I and coarse row I
    share a rank), ghost rows stay inside the domain, levels below the threshold are replicated"""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "multigrid-feanet_b200")]
    from FEANet.distributed import GHOST, SlabPartition

    L = int(np.log2(n))
    parts = [SlabPartition(n, L, world, r, dist_min_n) for r in range(world)]
    ld = parts[0].ld
    assert all(p.ld == ld for p in parts)
    for l in range(L):
        N = n // 2 ** l + 1
        levs = [p.levels[l] for p in parts]
        if l >= ld:
            assert all(not v["dist"] and v["own0"] == 0 and v["own1"] == N and v["nrows"] == N for v in levs)
            continue
        assert N >= dist_min_n and all(v["dist"] for v in levs)
        owner = np.zeros(N, int)
        for r, v in enumerate(levs):
            assert v["own0"] % 2 == 0 and (v["own1"] % 2 == 0 or v["own1"] == N)
            assert v["row0"] == max(0, v["own0"] - v["G"]) and v["row0"] + v["nrows"] == min(N, v["own1"] + v["G"])
            owner[v["own0"]:v["own1"]] += 1
            # deep halos: the rows a leg computes (owned + redundantly computed) start on an even row, and the 3 rows
            # above / 2 below them that the fused kernels stream are held in memory
            for lo, hi in (("dn0", "dn1"), ("up0", "up1")):
                assert v[lo] % 2 == 0 and v[lo] <= v["own0"] and v[hi] >= v["own1"]
                assert v["row0"] <= max(0, v[lo] - 3) and v["row0"] + v["nrows"] >= min(N, v[hi] + 3)
            assert v["own1"] - v["own0"] >= v["G"]  # the pushed boundary rows are owned rows
            if l + 1 < ld:
                c = parts[r].levels[l + 1]
                # the coarse right-hand side this rank restricts covers what its next down leg reads (computed rows + 2:
                # residual and restriction stencils; the Jacobi sweep from the zero guess is pointwise) ...
                assert v["dn0"] // 2 <= max(0, c["dn0"] - 2) and (v["dn1"] // 2 >= min(c["N"], c["dn1"] + 2) or v["dn1"] == N)
                # ... and the coarse iterate its up leg prolongs from is what the coarse up leg computed
                assert c["up0"] <= max(0, (v["up0"] - 1) // 2) and c["up1"] >= min(c["N"], (v["up1"] + 1) // 2 + 1)
            # the pre-smoothed iterate the up leg reads (computed rows +- 1) was computed by this rank's down leg
            assert v["dn0"] <= max(0, v["up0"] - 1) and v["dn1"] >= min(N, v["up1"] + 1)
            if l == 0:  # the norm of the owned rows needs the post-smoothed iterate on one more row: corrected rows +- 2
                assert v["dn0"] <= max(0, v["own0"] - 2) and v["dn1"] >= min(N, v["own1"] + 2)
            if l + 1 < ld:  # coarse rows I = own0/2 .. live on the same rank
                c = parts[r].levels[l + 1]
                assert c["own0"] == v["own0"] // 2 and (c["own1"] == v["own1"] // 2 or c["own1"] == (N - 1) // 2 + 1)
        assert (owner == 1).all()


def test_bench_algorithmic_bytes_match_survey():
    """64 B per fine DOF per V(1,1) cycle = 1.0745 GB at 4097^2; 4.30 GB at 8193^2 and at 1025^2 x 64"""
    sys.path[:0] = [ROOT]
    import bench

    b = bench.algorithmic_bytes_per_cycle(4096, 12)
    assert abs(b / 4097 ** 2 - 64) < 0.1 and abs(b / 1e9 - 1.0745) < 1e-3
    assert abs(bench.algorithmic_bytes_per_cycle(8192, 13) / 1e9 - 4.30) < 0.01
    assert abs(bench.algorithmic_bytes_per_cycle(1024, 8, B=64) / 1e9 - 4.30) < 0.02
    assert bench.algorithmic_bytes_per_cycle(4096, 12, key_bytes=1) > b
