#!/usr/bin/env python
"""bench.py -- V-cycles/s and GDOF/s of the multigrid V-cycle hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W          # our arm (sm_100a kernels through the C ABI)
    python bench.py --impl reference --steps K --warmup W   # the reference's CPU torch path on the box's host cores

A "step" is one V(1,1) cycle (smooth, residual+restrict, ..., prolong+correct, smooth, interior residual norm) over the
headline workload: isotropic Poisson, 4097 x 4097 nodal quad grid, 12 levels, single right-hand side, fp32.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "multigrid-feanet_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "GDOF/s of V(1,1) cycles (= V-cycles/s x DOF; 4097^2 Poisson at N=1), fp32"
REPEATS = 5  # median of >= 5 repeats of the K timed steps


def size_for(world, override=0):
    """grid intervals of the headline workload: N=1 the size the metric is quoted on (4097^2); N=2,4 BASELINE config 4
    (8193^2); N=8 config 5 (16385^2, north_star's 8-GPU target).  Every N>1 line also carries the SAME problem timed on
    one GPU in the same run ("strong_scaling") and the other size ("configs")."""
    return override or {1: 4096, 2: 8192, 4: 8192, 8: 16384}.get(world, 8192)


def model_u0(n, seed=123):
    """MM_Model_convergence.ipynb cell 3 random_data (seed enabled): the reference's own benchmark problem, f = 0"""
    np.random.seed(seed)
    coef = 100000 + 50000 * np.random.rand(2)
    return (coef[0] * np.random.random((n + 1, n + 1)).astype("f") + coef[1]).astype(np.float32)


def algorithmic_bytes_per_cycle(n, L, nu1=1, nu2=1, B=1, key_bytes=0):
    """compulsory fp32 traffic of one V-cycle, no credit for temporal blocking"""
    M = [(n // 2 ** l + 1) ** 2 for l in range(L)]
    words = 3 * (nu1 + nu2) * sum(M) + 2 * sum(2 * M[l] + M[l + 1] for l in range(L - 1)) + 2 * M[0]
    kb = key_bytes * ((nu1 + nu2 + 1) * sum(M) + M[0])
    return (4 * words + kb) * B


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.strip().split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


NCU_PROFILE = "r02_ncu_full_stream2_final.json"


def ncu_traffic(kernel_prefix, grid=None):
    """DRAM bytes (read + write) per launch of the dominant kernel from the committed `ncu --set full` summary
    (profiles/NCU_PROFILE, captured with tools/ncu_cycle.py on the same 4097^2 workload)"""
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", NCU_PROFILE)))
        for name, d in prof.items():
            if kernel_prefix in name and (grid is None or f"({grid}," in name):
                def mb(v):
                    x, unit = v.split()[0], (v.split() + [""])[1].lower()
                    return float(x) * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(unit, 1e6)

                return int(mb(d["dram__bytes_read.sum"]) + mb(d["dram__bytes_write.sum"]))
    except Exception:
        pass
    return None


def hbm_peak():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------
def _reference_cycles(n, steps, warm, budget_s):
    """times V-cycles of the f=0 model problem on the host cores: the UNMODIFIED reference classes when oracle/_ref was
    vendored (oracle/vendor_ref.py), else the ATen call-for-call port.  Returns (cycles, seconds, kind, what)"""
    import torch

    from oracle import ref_runner as RR

    L = int(np.log2(n))
    if RR.available():
        prob = RR.model_problem(n, model_u0(n))
        with torch.no_grad():
            t0 = time.perf_counter()
            prob.Solve([1, 1], n_iter=1)  # warm-up (first-touch of every level, oneDNN primitive caches)
            t_first = time.perf_counter() - t0
            k = max(1, min(steps, int(budget_s / max(t_first, 1e-3))))
            for _ in range(max(0, min(warm, 2) - 1)):
                prob.Solve([1, 1], n_iter=1)
            t0 = time.perf_counter()
            prob.Solve([1, 1], n_iter=k)  # k cycles, each followed by the residual norm + .item() (the solve() path)
            dt = time.perf_counter() - t0
        return k, dt, "reference", ("the UNMODIFIED reference (oracle/_ref: FEANet/*.py + MM_Model_convergence.ipynb "
                                    f"cell 3 Multigrid({n}).Solve([1,1], n_iter={k}))")
    from oracle import feanet_torch as FT

    levels = FT.make_levels(n, L)
    u = torch.from_numpy(model_u0(n)).reshape(1, 1, n + 1, n + 1)
    f = torch.zeros(1, 1, n + 1, n + 1)
    with torch.no_grad():
        t0 = time.perf_counter()
        u = FT.vcycle(levels, u, f)
        t_first = time.perf_counter() - t0
        k = max(1, min(steps, int(budget_s / max(t_first, 1e-3))))
        for _ in range(max(0, min(warm, 2) - 1)):
            u = FT.vcycle(levels, u, f)
        t0 = time.perf_counter()
        for _ in range(k):
            u = FT.vcycle(levels, u, f)
            r = f - levels[0].K(u)
            _ = torch.sqrt(torch.sum(r[:, :, 1:-1, 1:-1] ** 2)).item()
        dt = time.perf_counter() - t0
    return k, dt, "port", "ATen call-for-call restatement of the reference CPU torch path (oracle/feanet_torch.py)"


def run_reference(args):
    """The reference's CPU torch path on the host cores, on OUR arm's workload for this N (size_for)."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, args.gpus)
    n = size_for(world, args.n_multi if world > 1 else (args.n if args.n != 4096 else 0))
    note = None
    try:
        import psutil

        need = 40 * 4 * (n + 1) ** 2 * 1.4  # mesh arrays (int64 cells) + fields + ATen temporaries, all levels
        if psutil.virtual_memory().available < need:
            note = f"host RAM too small for {n + 1}^2 ({need / 1e9:.0f} GB): timed 8193^2 instead"
            n = 8192
    except Exception:
        pass
    L = int(np.log2(n))
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    k, dt, kind, what = _reference_cycles(n, args.steps, args.warmup, 150.0)
    val = k / dt
    dof = (n + 1) ** 2
    gd = val * dof / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": gd, "unit": "GDOF/s", "v_cycles_per_s": val, "n_gpus": 0,
            "steps": k, "warmup": min(args.warmup, 2), "ms_per_step": 1e3 * dt / k,
            "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "gdof_per_s": gd,
            "config": {"workload": workload_name(n, L), "n": n, "levels": L, "nu": [1, 1], "batch": 1, "note": note},
            "cpu_baseline": {"value": gd, "unit": "GDOF/s", "v_cycles_per_s": val, "cores": cores, "kind": kind,
                             "sample": f"{k} V-cycles (+ residual norm each) of the same {n + 1}^2 problem; {what}, "
                                       "torch threads = all host cores"},
            "e2e": {"value": gd, "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_name(n, L):
    """the SAME string in both arms (the driver compares them)"""
    return f"iso Poisson {n + 1}x{n + 1}, V(1,1), {L} levels, single RHS, f=0 model problem"


def cpu_baseline_sample(n, L, seconds=20.0):
    """bounded sample of the same workload on the host cores (rank 0, N=1 only), run in a child process: the reference
    package is also called FEANet, so it cannot share a process with the product"""
    code = ("import json,os,sys,numpy as np,torch;sys.path.insert(0,%r);import bench;"
            "torch.set_num_threads(os.cpu_count());"
            "k,dt,kind,what=bench._reference_cycles(%d,50,2,%f);"
            "print('CPUBASE'+json.dumps([k,dt,kind,what]))" % (ROOT, n, seconds))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    for ln in out.stdout.splitlines():
        if ln.startswith("CPUBASE"):
            k, dt, kind, what = json.loads(ln[7:])
            cores = os.cpu_count()
            return {"value": k / dt * (n + 1) ** 2 / 1e9, "unit": "GDOF/s", "v_cycles_per_s": k / dt, "cores": cores,
                    "kind": kind, "sample": f"{k} V-cycles (+ residual norm) at {n + 1}^2 after 1 warm-up, torch CPU "
                                            f"threads={cores}; {what}"}
    return {"value": None, "unit": "GDOF/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "failed: " + (out.stderr.strip().splitlines() or ["?"])[-1][:200]}


# ---------------------------------------------------------------------------------------------------------
def _replay(eng, k, zero_ctl):
    """k graph-replayed cycles; the device-side control block is re-armed before its history capacity runs out"""
    done = 0
    while done < k:
        eng.ctl.copy_(zero_ctl, non_blocking=True)
        c = min(k - done, eng.max_cycles - 1)
        for _ in range(c):
            eng._graph.replay()
        done += c


def time_engine(eng, steps, warm=3, repeats=REPEATS, reset=None):
    """median over `repeats` of CUDA-event timed blocks of `steps` graph-replayed cycles (ms per cycle, all blocks);
    `reset` (untimed) restores the initial iterate before every block: configurations on which the reference algorithm
    diverges would otherwise overflow, and a solve that has gone to inf / NaN stops doing work (device-side guard)"""
    import torch

    eng.refresh()
    eng._ctl_reset(0, -1.0, eng.max_cycles)  # eps2 < 0: never converge
    eng._ensure_graph()
    zero = eng.ctl.clone()
    _replay(eng, warm, zero)
    torch.cuda.synchronize()
    out = []
    for _ in range(repeats):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if reset is not None:
            reset()
        eng.ctl.copy_(zero, non_blocking=True)
        a.record()
        _replay(eng, steps, zero)
        b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b) / steps)
    return float(np.median(out)), out


def config_entry(name, eng, n, L, B, key_bytes, hist, steps=20, reset=None):
    peak, _ = hbm_peak()
    if reset is not None:
        reset()
    ms, runs = time_engine(eng, steps, reset=reset)
    balg = algorithmic_bytes_per_cycle(n, L, B=B, key_bytes=key_bytes)
    dof = (n + 1) ** 2 * B
    return {"config": name, "ms_per_cycle": ms, "ms_runs": runs, "v_cycles_per_s": 1e3 / ms,
            "gdof_per_s": dof / ms / 1e6, "algorithmic_GB": balg / 1e9,
            "cycle_roofline_frac": balg / (ms * 1e-3) / 1e9 / peak, "residual_history_rel": hist}


def secondary_configs(which):
    """BASELINE.json configs 2, 3 and the one-GPU legs of 4 and 5 at full size: ms/cycle (median of 5 x 20 graph-replayed
    cycles), GDOF/s, fraction of the algorithmic cycle roofline, relative residual history of 8 cycles"""
    import torch

    import mgfea
    from FEANet.drivers import HNet, Multigrid, SingleGrid, _InterfaceSingleGrid
    from FEANet.solver import LINEAR_4, VCycleEngine

    out = {}
    if "cfg2" in which:  # isotropic Poisson 1025^2, 8-level V-cycle, batch of 64 random right-hand sides
        n, L, B = 1024, 8, 64
        jacs = [SingleGrid(2, n // 2 ** l).jac for l in range(L)]
        eng = VCycleEngine(jacs, B=B, conv_rule=mgfea.CONV_MAX)
        g = torch.Generator(device="cuda").manual_seed(0)
        F = torch.randn(B, 1, n + 1, n + 1, generator=g, device="cuda")
        eng.set_u(torch.zeros(B, 1, n + 1, n + 1, device="cuda"))
        eng.set_f(SingleGrid(2, n).fnet(F))
        r0 = torch.sqrt(eng.residual_sumsq()).cpu().numpy()
        hist = [float(np.max(x / r0)) for x in eng.run(n_iter=8)]
        out["cfg2"] = config_entry("config 2: iso Poisson 1025^2 x 64 RHS (randn seed 0 through FNet), 8 levels, V(1,1); "
                                   "history = max over samples", eng, n, L, B, 0, hist)
        del eng, F
        torch.cuda.empty_cache()
    hw = np.load(os.path.join(ROOT, "tests", "golden", "ops.npz"))["hnet_w"]  # Model/.../iso_poisson_33x33.pth
    for tag, smoother in (("cfg3", "hjac"), ("cfg3_jacobi", "jac")):
        if tag in which:  # two-material circle 1:100, 4097^2, 12 levels, 16-channel linear R/P, w = [4, 1]
            n, L, B = 4096, 12, 1
            grids = [_InterfaceSingleGrid(2, n // 2 ** l, prop=(1, 100), shape=0) for l in range(L)]
            hnet = HNet(3)
            hnet.load_state_dict({f"convLayers.{i}.weight": torch.from_numpy(hw[i]).reshape(1, 1, 3, 3) for i in range(3)})
            R16 = np.repeat((LINEAR_4 / np.float32(4.0)).reshape(1, 9), 16, 0)
            P4 = np.repeat(LINEAR_4.reshape(1, 9), 16, 0)
            eng = VCycleEngine([g.jac for g in grids], B=B, smoother=smoother, hnet=hnet, prolong="table", rtab=R16,
                               r_scale=4.0, ptab=P4, p_scale=1.0)
            eng.set_u(torch.zeros(1, 1, n + 1, n + 1, device="cuda"))
            eng.set_f(grids[0].fnet(torch.ones(1, 1, n + 1, n + 1, device="cuda")))
            r0 = float(torch.sqrt(eng.residual_sumsq().sum()).item())
            hist = [x / r0 for x in eng.run(n_iter=8)]
            out[tag] = config_entry(f"config 3: two-phase circle 1:100, 4097^2, 12 levels, V(1,1), "
                                    f"{'learned HNet' if smoother == 'hjac' else 'Jacobi'} smoother, 16-ch linear R/P, "
                                    "w=[4,1], F=ones; the reference algorithm itself diverges at this depth (DESIGN 5, "
                                    "tests/golden/bands.json cfg3_*): throughput per cycle", eng, n, L, B, 1, hist,
                                    steps=10, reset=lambda: eng.u[0].zero_())
            del eng, grids
            torch.cuda.empty_cache()
    if "cfg3_single_pattern" in which:
        # config 3's operators (learned HNet smoother, 16-channel linear R/P, w = [4, 1]) on a SINGLE-material mesh, the
        # setting M-FEANet-mg_test.ipynb trains its HNet in: the finest level runs on the register-chained streaming
        # kernels (csrc/mgfea_hstream.cuh); on the two-phase mesh above they are opt-in (slower than the tile programs)
        n, L, B = 4096, 12, 1
        grids = [SingleGrid(2, n // 2 ** l) for l in range(L)]
        hnet = HNet(3)
        hnet.load_state_dict({f"convLayers.{i}.weight": torch.from_numpy(hw[i]).reshape(1, 1, 3, 3) for i in range(3)})
        R16 = np.repeat((LINEAR_4 / np.float32(4.0)).reshape(1, 9), 16, 0)
        P4 = np.repeat(LINEAR_4.reshape(1, 9), 16, 0)
        eng = VCycleEngine([g.jac for g in grids], B=B, smoother="hjac", hnet=hnet, prolong="table", rtab=R16,
                           r_scale=4.0, ptab=P4, p_scale=1.0)
        eng.set_u(torch.zeros(1, 1, n + 1, n + 1, device="cuda"))
        eng.set_f(grids[0].fnet(torch.ones(1, 1, n + 1, n + 1, device="cuda")))
        r0 = float(torch.sqrt(eng.residual_sumsq().sum()).item())
        hist = [x / r0 for x in eng.run(n_iter=8)]
        out["cfg3_single_pattern"] = config_entry(
            "config 3's cycle on one material: iso Poisson 4097^2, 12 levels, V(1,1), learned HNet smoother, 16-ch linear "
            "R/P, w=[4,1], F=ones", eng, n, L, B, 0, hist, steps=10, reset=lambda: eng.u[0].zero_())
        prev = mgfea.set_option("hstream_min_n", 0)  # the same cycle on the tile programs only
        eng2 = VCycleEngine([g.jac for g in grids], B=B, smoother="hjac", hnet=hnet, prolong="table", rtab=R16,
                            r_scale=4.0, ptab=P4, p_scale=1.0)
        eng2.set_u(torch.zeros(1, 1, n + 1, n + 1, device="cuda"))
        eng2.set_f(grids[0].fnet(torch.ones(1, 1, n + 1, n + 1, device="cuda")))
        ms2, _ = time_engine(eng2, 10, reset=lambda: eng2.u[0].zero_())
        mgfea.set_option("hstream_min_n", prev)
        out["cfg3_single_pattern"]["ms_per_cycle_tile_programs_only"] = ms2
        del eng, eng2, grids
        torch.cuda.empty_cache()
    for tag, n in (("cfg4_1gpu", 8192), ("cfg5_1gpu", 16384)):
        if tag in which:  # configs 4 / 5 on ONE GPU (the row-slab legs come from bench.py --gpus N): f = 0 model problem
            out[tag] = one_gpu_iso(n)
    if "cfg5_two_phase_1gpu" in which:  # config 5 "heterogeneous conductivity": two-phase circle 1:20 at 16385^2
        n, L = 16384, 14
        grids = [_InterfaceSingleGrid(2, n // 2 ** l, prop=(1, 20), shape=0) for l in range(L)]
        eng = VCycleEngine([g.jac for g in grids], B=1, smoother="jac")
        eng.set_u(torch.zeros(1, 1, n + 1, n + 1, device="cuda"))
        eng.set_f(grids[0].fnet(torch.ones(1, 1, n + 1, n + 1, device="cuda")))
        r0 = float(torch.sqrt(eng.residual_sumsq().sum()).item())
        hist = [x / r0 for x in eng.run(n_iter=8)]
        out["cfg5_two_phase_1gpu"] = config_entry("config 5 on 1 GPU: two-phase circle 1:20, 16385^2, 14 levels, V(1,1) "
                                                  "Jacobi, full weighting + bilinear prolongation, F=ones (the "
                                                  "reference algorithm diverges at this depth)", eng, n, L,
                                                  1, 1, hist, steps=10, reset=lambda: eng.u[0].zero_())
        del eng, grids
        torch.cuda.empty_cache()
    if "cfg5_element_1gpu" in which:  # config 5 in its general form: one conductivity per ELEMENT
        from FEANet.element import ElementMultigrid

        n, L = 16384, 14
        g = torch.Generator(device="cuda").manual_seed(5)
        a = torch.exp(torch.rand((n, n), generator=g, device="cuda") * np.log(25.0) + np.log(0.2))  # 0.2 .. 5, log-uniform
        mg = ElementMultigrid(n, a)
        del a
        N = n + 1
        mg.u[0].zero_()
        mg.f[0].view.copy_(mg.grids[0].fnet(torch.ones(1, 1, N, N, device="cuda")))
        r0 = float(torch.sqrt(mg.residual_sumsq().sum()).item())
        hist = []
        for _ in range(6):
            hist.append(float(torch.sqrt(mg.cycle().sum()).item()) / r0)
        runs = []
        for _ in range(REPEATS):
            mg.u[0].zero_()
            torch.cuda.synchronize()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            for _ in range(5):
                mg.cycle(want_norm=True)
            eb.record()
            torch.cuda.synchronize()
            runs.append(ea.elapsed_time(eb) / 5)
        ms = float(np.median(runs))
        balg = algorithmic_bytes_per_cycle(n, L, key_bytes=4)  # + 4 B per node per K-apply for the element field
        peak, _ = hbm_peak()
        out["cfg5_element_1gpu"] = {
            "config": "config 5, general heterogeneous conductivity: 16385^2, one log-uniform conductivity in [0.2, 5] per "
                      "ELEMENT, 14 levels (4-child mean coarsening), V(1,1) Jacobi, F=ones; per-element kernels "
                      "(mgfea_elem_*, one pass per operator, not fused yet), eager launches incl. the residual norm",
            "ms_per_cycle": ms, "ms_runs": runs, "v_cycles_per_s": 1e3 / ms, "gdof_per_s": N * N / ms / 1e6,
            "algorithmic_GB": balg / 1e9, "cycle_roofline_frac": balg / (ms * 1e-3) / 1e9 / peak,
            "residual_history_rel": hist}
        del mg
        torch.cuda.empty_cache()
    return out


def one_gpu_iso(n, steps=20):
    """the iso f=0 model problem of size n on the current GPU (single-GPU engine): the same-problem baseline of the
    row-slab runs"""
    import torch

    from FEANet.drivers import Multigrid

    L = int(np.log2(n))
    prob = Multigrid(n)
    eng = prob._engine(1, 1, 0, B=1)
    g = torch.Generator(device="cuda").manual_seed(123)
    eng.set_u(1.2e5 * torch.rand((1, 1, n + 1, n + 1), generator=g, device="cuda") + 1.3e5)
    eng.set_f(torch.zeros(1, 1, n + 1, n + 1, device="cuda"))
    r0 = float(torch.sqrt(eng.residual_sumsq().sum()).item())
    hist = [x / r0 for x in eng.run(n_iter=8)]
    e = config_entry(f"iso Poisson {n + 1}^2, {L} levels, V(1,1), single RHS, f=0 model problem, 1 GPU", eng, n, L, 1, 0,
                     hist, steps=steps)
    del eng, prob
    torch.cuda.empty_cache()
    return e


# ---------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import mgfea
    from FEANet.drivers import Multigrid

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, L = args.n, int(np.log2(args.n))
    N = n + 1
    dof = N * N
    steps, warm = args.steps, max(args.warmup, 3)

    prob = Multigrid(n)  # every rank: one independent replica of the workload (see DESIGN.md "Multi-GPU")
    u0_host = torch.from_numpy(model_u0(n, seed=123 + rank)).reshape(1, 1, N, N).pin_memory()
    f_host = torch.zeros(1, 1, N, N).pin_memory()
    prob.initial_v = u0_host
    prob.grids[0].f = f_host
    eng = prob._engine(1, 1, 0, B=1)
    eng.set_u(u0_host)
    eng.set_f(f_host)
    eng.refresh()
    eng._ctl_reset(0, -1.0, eng.max_cycles)  # eps2 < 0: never converge; the history ring is capped at max_cycles
    eng._ensure_graph()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ctl.cycle would hit max_cycles and set done; re-arm it on the device without a host sync
    zero_ctl = eng.ctl.clone()

    def rearm():
        eng.ctl.copy_(zero_ctl, non_blocking=True)

    sampler = ClockSampler(local) if rank == 0 else None
    # untimed load phase so that the clock samples are taken under this workload, doubles as warm-up
    t_end = time.time() + 1.0
    while time.time() < t_end:
        rearm()
        for _ in range(64):
            eng._graph.replay()
        torch.cuda.synchronize()
    rearm()
    for _ in range(warm):
        eng._graph.replay()
    barrier()
    ms_runs = []
    for _ in range(REPEATS):  # median of REPEATS blocks of exactly `steps` cycles
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rearm()
        ev0.record()
        done = 0
        while done < steps:
            k = min(steps - done, eng.max_cycles - 1)
            for _ in range(k):
                eng._graph.replay()
            done += k
            if done < steps:
                rearm()
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_runs.append(float(t.item()))
    ms = float(np.median(ms_runs))
    # launches per cycle: count one eager cycle (graph replays do not pass through the host-side counter)
    c0 = mgfea.launch_count()
    rearm()
    eng.cycle(use_ctl=True)
    torch.cuda.synchronize()
    per_cycle = mgfea.launch_count() - c0
    # keep the GPU under the same load a little longer for the sampler
    t_end = time.time() + 0.3
    while time.time() < t_end:
        rearm()
        for _ in range(64):
            eng._graph.replay()
        torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None

    cycles_per_s = world * steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (level-0 fused kernels), CUDA events on the launching stream
    import ctypes

    peak, peak_src = hbm_peak()
    g0 = eng._grids[0]
    rt = eng._keep[0]
    M0, M1 = dof, (n // 2 + 1) ** 2
    reps = 30

    def time_kernel(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    def down():
        mgfea.check(mgfea.lib().mgfea_smooth_residual_restrict(
            ctypes.byref(g0), eng.u[0].ptr, eng.u_alt[0].ptr, eng.f[0].ptr, 1, 0, None, 0, eng.f[1].ptr,
            eng.f[1].pitch, eng.f[1].plane, rt.data_ptr(), 1, 1, 4.0, None, 1, mgfea.stream_ptr()))

    def up():
        g1 = eng._grids[1]
        mgfea.check(mgfea.lib().mgfea_prolong_correct_smooth(
            ctypes.byref(g0), ctypes.byref(g1), eng.u[1].ptr, eng.u_alt[0].ptr, eng.u[0].ptr, eng.f[0].ptr,
            mgfea.PROLONG_BILINEAR, None, 0, 0, 0.0, None, 1, 0, None, 0, 1, mgfea.stream_ptr()))

    def up_norm():  # what the cycle actually launches last: the up leg with the interior residual norm fused in
        g1 = eng._grids[1]
        mgfea.check(mgfea.lib().mgfea_prolong_correct_smooth_norm(
            ctypes.byref(g0), ctypes.byref(g1), eng.u[1].ptr, eng.u_alt[0].ptr, eng.u[0].ptr, eng.f[0].ptr,
            mgfea.PROLONG_BILINEAR, None, 0, 0, 0.0, None, 1, 0, None, 0, eng.sumsq.data_ptr(), 1, mgfea.stream_ptr()))

    ms_down, ms_up, ms_upn = time_kernel(down), time_kernel(up), time_kernel(up_norm)
    alg_down = 4 * (3 * M0 + 2 * M0 + M1)       # pre-smooth (u,f -> u) + residual->restrict (u,f -> f_c)
    alg_up = 4 * (2 * M0 + M1 + 3 * M0)         # prolong->correct (v_c,u -> u) + post-smooth (u,f -> u)
    alg_upn = alg_up + 4 * 2 * M0               # + convergence norm (u,f)
    # the dominant kernel = the longest launch of the cycle
    kname, kms, kalg = "mg_stream2_kernel<1, 0, 0, 0, 1> level-0 up leg (prolong+correct+smooth+residual norm)", ms_upn, alg_upn
    traffic = ncu_traffic("mg_stream2_kernel<1, 0", 280) if n == 4096 else None
    ach = kalg / (kms * 1e-3) / 1e9
    balg = algorithmic_bytes_per_cycle(n, L)
    cyc_ms = ms / steps
    cyc_ach = balg / (cyc_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "frac_actual": (traffic / (kms * 1e-3) / 1e9 / peak) if traffic else None,
                "traffic_source": f"from profile: profiles/{NCU_PROFILE} (ncu --set full of the same workload; not "
                                  "measured in this run)" if traffic else None,
                "note": "achieved = algorithmic bytes (SURVEY 8d: every logical operator reads its inputs and writes its "
                        "outputs once) / measured time; the fused kernel moves only `traffic` bytes through DRAM, so "
                        "achieved may exceed the copy-bandwidth peak; traffic / time is the DRAM rate actually sustained",
                "dram_rate": (traffic / (kms * 1e-3) / 1e9) if traffic else None,
                "kernel": kname, "kernel_ms": kms, "algorithmic_bytes_per_launch": kalg, "peak_source": peak_src,
                "other_kernels": {"down_leg": {"ms": ms_down, "algorithmic_bytes": alg_down,
                                               "achieved": alg_down / (ms_down * 1e-3) / 1e9,
                                               "traffic": ncu_traffic("mg_stream2_kernel<0, 0", 280) if n == 4096 else None},
                                  "up_leg_without_norm": {"ms": ms_up, "algorithmic_bytes": alg_up,
                                                          "achieved": alg_up / (ms_up * 1e-3) / 1e9}},
                "cycle": {"algorithmic_bytes": balg, "ms": cyc_ms, "achieved": cyc_ach, "frac": cyc_ach / peak}}

    # ---- end to end through the reference-facing API: Multigrid.Solve from HOST buffers to 1e-8 relative
    def e2e_once():
        prob.initial_v = u0_host
        prob.grids[0].f = f_host
        r0 = None
        t0 = time.perf_counter()
        res = prob.Solve([1, 1], n_iter=args.e2e_cycles, chunk=args.e2e_cycles)
        u_host = prob.grids[0].v  # Solve returns the solution on the host for host inputs (D2H inside)
        dt = time.perf_counter() - t0
        return dt, res, u_host

    for _ in range(3):  # warm-up: graph, pinned result blocks of the caching host allocator (two alternate)
        e2e_once()
    barrier()
    e2e_runs = []
    for _ in range(5):
        dt, res, _ = e2e_once()
        e2e_runs.append(dt)
    torch.cuda.synchronize()
    e2e_dt = float(np.median(e2e_runs))
    te = torch.tensor([e2e_dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_dt = float(te.item())
    e2e_val = world * args.e2e_cycles / e2e_dt

    # ---- time to 1e-8 relative residual (device resident, f = 0 model problem)
    eng.set_u(u0_host)
    eng.set_f(f_host)
    torch.cuda.synchronize()
    r0 = float(torch.sqrt(eng.residual_sumsq().sum()).item())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    # convergence is evaluated on the device and cycles past it are no-ops, so the host may enqueue a generous chunk
    # and synchronise once (the reference synchronises with .item() after every cycle)
    # n_iter=1: the reference's loop starts from res = 1, so an absolute EPS >= 1 (this u0 is O(1e5): r0 ~ 1e9) would run
    # no cycle at all; at least one cycle, then until res <= EPS
    hist = eng.run(n_iter=1, EPS=1e-8 * r0, chunk=16)
    torch.cuda.synchronize()
    t_tol = time.perf_counter() - t0

    # ---- time to 1e-8 relative residual with a NONZERO right-hand side (fp32 alone stalls near 1e-6, BASELINE.md
    # section 2): fp64 defect correction around the fp32 cycle (Multigrid.SolveMixed), device resident
    mixed = None
    if not args.no_mixed:
        g = torch.Generator(device="cuda").manual_seed(0)
        Fr = torch.randn(1, 1, N, N, generator=g, device="cuda")
        prob.grids[0].f = prob.grids[0].fnet(Fr)
        prob.initial_v = torch.zeros(1, 1, N, N, device="cuda")
        prob.SolveMixed([1, 1], n_iter=2)  # builds the engine + graph
        engm = prob._mixed_engine
        r0m = float(torch.sqrt(engm.r0_sumsq.sum()).item())
        tms = []
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            hm = prob.SolveMixed([1, 1], EPS=1e-8 * r0m, chunk=4)
            torch.cuda.synchronize()
            tms.append(time.perf_counter() - t0)
        tm = float(np.median(tms))
        mixed = {"cycles_to_1e-8_rel": len(hm), "ms": 1e3 * tm, "ms_per_cycle": 1e3 * tm / max(len(hm), 1),
                 "ms_runs": [1e3 * t for t in tms],
                 "final_rel_residual": hm[-1] / r0m,
                 "what": "randn right-hand side (seed 0) through FNet, u0 = 0; iterate / residual in fp64, V-cycle in "
                         "fp32; includes the H2D-free setup copies of u0 and f into the fp64 buffers"}
        prob.grids[0].f = f_host
        prob.initial_v = u0_host

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    configs = None
    if world == 1 and not args.no_configs:
        del eng
        prob._engines.clear()
        torch.cuda.empty_cache()
        configs = secondary_configs(["cfg2", "cfg3", "cfg3_jacobi", "cfg3_single_pattern", "cfg4_1gpu", "cfg5_1gpu", "cfg5_two_phase_1gpu", "cfg5_element_1gpu"]
                                    if not args.configs else args.configs.split(","))
    cpu = cpu_baseline_sample(n, L) if (world == 1 and not args.no_cpu_baseline) else None
    line = {"metric": METRIC, "value": cycles_per_s * dof / 1e9, "unit": "GDOF/s", "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "v_cycles_per_s": cycles_per_s, "repeats": REPEATS, "ms_per_step_runs": [m / steps for m in ms_runs],
            "time_to_1e-8_rel_ms": 1e3 * t_tol, "cycles_to_1e-8_rel": len(hist), "mixed_precision_solve": mixed,
            "config": {"workload": workload_name(n, L), "n": n, "levels": L, "nu": [1, 1], "batch": 1,
                       "what": "MM_Model_convergence.ipynb cell 3 problem; interior residual norm fused in every cycle; "
                               "value = median of 5 timed blocks of `steps` graph-replayed cycles",
                       "l2": "inputs larger than L2 (level-0 u, u', f = 201 MB > 126 MB); no explicit flush",
                       "replicas": world, "loader": "tma" if args.loader == "tma" else "cp.async"},
            "clocks": clocks,
            "e2e": {"value": e2e_val * dof / 1e9, "unit": "GDOF/s", "v_cycles_per_s": e2e_val,
                    "h2d_bytes_per_step": int(2 * 4 * dof / args.e2e_cycles),
                    "d2h_bytes_per_step": int((4 * dof + 8 * args.e2e_cycles) / args.e2e_cycles),
                    "what": f"Multigrid.Solve(n_iter={args.e2e_cycles}) from pinned host u0,f: H2D of both fields, "
                            f"{args.e2e_cycles} cycles, D2H of the residual history and of the solution; "
                            "bytes are per V-cycle; median of 5 solves", "ms_per_solve": 1e3 * e2e_dt,
                    "ms_runs": [1e3 * t for t in e2e_runs]},
            "gpu_launches": int(per_cycle * steps * REPEATS), "gpu_launches_per_step": int(per_cycle),
            "roofline": roofline, "configs": configs}
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def mg_r0(mg):
    """initial fp64 residual norm of the slab problem (collective)"""
    mg.SolveMixed(n_iter=0, EPS=float("inf"))
    return mg.r0


def slab_setup(n, prop, rank, use_graph, **kw):
    """collective: the f=0 model problem of size n split into row slabs, ghost rows exchanged, cycle graph captured"""
    import torch
    import torch.distributed as dist

    from FEANet.distributed import SlabMultigrid

    mg = SlabMultigrid(n, prop=prop, **kw)

    def u_rows(row0, nrows, NN):  # same random family as the reference's model problem, generated per rank on the device
        g = torch.Generator(device="cuda").manual_seed(123 + rank)
        return (1.2e5 * torch.rand((nrows, NN), generator=g, device="cuda") + 1.3e5)

    mg.fill_local(u_rows)
    mg.exchange_initial()
    graphed = mg.enable_graph() if use_graph else False
    gflag = torch.tensor([1.0 if graphed else 0.0], device="cuda")
    dist.all_reduce(gflag, op=dist.ReduceOp.MIN)
    if float(gflag.item()) < 0.5 and graphed:  # all ranks or none
        mg._graph = None
        graphed = False
    return mg, graphed


def time_slab(mg, steps, warm, repeats=REPEATS):
    """collective: median over `repeats` blocks of `steps` cycles, each block timed with CUDA events on every rank between
    barriers, max over ranks"""
    import torch
    import torch.distributed as dist

    for _ in range(warm):
        mg.cycle()
    runs = []
    for _ in range(repeats):
        dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            mg.cycle()
        ev1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        runs.append(float(t.item()) / steps)
    return float(np.median(runs)), runs


def slab_parity(world, rank, n=2048, cycles=3):
    """REAL multi-GPU parity self-check (one GPU per rank, concurrent kernels, peer stores over NVLink): `cycles` V-cycles
    of a seeded n^2 problem with three distributed levels on the N ranks against the same solve on rank 0 alone.
    Bit-identical solution and residual history, or the run fails."""
    import torch
    import torch.distributed as dist

    from FEANet.distributed import SlabMultigrid
    from FEANet.drivers import Multigrid

    rs = np.random.RandomState(3)
    u0 = rs.standard_normal((n + 1, n + 1)).astype(np.float32)
    f = 0.01 * rs.standard_normal((n + 1, n + 1)).astype(np.float32)
    mg = SlabMultigrid(n, dist_min_n=513)
    mg.set_problem(torch.from_numpy(u0), torch.from_numpy(f))
    hist = mg.Solve(n_iter=cycles)
    sol = mg.gather_solution().numpy()
    out = {"n": n, "cycles": cycles, "ranks": world, "distributed_levels": mg.part.ld,
           "exchange": "peer" if mg.peer is not None else "nccl"}
    mg.close()
    ok = torch.ones(1, device="cuda")
    if rank == 0:
        prob = Multigrid(n)
        prob.initial_v = torch.from_numpy(u0)
        prob.grids[0].f = torch.from_numpy(f).reshape(1, 1, n + 1, n + 1)
        ref = prob.Solve([1, 1], n_iter=cycles)
        want = prob.grids[0].v.numpy()[0, 0]
        out["bit_identical"] = bool(np.array_equal(sol, want))
        out["max_abs_diff"] = float(np.abs(sol.astype(np.float64) - want).max())
        out["hist_rel"] = float(np.max(np.abs(np.array(hist) - np.array(ref)) / np.array(ref)))
        out["history"] = hist
        if not out["bit_identical"] or out["hist_rel"] > 1e-12:
            ok.zero_()
        del prob
        torch.cuda.empty_cache()
    dist.broadcast(ok, 0)
    out["ok"] = bool(ok.item() > 0.5)
    return out


def run_multi(args):
    """N > 1: ONE problem partitioned into row slabs (FEANet.distributed): peer-store halo exchange over NVLink,
    replicated coarse levels.  Strong scaling: the line carries the same problem timed on ONE GPU in the same run."""
    import torch
    import torch.distributed as dist

    import mgfea

    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa = mgfea.bind_to_gpu_numa(local)  # pinned host buffers of the e2e leg land on the GPU's own NUMA node
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = size_for(world, args.n_multi)
    N = n + 1
    dof = N * N
    L = int(np.log2(n))
    steps, warm = args.steps, max(args.warmup, 3)
    prop = tuple(float(x) for x in args.prop.split(",")) if args.prop else None
    peak, peak_src = hbm_peak()

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    mg, graphed = slab_setup(n, prop, rank, not args.no_graph)  # --prop a,b: two-phase circle (keyed streaming kernels)
    lev = mg.part.levels[0]
    sampler = ClockSampler(local) if rank == 0 else None
    # every rank must issue the SAME sequence of exchange steps: agree on the number of load-phase cycles first
    for _ in range(3):
        mg.cycle()
    barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        mg.cycle()
    torch.cuda.synchronize()
    tc = torch.tensor([(time.perf_counter() - t0) / 5], dtype=torch.float64, device="cuda")
    dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    n_load = int(max(5, min(2000, 1.0 / max(float(tc.item()), 1e-5))))
    for _ in range(n_load):
        mg.cycle()
    ms_cycle, ms_runs = time_slab(mg, steps, warm)
    c0 = mgfea.launch_count()  # graph replays bypass the host-side counter: count one eager cycle on every rank
    mg._cycle_eager()
    barrier()
    launches = (mgfea.launch_count() - c0) * steps * REPEATS
    for _ in range(max(5, n_load // 3)):
        mg.cycle()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    cycles_per_s = 1e3 / ms_cycle
    if mg.peer is not None:
        mg.peer.check()  # a timed-out peer wait invalidates the run: fail loudly

    # dominant kernel on this rank: level-0 slab down leg
    ops = mg.ops
    reps = 20

    def down():
        ops.down(0, mg.u[0], mg.u_alt[0], mg.f[0], mg.f[1] if mg.part.ld > 1 else ops.coarse_f())

    for _ in range(3):
        down()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        down()
    b.record()
    torch.cuda.synchronize()
    kms = a.elapsed_time(b) / reps
    mloc = (lev["own1"] - lev["own0"]) * N
    kalg = 4 * (5 * mloc + mloc // 4)
    ach = kalg / (kms * 1e-3) / 1e9
    balg = algorithmic_bytes_per_cycle(n, L)

    # end to end: local rows from pinned host memory, e2e_cycles cycles (device-side stopping rule), D2H of owned rows
    hu = torch.empty((lev["nrows"], N), dtype=torch.float32).pin_memory()
    hu.copy_(mg.u[0][0, :, :N])
    hf = torch.zeros((lev["nrows"], N), dtype=torch.float32).pin_memory()
    hout = torch.empty((lev["own1"] - lev["own0"], N), dtype=torch.float32).pin_memory()

    def e2e_once():
        mg.u[0][0, :, :N].copy_(hu, non_blocking=True)
        mg.f[0][0, :, :N].copy_(hf, non_blocking=True)
        hist = mg.Solve(n_iter=args.e2e_cycles)
        hout.copy_(mg.u[0][0, lev["own0"] - lev["row0"]: lev["own1"] - lev["row0"], :N], non_blocking=True)
        torch.cuda.synchronize()
        return hist

    e2e_once()
    e2e_runs = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        e2e_once()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_runs.append(float(te.item()))
    e2e_dt = float(np.median(e2e_runs))
    # ---- time to 1e-8 relative residual, nonzero right-hand side: fp64 defect correction on the slabs (device resident)
    mixed = None
    if mg.peer is not None and not args.no_mixed:
        def f_rows(row0, nrows, NN):
            g = torch.Generator(device="cuda").manual_seed(777 + rank)
            return 1e-3 * torch.randn((nrows, NN), generator=g, device="cuda", dtype=torch.float64)

        mg.fill_local64(f_rows)
        mg.SolveMixed(n_iter=3)  # warm-up: first step eager, then the step graph is captured and replayed
        mg.fill_local64(f_rows)
        barrier()
        t0 = time.perf_counter()
        hm = mg.SolveMixed(EPS=1e-8 * mg_r0(mg), max_cycles=60)
        torch.cuda.synchronize()
        tm = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        mixed = {"cycles_to_1e-8_rel": len(hm), "ms": 1e3 * float(tm.item()), "final_rel_residual": hm[-1] / mg.r0,
                 "what": "1e-3 * randn right-hand side, u0 = 0; iterate / residual fp64 on the slabs, V-cycle fp32; "
                         "host checks the all-reduced norm every cycle"}
    exchange, ld, graph_err, peer_err = ("peer" if mg.peer is not None else "nccl"), mg.part.ld, mg._graph_err, \
        getattr(mg, "peer_error", None)
    mg.close()

    # ---- strong scaling: the SAME problem on ONE GPU (rank 0, same run), then the other BASELINE size both ways
    def same_problem(nn, slab_ms):
        e = None
        if rank == 0 and not prop:
            e = one_gpu_iso(nn)
        barrier()
        if e is None:
            return None
        return {"n": nn, "one_gpu_ms_per_cycle": e["ms_per_cycle"], "one_gpu_gdof_per_s": e["gdof_per_s"],
                "one_gpu_cycle_roofline_frac": e["cycle_roofline_frac"], "n_gpu_ms_per_cycle": slab_ms,
                "speedup": e["ms_per_cycle"] / slab_ms, "efficiency": e["ms_per_cycle"] / slab_ms / world}

    strong = same_problem(n, ms_cycle)
    configs = {}
    if not args.no_configs and not prop:
        n2 = 8192 if n != 8192 else 16384
        mg2, g2 = slab_setup(n2, None, rank, not args.no_graph)
        ms2, runs2 = time_slab(mg2, max(10, steps // 2), warm)
        if mg2.peer is not None:
            mg2.peer.check()
        mg2.close()
        st2 = same_problem(n2, ms2)
        configs[f"cfg{'4' if n2 == 8192 else '5'}_{world}gpu"] = {
            "config": f"iso Poisson {n2 + 1}^2, {int(np.log2(n2))} levels, V(1,1), f=0 model problem, {world} row slabs",
            "ms_per_cycle": ms2, "ms_runs": runs2, "gdof_per_s": (n2 + 1) ** 2 / ms2 / 1e6, "cuda_graph": bool(g2),
            "cycle_roofline_frac": algorithmic_bytes_per_cycle(n2, int(np.log2(n2))) / (ms2 * 1e-3) / 1e9 / (peak * world),
            "strong_scaling": st2}
    parity = None if args.no_parity else slab_parity(world, rank)

    xname = ("halo rows stored straight into the neighbours' ghost rows over NVLink peer memory, one exchange kernel "
             "per step, no NCCL on the data path") if exchange == "peer" else "NCCL send/recv halo exchange"
    if rank == 0:
        line = {"metric": METRIC, "value": cycles_per_s * dof / 1e9, "unit": "GDOF/s", "n_gpus": world, "steps": steps,
                "warmup": warm, "ms_per_step": ms_cycle, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "v_cycles_per_s": cycles_per_s,
                "repeats": REPEATS, "ms_per_step_runs": ms_runs, "mixed_precision_solve": mixed,
                "strong_scaling": strong, "slab_parity": parity, "configs": configs,
                "config": {"workload": workload_name(n, L) if not prop else
                           f"two-phase circle {args.prop} Poisson {N}x{N}, V(1,1), {L} levels, single RHS, F=0",
                           "partition": f"{world} row slabs ({xname}, levels N<2049 replicated); "
                                        f"{dof / world / 1e6:.1f} MDOF per GPU; value = median of 5 timed blocks",
                           "n": n, "levels": L, "nu": [1, 1], "batch": 1, "first_replicated_level": ld,
                           "cuda_graph": bool(graphed), "graph_error": graph_err, "exchange": exchange,
                           "peer_error": peer_err, "numa_node_rank0": numa,
                           "l2": "inputs larger than L2 per GPU at level 0; no explicit flush"},
                "clocks": clocks,
                "e2e": {"value": args.e2e_cycles * dof / e2e_dt / 1e9, "unit": "GDOF/s",
                        "h2d_bytes_per_step": int(2 * 4 * dof / args.e2e_cycles),
                        "d2h_bytes_per_step": int(4 * dof / args.e2e_cycles),
                        "what": f"SlabMultigrid.Solve(n_iter={args.e2e_cycles}) from pinned host slabs on every rank, "
                                "D2H of the owned rows; bytes are whole-job per V-cycle; median of 3 solves",
                        "ms_per_solve": 1e3 * e2e_dt, "ms_runs": [1e3 * t for t in e2e_runs]},
                "gpu_launches": int(launches), "gpu_launches_per_step": int(launches // (steps * REPEATS)),
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "traffic": None, "kernel": "mg_stream2_kernel<down> level-0 slab (rank 0)",
                             "kernel_ms": kms, "algorithmic_bytes_per_launch": kalg, "peak_source": peak_src,
                             "cycle": {"algorithmic_bytes": balg, "ms": ms_cycle,
                                       "achieved": balg / (ms_cycle * 1e-3) / 1e9,
                                       "frac": balg / (ms_cycle * 1e-3) / 1e9 / (peak * world)}}}
        print(json.dumps(line), flush=True)
    # teardown: leave without running the process-group destructor (it can block on graph-captured communicators)
    barrier()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0 if (parity is None or parity["ok"]) else 3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--e2e-cycles", type=int, default=13)
    ap.add_argument("--loader", default="tma", choices=["tma", "cpasync"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--prop", default="", help="N > 1 only: conductivities a,b of a two-phase circle inclusion (e.g. 1,20)")
    ap.add_argument("--no-mixed", action="store_true", help="skip the fp64 defect-correction time-to-tolerance run")
    ap.add_argument("--n-multi", type=int, default=0, help="grid intervals for the row-slab run at N > 1 (default by N)")
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary BASELINE configs block")
    ap.add_argument("--configs", default="", help="comma list out of cfg2,cfg3,cfg3_jacobi,cfg3_single_pattern,cfg4_1gpu,cfg5_1gpu,"
                                                  "cfg5_two_phase_1gpu,cfg5_element_1gpu (default: all at N=1)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the slab-vs-single-GPU bit-identity self-check")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 200:
            args.steps = 10
        run_reference(args)
        return
    import mgfea

    mgfea.set_loader(args.loader == "tma")
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        if args.steps == 200:
            args.steps = 50
        run_multi(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
