"""V-cycle engine shared by the reference-facing drivers (FEANet.drivers, FEANet.multigrid).

Holds the padded per-level buffers in HBM (u ping-pong pair + f per level), the level descriptors handed to the C ABI,
and drives ``mgfea_vcycle``.  The steady-state loop replays a CUDA graph of one cycle (residual norm fused into the last
kernel, convergence evaluated on the device in an ``mgfea_ctl`` block), so the host only synchronises every few cycles
instead of calling ``.item()`` per cycle like the reference does (MM_Model_convergence.ipynb cell 3 ``Solve``).
"""
import ctypes
import math

import numpy as np
import torch

import mgfea
from mgfea import CycleCfg, Ctl, Field, Grid, LevelBufs, check, lib, stream_ptr

FULL_WEIGHTING_16 = (np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=np.float32) / np.float32(16.0))
LINEAR_4 = (np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=np.float32) / np.float32(4.0))


class VCycleEngine:
    """levels: list of FEANet.jacobi.JacobiBlock (finest first); each carries its KNet (weights, pattern keys), its
    omega/d table and its Dirichlet masks."""

    def __init__(self, jacs, B=1, nu1=1, nu2=1, smoother="jac", hnet=None, prolong="bilinear", rtab=None, r_scale=4.0,
                 ptab=None, p_scale=None, w_param=None, quirk_level0=False, conv_rule=mgfea.CONV_SUM,
                 max_cycles=256, f0_store=None, compute_norm=True, zero_guess=False):
        self.dev = mgfea.require_cuda()
        self.jacs = list(jacs)
        self.L = len(self.jacs)
        self.B = B
        self.nu1, self.nu2 = int(nu1), int(nu2)
        self.smoother = smoother
        self.hnet = hnet
        self.prolong = prolong
        self.quirk_level0 = quirk_level0
        self.conv_rule = conv_rule
        self.max_cycles = max_cycles
        self.compute_norm, self.zero_guess = bool(compute_norm), bool(zero_guess)
        self._rtab_src = rtab  # numpy array, torch tensor / Parameter (live) or None (full weighting /16)
        self._ptab_src = ptab
        self.r_scale, self.p_scale = r_scale, p_scale
        self.w_param = w_param  # live MultiGrid.w (2,) -> device scales
        self._tabs = {k: mgfea.DeviceTable() for k in ("r", "p", "w", "h")}
        self.u = [Field(B, j.nnode_edge, self.dev) for j in self.jacs]
        self.u_alt = [Field(B, j.nnode_edge, self.dev) for j in self.jacs]
        self.f = [Field(B, j.nnode_edge, self.dev) for j in self.jacs]
        if f0_store is not None:  # caller-owned level-0 right-hand side (peer-mapped memory of the row-slab path)
            self.f[0] = Field(B, self.jacs[0].nnode_edge, self.dev, store=f0_store)
        self.sumsq = torch.zeros(B, dtype=torch.float64, device=self.dev)
        self.hist = torch.zeros((max_cycles, B), dtype=torch.float64, device=self.dev)
        self.ctl = torch.zeros(8, dtype=torch.int32, device=self.dev)  # mgfea_ctl (32 bytes)
        self._ctl_host = torch.zeros(8, dtype=torch.int32).pin_memory()
        self._graph = None
        self._graph_key = None
        self._keep = []
        self.refresh()
        # first use allocates the library's per-device scratch (must not happen inside a graph capture)
        check(lib().mgfea_residual_norm(ctypes.byref(self._grids[0]), self.u[0].ptr, self.f[0].ptr,
                                        self.sumsq.data_ptr(), None, None, B, stream_ptr()))

    # -- descriptors ---------------------------------------------------------------------------------------
    def _table(self, which, src, default):
        if src is None:
            src = default
        t = src if isinstance(src, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(src, dtype=np.float32))
        t = t.detach().reshape(-1, 9)
        return self._tabs[which].get(t.contiguous() if not t.is_contiguous() else t)

    def refresh(self):
        """(re)build the C descriptors; picks up live weight changes (KNet.net2, R/P kernels, w, HNet layers)"""
        grids = (Grid * self.L)()
        bufs = (LevelBufs * self.L)()
        for l, j in enumerate(self.jacs):
            grids[l] = j.grid_struct(self.u[l])
            bufs[l].u, bufs[l].u_alt, bufs[l].f = self.u[l].ptr, self.u_alt[l].ptr, self.f[l].ptr
        self._grids, self._bufs = grids, bufs
        cfg = CycleCfg()
        cfg.nu1, cfg.nu2 = self.nu1, self.nu2
        cfg.smoother = mgfea.SMOOTH_HJACOBI if self.smoother == "hjac" else mgfea.SMOOTH_JACOBI
        keep = []
        if self.smoother == "hjac":
            ws = [l.weight for l in self.hnet.convLayers]
            sid = tuple((w.data_ptr(), w._version) for w in ws)
            if getattr(self, "_hw_dev", None) is None or self._hw_dev.shape[0] != len(ws):
                self._hw_dev, self._hw_sid = torch.zeros((len(ws), 9), dtype=torch.float32, device=self.dev), None
            if sid != self._hw_sid:  # persistent device buffer: the pointer stays valid for captured graphs
                self._hw_dev.copy_(torch.stack([w.detach().reshape(9).float().cpu() for w in ws]))
                self._hw_sid = sid
            cfg.hw, cfg.nlayers = self._hw_dev.data_ptr(), len(ws)
        rt = self._table("r", self._rtab_src, FULL_WEIGHTING_16)
        keep.append(rt)
        cfg.rtab, cfg.rtab_n = rt.data_ptr(), rt.shape[0]
        cfg.prolong_mode = mgfea.PROLONG_TABLE if self.prolong == "table" else mgfea.PROLONG_BILINEAR
        if self.prolong == "table":
            pt = self._table("p", self._ptab_src, LINEAR_4)
            keep.append(pt)
            cfg.ptab, cfg.ptab_n = pt.data_ptr(), pt.shape[0]
        if self.w_param is not None:
            w = self._tabs["w"].get(self.w_param)
            keep.append(w)
            cfg.r_has_scale = cfg.p_has_scale = 1
            cfg.r_scale_dev, cfg.p_scale_dev = w.data_ptr(), w.data_ptr() + 4
        else:
            cfg.r_has_scale = int(self.r_scale is not None)
            cfg.r_scale_host = float(self.r_scale or 0.0)
            cfg.p_has_scale = int(self.p_scale is not None)
            cfg.p_scale_host = float(self.p_scale or 0.0)
        cfg.quirk_level0 = int(self.quirk_level0)
        cfg.compute_norm = int(self.compute_norm)
        cfg.zero_guess = int(self.zero_guess)
        self._cfg = cfg
        self._keep = keep
        sig = (cfg.hw, cfg.rtab, cfg.ptab, cfg.r_scale_dev, tuple(g.ktab for g in grids), tuple(g.bc_idx for g in grids),
               self.nu1, self.nu2)
        if sig != self._graph_key:
            self._graph, self._graph_key = None, sig
            self._graph64 = None

    # -- problem data --------------------------------------------------------------------------------------
    def set_u(self, u):
        self._load(self.u[0], u)

    def set_f(self, f):
        self._load(self.f[0], f)

    def _load(self, dst: Field, x):
        fld = getattr(x, "_mgfea_field", None)
        if fld is dst:
            return
        x = torch.as_tensor(x)
        if x.dim() == 2:
            x = x[None, None]
        elif x.dim() == 3:
            x = x[:, None]
        if x.shape[0] != dst.B and x.shape[0] == 1:
            x = x.expand(dst.B, -1, -1, -1)
        if tuple(x.shape) != (dst.B, 1, dst.N, dst.N):
            raise mgfea.MgfeaError(f"field shape {tuple(x.shape)} does not match level ({dst.B},1,{dst.N},{dst.N})")
        # strided copy straight into the padded layout (H2D when x lives on the host)
        dst.view.copy_(x.to(dtype=torch.float32), non_blocking=True)

    # -- cycles --------------------------------------------------------------------------------------------
    def _ctl_reset(self, min_cycles, eps2, max_cycles):
        c = Ctl()
        c.cycle, c.done, c.min_cycles, c.max_cycles = 0, 0, int(min_cycles), int(max_cycles)
        c.conv_rule, c.eps2 = self.conv_rule, eps2
        raw = np.frombuffer(bytes(c), dtype=np.int32).copy()
        self.ctl.copy_(torch.from_numpy(raw), non_blocking=False)

    def cycle(self, use_ctl=False):
        """one V-cycle, eager launch sequence (no graph); the interior residual sum of squares lands in self.sumsq"""
        check(lib().mgfea_vcycle(self._grids, self._bufs, self.L, ctypes.byref(self._cfg), self.sumsq.data_ptr(),
                                 self.ctl.data_ptr() if use_ctl else None, self.hist.data_ptr() if use_ctl else None,
                                 self.B, stream_ptr()))

    def residual_sumsq(self):
        """per-sample interior sum of squares of f - K u on level 0 (device tensor, float64)"""
        check(lib().mgfea_residual_norm(ctypes.byref(self._grids[0]), self.u[0].ptr, self.f[0].ptr,
                                        self.sumsq.data_ptr(), None, None, self.B, stream_ptr()))
        return self.sumsq

    def _ensure_graph(self):
        if self._graph is not None:
            return
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        # warm the lazily initialised pieces (function attributes, tensor maps) on a side stream with the done flag set
        # so that no field is modified: every kernel returns immediately when ctl->done != 0
        saved = self.ctl.clone()
        self.ctl[1] = 1
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.cycle(use_ctl=True)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.ctl.copy_(saved)
        with torch.cuda.graph(g):
            self.cycle(use_ctl=True)
        self._graph = g

    def run(self, n_iter=None, EPS=None, max_cycles=None, chunk=4, use_graph=True):
        """Multigrid.Solve loop: repeat cycles while (res > EPS or n < n_iter).  Returns the list of per-cycle interior
        residual 2-norms (whole batch for CONV_SUM; per-sample array rows for CONV_MAX via self.last_hist)."""
        if n_iter is None:
            if EPS is None:
                print("At least one of EPS and n_iter have to be assigned")
                return None
            n_iter = 0
        elif EPS is None:
            EPS = math.inf
        cap = min(max_cycles or self.max_cycles, self.max_cycles)
        if n_iter > cap:
            raise mgfea.MgfeaError(f"n_iter={n_iter} exceeds the history capacity {cap}")
        eps2 = float(EPS) * float(EPS) if math.isfinite(EPS) else 1.7e308
        if n_iter == 0 and EPS >= 1.0:
            # the reference's loop starts from `res = 1` (MM_Model_convergence.ipynb cell 3 Solve): with EPS >= 1 and no
            # n_iter it runs no cycle at all
            self.last_hist = np.zeros((0, self.B))
            return []
        self.refresh()
        self._ctl_reset(n_iter, eps2, cap)
        if use_graph:
            self._ensure_graph()
        done = False
        while not done:
            for _ in range(chunk):
                if use_graph:
                    self._graph.replay()
                else:
                    self.cycle(use_ctl=True)
            self._ctl_host.copy_(self.ctl, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            done = bool(self._ctl_host[1].item())
        ncyc = int(self._ctl_host[0].item())
        h = self.hist[:ncyc].cpu().numpy()
        self.last_hist = h
        self._warn_if_capped(h, ncyc, cap, n_iter, eps2)
        if self.conv_rule == mgfea.CONV_SUM:
            return [float(math.sqrt(v)) for v in h.sum(axis=1)]
        return [np.sqrt(row) for row in h]

    def _warn_if_capped(self, h, ncyc, cap, n_iter, eps2):
        """the reference loops until converged; this engine stops at the history capacity -- say so instead of returning a
        history that silently did not reach EPS"""
        if ncyc >= cap and ncyc > n_iter and eps2 < 1.7e308 and ncyc > 0:
            last = h[-1].sum() if self.conv_rule == mgfea.CONV_SUM else h[-1].max()
            if not last <= eps2:
                import warnings

                warnings.warn(f"mgfea: stopped after {ncyc} cycles (max_cycles) with residual {math.sqrt(last):.3e} > EPS "
                              f"{math.sqrt(eps2):.3e}; raise max_cycles (reference: loops until converged)", RuntimeWarning)

    # -- fp64 defect correction around the fp32 cycle (the reference's remedy is `.double()`) ----------
    def _mixed_setup(self):
        if not (self.zero_guess and not self.compute_norm):
            raise mgfea.MgfeaError("run_mixed needs an engine built with zero_guess=True, compute_norm=False")
        if getattr(self, "u64", None) is None:
            N, pitch = self.u[0].N, self.u[0].pitch
            self.u64 = torch.zeros((self.B, N, pitch), dtype=torch.float64, device=self.dev)
            self.f64 = torch.zeros((self.B, N, pitch), dtype=torch.float64, device=self.dev)
            self._graph64 = None

    def _load64(self, dst, x, zero_ring):
        """(B,1,N,N) problem data -> padded fp64 buffer.  fp32 input (the reference's dtype) goes through the padded fp32
        layout (in place for our own strided views) and ONE widening kernel; fp64 input is copied as is."""
        x = x0 = torch.as_tensor(x)
        while x.dim() < 4:
            x = x[None]
        N = self.u[0].N
        if x.shape[0] != self.B and x.shape[0] == 1:
            x = x.expand(self.B, -1, -1, -1)
        if tuple(x.shape) != (self.B, 1, N, N):
            raise mgfea.MgfeaError(f"field shape {tuple(x.shape)} does not match level ({self.B},1,{N},{N})")
        if x.dtype == torch.float64:
            dst[:, :, :N].copy_(x[:, 0], non_blocking=True)
            if zero_ring:  # the first reset_boundary of the reference's sweep (default Dirichlet ring)
                dst[:, 0, :] = 0
                dst[:, N - 1, :] = 0
                dst[:, :, 0] = 0
                dst[:, :, N - 1:] = 0
            return
        own = x0.dim() == 4 and x0.shape[0] == self.B and x0.dtype == torch.float32  # possibly one of our padded views
        fld = mgfea.as_field(x0 if own else x.to(dtype=torch.float32), self.dev)
        check(lib().mgfea_widen_f64(fld.ptr, dst.data_ptr(), N, fld.pitch, fld.plane, self.B, int(zero_ring), stream_ptr()))

    def _mixed_step(self, use_ctl=True):
        """e = V-cycle(0, r) in fp32; u64 += e; r = f64 - K u64 (fp64, rounded to fp32 into the cycle's rhs) + norm"""
        g0 = ctypes.byref(self._grids[0])
        ctl = self.ctl.data_ptr() if use_ctl else None
        self.cycle(use_ctl=use_ctl)
        check(lib().mgfea_correct_f64(g0, self.u64.data_ptr(), self.u[0].ptr, ctl, self.B, stream_ptr()))
        check(lib().mgfea_defect_f64(g0, self.u64.data_ptr(), self.f64.data_ptr(), self.f[0].ptr, self.sumsq.data_ptr(),
                                     ctl, self.hist.data_ptr() if use_ctl else None, self.B, stream_ptr()))

    def run_mixed(self, u0, f, n_iter=None, EPS=None, max_cycles=None, chunk=4, use_graph=True):
        """Multigrid.Solve loop with the iterate, right-hand side and residual in fp64 and one fp32 V-cycle (from the
        zero guess, on the fp64 residual) per step.  Returns the fp64 interior residual 2-norms after every cycle (the
        history the reference produces after `.double()`); the solution is `self.solution64`."""
        if n_iter is None:
            if EPS is None:
                print("At least one of EPS and n_iter have to be assigned")
                return None
            n_iter = 0
        elif EPS is None:
            EPS = math.inf
        cap = min(max_cycles or self.max_cycles, self.max_cycles)
        if n_iter > cap:
            raise mgfea.MgfeaError(f"n_iter={n_iter} exceeds the history capacity {cap}")
        eps2 = float(EPS) * float(EPS) if math.isfinite(EPS) else 1.7e308
        self._mixed_setup()
        self.refresh()
        if self._grids[0].bc_idx:
            raise mgfea.MgfeaError("run_mixed supports the default Dirichlet ring only")
        self._load64(self.u64, u0, True)
        self._load64(self.f64, f, False)
        self._ctl_reset(n_iter, eps2, cap)
        # r_0 (not part of the history: Solve records the residual AFTER each cycle)
        check(lib().mgfea_defect_f64(ctypes.byref(self._grids[0]), self.u64.data_ptr(), self.f64.data_ptr(), self.f[0].ptr,
                                     self.sumsq.data_ptr(), None, None, self.B, stream_ptr()))
        self.r0_sumsq = self.sumsq.clone()
        if use_graph and self._graph64 is None:
            torch.cuda.synchronize()
            saved = self.ctl.clone()
            self.ctl[1] = 1  # warm-up pass with the done flag set: nothing is modified
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._mixed_step()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self.ctl.copy_(saved)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._mixed_step()
            self._graph64 = g
        done = False
        while not done:
            for _ in range(chunk):
                if use_graph:
                    self._graph64.replay()
                else:
                    self._mixed_step()
            self._ctl_host.copy_(self.ctl, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            done = bool(self._ctl_host[1].item())
        ncyc = int(self._ctl_host[0].item())
        h = self.hist[:ncyc].cpu().numpy()
        self.last_hist = h
        if self.conv_rule == mgfea.CONV_SUM:
            return [float(math.sqrt(v)) for v in h.sum(axis=1)]
        return [np.sqrt(row) for row in h]

    @property
    def solution64(self):
        N = self.u[0].N
        return self.u64[:, None, :, :N]

    @property
    def solution(self):
        return self.u[0].view

    def solution_to_host(self):
        """D2H of the level-0 solution into a PINNED host tensor (contiguous (B,1,N,N)) that belongs to the caller.
        torch's caching host allocator hands the same pinned block back once the previous result has been dropped, so a
        solve loop pays neither a cudaHostAlloc nor a host-side copy per call, and a result the caller keeps is never
        overwritten by a later solve."""
        out = torch.empty((self.B, 1, self.u[0].N, self.u[0].N), dtype=torch.float32, pin_memory=True)
        out.copy_(self.u[0].view, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out
