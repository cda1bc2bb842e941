"""Row-slab multi-GPU V-cycle (the reference has no distributed code at all).

One process per GPU (``torch.distributed``).  The fine levels (N >= ``dist_min_n``) are split into contiguous row slabs,
boundaries aligned so that fine row 2I and coarse row I live on the same rank; every rank keeps its owned rows plus
``GHOST`` ghost rows per side, refreshed by a nearest-neighbour exchange (NCCL send/recv over NVLink) after each kernel
that changes them.  Levels below the threshold are agglomerated: the first replicated right-hand side is all-gathered
and every rank runs the small coarse cycle redundantly (no scatter on the way up).  The interior residual norm is an
all-reduce of one double per sample.

The arithmetic is identical to the single-GPU cycle, so results are bit-identical for any number of ranks
(tests/test_distributed_cpu.py checks that on CPU with gloo, using the oracle as the local operator;
tests/test_gpu_parity.py::test_slab_* emulates the ranks on one GPU).
"""
import ctypes
import math
import os

import numpy as np
import torch
import torch.distributed as dist

GHOST = 4  # margin rows kept beyond the redundantly computed range: the fused legs stream 3 rows above / 2 below it


def halo_depths(ld):
    """Communication-avoiding deep halos.  With `ld` distributed levels a rank computes, besides its owned rows, `a[l]`
    extra rows per side of the down leg of level l and `y[l]` of the up leg REDUNDANTLY (bit-identical to what the
    neighbour computes for the same rows), so that no level needs a halo exchange inside the cycle:

      up leg     y[0] = 0; y[l] = 2: level l-1 prolongs from coarse rows own +- (y[l-1]/2 + 1)
      down leg   the last distributed level only feeds the gathered replicated right-hand side: a[ld-1] = 4 (>= y + 1,
                 the rows of the pre-smoothed iterate its own up leg reads); level l must produce the coarse right-hand
                 side on own_{l+1} +- (a[l+1] + 2) (residual + restriction stencils), i.e. a[l] = 2 a[l+1] + 4

    Per cycle only TWO exchange steps remain: the gather of the first replicated right-hand side, and after the finest
    up leg the push of a[0] + GHOST rows of the new iterate together with the all-reduce of the residual norm.
    Returns (a, y, G): G[l] = a[l] + GHOST ghost rows per side are held in memory."""
    a = [0] * ld
    for l in range(ld - 1, -1, -1):
        a[l] = 4 if l == ld - 1 else 2 * a[l + 1] + 4
    y = [0 if l == 0 else 2 for l in range(ld)]
    return a, y, [v + GHOST for v in a]


class SlabPartition:
    """Host-side description of which global rows of every level a rank owns / holds / computes."""

    def __init__(self, n, L, world, rank, dist_min_n=2049):
        self.n, self.L, self.world, self.rank, self.dist_min_n = n, L, world, rank, dist_min_n
        # number of distributed levels: the size threshold, then as many as leave every rank at least G[l] owned rows
        ld = 0
        if world > 1:
            while ld < L and (n // 2 ** ld + 1) >= dist_min_n and (n // 2 ** ld) % (2 * world) == 0:
                ld += 1
            while ld > 0 and any((n // 2 ** l) // world < g for l, g in enumerate(halo_depths(ld)[2])):
                ld -= 1
        a, y, G = halo_depths(ld)
        self.levels = []
        for l in range(L):
            nl = n // (2 ** l)
            N = nl + 1
            if l < ld:
                per = nl // world
                own0 = rank * per
                own1 = (rank + 1) * per + (1 if rank == world - 1 else 0)
                row0 = max(0, own0 - G[l])
                row1 = min(N, own1 + G[l])
                lev = dict(N=N, dist=True, own0=own0, own1=own1, row0=row0, nrows=row1 - row0, G=G[l],
                           dn0=max(0, own0 - a[l]), dn1=min(N, own1 + a[l]), up0=max(0, own0 - y[l]),
                           up1=min(N, own1 + y[l]))
            else:
                lev = dict(N=N, dist=False, own0=0, own1=N, row0=0, nrows=N, G=0, dn0=0, dn1=N, up0=0, up1=N)
            self.levels.append(lev)
        self.ld = ld  # first replicated level
        if self.ld == L and world > 1:
            raise ValueError("the coarsest levels must be replicated: lower dist_min_n or use fewer ranks")

    def owned_rows(self, l, rank):
        nl = self.n // (2 ** l)
        per = nl // self.world
        return rank * per, (rank + 1) * per + (1 if rank == self.world - 1 else 0)


def _staged(t, group):
    """gloo has no CUDA point-to-point / all-gather: stage through the host (used by the one-GPU emulation tests)"""
    return t.is_cuda and dist.get_backend(group) == "gloo"


def halo_exchange(arr, lev, rank, world, group=None):
    """refresh the ghost rows of a local slab array (B=1: [1][nrows][pitch]) from the neighbouring ranks"""
    if world == 1 or not lev["dist"]:
        return
    r0, own0, own1, G = lev["row0"], lev["own0"], lev["own1"], lev["G"]
    a = arr[0]
    pairs = []  # (send view, recv view, peer)
    if rank > 0:  # neighbour above (smaller row indices)
        pairs.append((a[own0 - r0: own0 - r0 + G], a[own0 - r0 - G: own0 - r0], rank - 1))
    if rank < world - 1:  # neighbour below
        pairs.append((a[own1 - r0 - G: own1 - r0], a[own1 - r0: own1 - r0 + G], rank + 1))
    stage = _staged(arr, group)
    ops, bufs = [], []
    for send, recv, peer in pairs:
        sb = send.cpu() if stage else send
        rb = torch.empty(recv.shape, dtype=recv.dtype) if stage else recv
        bufs.append((recv, rb))
        ops += [dist.P2POp(dist.isend, sb, peer, group), dist.P2POp(dist.irecv, rb, peer, group)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    if stage:
        for recv, rb in bufs:
            recv.copy_(rb)


class SlabExchangePlan:
    """Pure host logic of the peer exchange: the byte layout of a rank's block (identical on every rank, arrays sized
    for the largest slab) and, per exchange step, WHAT goes WHERE -- copy jobs as (source offset, destination rank,
    destination offset, bytes), the flags to raise in the targets' mailboxes, the flags to wait for, the CTA count.
    No CUDA here, so tests/test_distributed_cpu.py checks every rank's plans against each other."""
    # mailbox (bytes): flags written by the neighbours / by every rank, partial-sum slots, this rank's step counters
    FLAG_UP, FLAG_DOWN, ALL_FLAGS, SLOTS, SEQ_NB, SEQ_ALL, ERR, PARTIAL, TOTAL, MAILBOX = 0, 16, 64, 256, 512, 528, 544, 576, 592, 4096
    # fused halo push of the finest up leg (mgfea_slab_prolong_correct_smooth_push): flags raised ONCE per launch by the
    # neighbour's kernel, this rank's expected value, and the two local ticket words
    FUSE_UP, FUSE_DOWN, SEQ_FUSE, TICKET = 32, 48, 560, 608
    MAX_PEERS = 8

    def __init__(self, part, pitch_for):
        self.part = part
        world, rank, ld = part.world, part.rank, part.ld
        if world > self.MAX_PEERS:
            raise ValueError(f"peer exchange supports up to {self.MAX_PEERS} ranks")
        self.parts = [part if q == rank else SlabPartition(part.n, part.L, world, q, part.dist_min_n) for q in range(world)]
        off, self.off, self.pitch = self.MAILBOX, {}, {}
        for l in range(ld + 1):
            N = part.levels[l]["N"]
            self.pitch[l] = pitch_for(N)
            rows = max(p.levels[l]["nrows"] for p in self.parts)
            size = (rows * self.pitch[l] * 4 + 255) // 256 * 256
            for name in (("u", "u_alt", "f") if l < ld else ("f",)):
                self.off[(name, l)] = off
                off += size
            if l == 0 and l < ld:  # fp64 iterate / right-hand side of the defect-correction solve (SolveMixed)
                for name in ("u64", "f64"):
                    self.off[(name, l)] = off
                    off += 2 * size
        self.nbytes = off

    @staticmethod
    def esize(name):
        return 8 if name.endswith("64") else 4

    def plan(self, halos, gather=False, reduce=False):
        """jobs [(src_off, dst_rank, dst_off, nbytes)], signals [(rank, flag_off)], waits [flag_off], seq_off, grid,
        red = (src_off, dst_off, n, stride) or None"""
        part, rank, world = self.part, self.part.rank, self.part.world
        jobs = []
        for name, l in halos:  # G boundary rows of my owned range -> the neighbours' ghost rows
            lev, rowb, off = part.levels[l], self.pitch[l] * self.esize(name), self.off[(name, l)]
            for q, first in ((rank - 1, lev["own0"]), (rank + 1, lev["own1"] - lev["G"])):
                if 0 <= q < world:
                    jobs.append((off + (first - lev["row0"]) * rowb, q,
                                 off + (first - self.parts[q].levels[l]["row0"]) * rowb, lev["G"] * rowb))
        if gather:  # my owned rows of the first replicated level's right-hand side -> every peer
            ld = part.ld
            per, rowb, off = (part.n // 2 ** ld) // world, self.pitch[ld] * 4, self.off[("f", ld)]
            jobs += [(off + rank * per * rowb, q, off + rank * per * rowb, per * rowb) for q in range(world) if q != rank]
        red = None
        if reduce:  # my partial sum -> slot [rank] of every mailbox (mine included); summed in rank order after the wait
            jobs += [(self.PARTIAL, q, self.SLOTS + 16 * rank, 16) for q in range(world)]
            red = (self.SLOTS, self.TOTAL, world, 16)
        if gather or reduce:
            signals = [(q, self.ALL_FLAGS + 16 * rank) for q in range(world) if q != rank]
            waits = [self.ALL_FLAGS + 16 * q for q in range(world) if q != rank]
            seq = self.SEQ_ALL
        else:
            signals, waits, seq = [], [], self.SEQ_NB
            if rank > 0:
                signals.append((rank - 1, self.FLAG_DOWN))
                waits.append(self.FLAG_UP)
            if rank < world - 1:
                signals.append((rank + 1, self.FLAG_UP))
                waits.append(self.FLAG_DOWN)
        # CTAs of the step: a function of the step alone (NOT of the rank): the flags count the pushing CTAs.  About one
        # 16-byte chunk per thread for the halo rows, more CTAs when whole coarse slabs travel
        halo_chunks = sum(part.levels[l]["G"] * self.pitch[l] * self.esize(nm) // 16 for nm, l in halos) * 2
        gather_chunks = ((part.n // 2 ** part.ld) // world) * self.pitch[part.ld] * 4 // 16 * (world - 1) if gather else 0
        grid = int(min(296, max(1, (halo_chunks + gather_chunks // 4 + 255) // 256)))
        return dict(jobs=jobs, signals=signals, waits=waits, seq=seq, grid=grid, red=red)


class PeerSlabMemory(SlabExchangePlan):
    """Slab arrays + exchange mailbox of one rank in a block the other ranks of the node map through cudaIpc
    (mgfea.PeerBlock), and the exchange steps over it (mgfea_p2p_exchange): the ranks store their boundary rows straight
    into the neighbours' ghost rows over NVLink, one kernel per step, no collective library on the data path.

    The layout is identical on every rank (SlabExchangePlan), so a peer's address of any row is
    base[q] + offset(array) + (global row - row0 of rank q) * pitch.
    """

    def __init__(self, part, group=None):
        import mgfea

        super().__init__(part, mgfea.pitch_for)
        self.mg, self.group = mgfea, group
        self.block = mgfea.PeerBlock(self.nbytes)
        self.bases = None
        self.partial = self.block.tensor(self.PARTIAL, (2,), torch.float64)  # [0]: this rank's interior sum of squares
        self.total = self.block.tensor(self.TOTAL, (1,), torch.float64)      # all-rank sum (same bits on every rank)
        self.err = self.block.tensor(self.ERR, (1,), torch.int32)
        self._steps = {}

    def connect(self, handles):
        """map the blocks of the other ranks (handles[q] = their PeerBlock.handle(), any transport)"""
        rank = self.part.rank
        self.bases = [self.block.base if q == rank else self.block.open_peer(handles[q]) for q in range(len(handles))]

    def array(self, name, l):
        lev = self.part.levels[l]
        dtype = torch.float64 if name.endswith("64") else torch.float32
        return self.block.tensor(self.off[(name, l)], (1, lev["nrows"], self.pitch[l]), dtype)

    def _build(self, halos, gather, reduce):
        """resolve a plan into the mgfea_xchg descriptor of this rank (addresses in the mapped blocks)"""
        mg, me = self.mg, self.bases[self.part.rank]
        pl = self.plan(halos, gather, reduce)
        x = mg.Xchg()
        for j, (so, q, do, nb) in enumerate(pl["jobs"]):
            x.src[j], x.dst[j], x.bytes[j] = me + so, self.bases[q] + do, nb
        for i, (q, fo) in enumerate(pl["signals"]):
            x.signal[i] = self.bases[q] + fo
        for i, fo in enumerate(pl["waits"]):
            x.wait[i] = me + fo
        x.njobs, x.nsignal, x.nwait = len(pl["jobs"]), len(pl["signals"]), len(pl["waits"])
        x.seq, x.err, x.grid = me + pl["seq"], me + self.ERR, pl["grid"]
        if pl["red"] is not None:
            so, do, n, stride = pl["red"]
            x.red_src, x.red_dst, x.nred, x.red_stride = me + so, me + do, n, stride
        x.mode = mg.XCHG_PUSH | mg.XCHG_WAIT
        return x

    def fused_waits(self):
        """(flag offsets this rank waits on after a fused push of its neighbours, seq offset): the neighbour ABOVE pushes
        its last rows DOWN into this rank's upper ghost rows and raises FUSE_DOWN here, and vice versa"""
        rank, world = self.part.rank, self.part.world
        return ([self.FUSE_DOWN] if rank > 0 else []) + ([self.FUSE_UP] if rank < world - 1 else []), self.SEQ_FUSE

    def push_desc(self, name="u", l=0):
        """mgfea_slab_push for this rank's finest up leg: where the first / last G owned rows of `name` go"""
        mg, part = self.mg, self.part
        rank, world, lev = part.rank, part.world, part.levels[l]
        rowb = self.pitch[l] * self.esize(name)
        d = mg.SlabPush()
        if rank > 0:  # element (global row 0, column 0) of the upper neighbour's array, in this process
            q = rank - 1
            d.up = self.bases[q] + self.off[(name, l)] - self.parts[q].levels[l]["row0"] * rowb
            d.flag_up = self.bases[q] + self.FUSE_UP
        if rank < world - 1:
            q = rank + 1
            d.dn = self.bases[q] + self.off[(name, l)] - self.parts[q].levels[l]["row0"] * rowb
            d.flag_dn = self.bases[q] + self.FUSE_DOWN
        d.ticket = self.bases[rank] + self.TICKET
        d.rows, d.own0, d.own1 = lev["G"], lev["own0"], lev["own1"]
        return d

    def step(self, halos, gather=False, reduce=False, fused=False, ctl=None):
        """one exchange step (a single kernel on the current stream); every rank must issue the same sequence.
        fused: also wait for the flags of the neighbours' fused-push kernels (their finest up leg of this cycle);
        ctl = (mgfea_ctl address, history address, capacity): device-side stopping rule on the reduced total"""
        key = (tuple(halos), gather, reduce, fused, ctl)
        x = self._steps.get(key)
        if x is None:
            if halos or gather or reduce:
                x = self._build(halos, gather, reduce)
            else:  # wait-only step
                x = self.mg.Xchg()
                x.err, x.grid, x.mode = self.bases[self.part.rank] + self.ERR, 1, 0
            if fused:
                waits, seq = self.fused_waits()
                me = self.bases[self.part.rank]
                for i, fo in enumerate(waits):
                    x.wait2[i] = me + fo
                x.nwait2, x.seq2 = len(waits), me + seq
            if ctl is not None:
                x.ctl, x.hist, x.hist_cap = ctl
            self._steps[key] = x
        if x.mode == 0 and x.nwait2 == 0:
            return
        self.mg.check(self.mg.lib().mgfea_p2p_exchange(ctypes.byref(x), self.mg.stream_ptr()))

    def check(self):
        """raises if a wait timed out (a peer died or fell out of step); synchronises"""
        e = int(self.err.item())
        if e:
            raise self.mg.MgfeaError(f"peer exchange: wait on flag {e - 1} timed out on rank {self.part.rank}")


class CudaSlabOps:
    """local operators on the GPU: the slab forms of the streaming kernels + the replicated coarse engine"""

    def __init__(self, part, nu1=1, nu2=1, coarse_f_store=None, prop=None, shape=0):
        import mgfea
        from .jacobi import JacobiBlock
        from .mesh import MeshSquare
        from .model import KNet
        from .solver import FULL_WEIGHTING_16, VCycleEngine

        if (nu1, nu2) != (1, 1):
            raise mgfea.MgfeaError("the row-slab path implements V(1,1)")
        self.mg = mgfea
        self.part = part
        self.dev = mgfea.require_cuda()
        n, L = part.n, part.L
        self.jacs = []
        for l in range(L):  # light-weight levels: no full-size host tensors (a 16385^2 mask alone is 1 GB)
            if prop is None:
                mesh = MeshSquare(2, n // 2 ** l + 1)
            else:  # two-phase inclusion (FEANet/mesh.py:4-120): every rank keeps the GLOBAL uint8 key map of a level
                from .mesh import MeshCenterInterface
                mesh = MeshCenterInterface(2, list(prop), n // 2 ** l + 1, shape=shape)
            self.jacs.append(JacobiBlock(KNet(mesh), mesh, 2 / 3., None, None))
        self.rtab = torch.from_numpy(FULL_WEIGHTING_16.reshape(1, 9).copy()).to(self.dev)
        # the replicated coarse cycle always starts from a zero guess and nobody reads its residual norm -- unless it IS
        # the whole problem (ld == 0: single rank / tiny grids)
        sub = part.ld > 0
        self.coarse = VCycleEngine(self.jacs[part.ld:], B=1, nu1=nu1, nu2=nu2, f0_store=coarse_f_store,
                                   compute_norm=not sub, zero_guess=sub) if part.ld < L else None
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=self.dev)

    def alloc(self, l):
        lev = self.part.levels[l]
        if l == self.part.ld and self.coarse is not None:
            return None  # the replicated engine owns these buffers
        return torch.zeros((1, lev["nrows"], self.mg.pitch_for(lev["N"])), dtype=torch.float32, device=self.dev)

    def coarse_f(self):
        return self.coarse.f[0].store

    def coarse_u(self):
        return self.coarse.u[0].store

    def _grid(self, l, arr):
        lev = self.part.levels[l]
        fld = self.mg.Field(1, lev["N"], self.dev, store=arr)  # pitch / N carrier; plane = local rows * pitch
        g = self.jacs[l].grid_struct(fld)
        g.plane = arr.shape[1] * arr.shape[2]
        return g

    def _slab(self, l, leg="own"):
        """rows held + rows COMPUTED by a leg: the owned rows, or the owned rows plus the deep-halo rows of that leg"""
        lev = self.part.levels[l]
        lo, hi = {"own": ("own0", "own1"), "dn": ("dn0", "dn1"), "up": ("up0", "up1")}[leg]
        return self.mg.Slab(lev["row0"], lev["nrows"], lev[lo], lev[hi])

    def down(self, l, u_in, u_out, f, fc):
        mg = self.mg
        g = self._grid(l, u_out)
        s, sc = self._slab(l, "dn"), self._slab(l + 1)
        mg.check(mg.lib().mgfea_slab_smooth_residual_restrict(
            ctypes.byref(g), ctypes.byref(s), u_in.data_ptr() if u_in is not None else None, u_out.data_ptr(),
            f.data_ptr(), fc.data_ptr(), ctypes.byref(sc), fc.shape[2], fc.shape[1] * fc.shape[2], self.rtab.data_ptr(),
            1, 4.0, None, 1, mg.stream_ptr()))

    def up(self, l, vc, u_in, u_out, f, want_norm, push=None, ctl=None):
        """push: mgfea.SlabPush -- the kernel also stores its boundary rows into the neighbours' ghost rows (fused halo);
        ctl: device address of the solve's mgfea_ctl (read only: a converged solve is left untouched)"""
        mg = self.mg
        g = self._grid(l, u_out)
        s, sc = self._slab(l, "up"), self._slab(l + 1)
        mg.check(mg.lib().mgfea_slab_prolong_correct_smooth_push(
            ctypes.byref(g), ctypes.byref(s), vc.data_ptr(), ctypes.byref(sc), vc.shape[2], vc.shape[1] * vc.shape[2],
            u_in.data_ptr(), u_out.data_ptr(), f.data_ptr(), self._sumsq.data_ptr() if want_norm else None,
            ctypes.byref(push) if push is not None else None, ctl, 1, mg.stream_ptr()))
        return self._sumsq if want_norm else None

    def coarse_cycle(self):
        """one V-cycle from a zero guess on the replicated levels; rhs / result live in the engine's level-0 buffers"""
        self.coarse.cycle()


class SlabMultigrid:
    """V(1,1) solver for the iso Poisson problem on row slabs.  `ops` supplies the local operators."""

    def __init__(self, n, ops_factory=CudaSlabOps, L=None, dist_min_n=2049, group=None, p2p=None, prop=None, shape=0):
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.group = group
        self.n = n
        self.L = int(math.log2(n)) if L is None else L
        self.part = SlabPartition(n, self.L, self.world, self.rank, dist_min_n)
        ld = self.part.ld
        # peer-memory exchange (NVLink P2P, no NCCL on the data path) whenever the CUDA operators run on > 1 rank
        self.peer = None
        if p2p is None:
            p2p = os.environ.get("MGFEA_P2P", "1") != "0"
        if p2p and self.world > 1 and 0 < ld < self.L and ops_factory is CudaSlabOps and torch.cuda.is_available():
            self.peer = self._try_peer_memory()
        kw = {} if prop is None else {"prop": prop, "shape": shape}  # two-phase conductivity (CUDA operators only)
        if self.peer is not None:
            self.ops = ops_factory(self.part, coarse_f_store=self.peer.array("f", ld), **kw)
            self.ops._sumsq = self.peer.partial
            self.u = [self.peer.array("u", l) for l in range(ld)]
            self.u_alt = [self.peer.array("u_alt", l) for l in range(ld)]
            self.f = [self.peer.array("f", l) for l in range(ld)]
        else:
            self.ops = ops_factory(self.part, **kw)
            self.u = [self.ops.alloc(l) for l in range(ld)]
            self.u_alt = [self.ops.alloc(l) for l in range(ld)]
            self.f = [self.ops.alloc(l) for l in range(ld)]
        # halo push fused into the finest up leg (the kernel's boundary strips store straight into the neighbours' ghost
        # rows).  Measured at 8 GPUs / 16385^2 (profiles/r02_slab_trace_n8_fused.log vs _nofuse.log): the exchange step
        # after the leg shrinks from 30 to 11-17 us, but the leg itself grows from 92-95 to 116-118 us -- the strips that
        # issue the peer stores wait for NVLink, and a one-wave streaming kernel is as slow as its slowest strip -- so the
        # separate exchange kernel stays the default; MGFEA_FUSED_PUSH=1 selects the fused path (parity-tested both ways)
        self.fused_push = self.peer is not None and os.environ.get("MGFEA_FUSED_PUSH", "0") == "1"
        self._push0 = self.peer.push_desc("u", 0) if self.fused_push else None
        if self.peer is not None:  # mgfea_ctl + residual history of the device-side stopping rule (free-running by default)
            self.max_cycles = 256
            self.ctl = torch.zeros(8, dtype=torch.int32, device=self.ops.dev)
            self.hist = torch.zeros(self.max_cycles, dtype=torch.float64, device=self.ops.dev)
            self._ctl_set(0, -1.0, 2 ** 31 - 1)
        self.residuals = []
        self._graph = None
        self._graph64 = None
        self._mixed_steps = 0
        self._graph_out = None
        self._graph_err = None

    def _try_peer_memory(self):
        """collective: all ranks map each other's block, or all fall back to the NCCL exchange"""
        mem, handle, why = None, None, None
        try:
            mem = PeerSlabMemory(self.part, self.group)
            handle = mem.block.handle()
        except Exception as e:  # noqa: BLE001  (allocation / cudaIpc export unavailable)
            why = repr(e)
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=self.group)
        if why is None and all(h is not None for h in handles):
            try:
                mem.connect(handles)
            except Exception as e:  # noqa: BLE001  (peers without P2P access, IPC blocked by the container, ...)
                why = repr(e)
        elif why is None:
            why = "a peer could not export its block"
        ok = [None] * self.world
        dist.all_gather_object(ok, why, group=self.group)
        bad = [w for w in ok if w is not None]
        self.peer_error = bad[0] if bad else None
        return mem if not bad else None

    # ---- problem data: every rank takes its rows (owned + ghost) from the full host arrays
    def set_problem(self, u0_full, f_full):
        lev = self.part.levels[0]
        N = lev["N"]
        rows = slice(lev["row0"], lev["row0"] + lev["nrows"])
        if self.part.ld == 0:  # everything replicated (tiny problems / single rank)
            self.ops.coarse.set_u(u0_full)
            self.ops.coarse.set_f(f_full)
            return
        for dst, src in ((self.u[0], u0_full), (self.f[0], f_full)):
            src = torch.as_tensor(src, dtype=torch.float32).reshape(N, N)
            dst[0, :, :N].copy_(src[rows], non_blocking=True)

    def fill_local(self, fn_u, fn_f=None):
        """set the problem from per-rank generators: fn(row0, nrows, N) -> (nrows, N) float32 tensor of the global rows
        [row0, row0+nrows) (owned rows only need to be right: ghost rows are refreshed by the first halo exchange)"""
        lev = self.part.levels[0]
        N = lev["N"]
        if self.part.ld == 0:
            self.ops.coarse.set_u(fn_u(0, N, N))
            self.ops.coarse.set_f(fn_f(0, N, N) if fn_f else torch.zeros(N, N))
            return
        self.u[0][0, :, :N].copy_(fn_u(lev["row0"], lev["nrows"], N))
        if fn_f is not None:
            self.f[0][0, :, :N].copy_(fn_f(lev["row0"], lev["nrows"], N))
        else:
            self.f[0].zero_()

    def gather_solution(self):
        """full level-0 solution on every rank (host tensor); for tests / result delivery"""
        lev = self.part.levels[0]
        N = lev["N"]
        if self.part.ld == 0:
            return self.ops.coarse.solution.cpu().reshape(N, N).clone()
        own = self.u[0][0, lev["own0"] - lev["row0"]: lev["own1"] - lev["row0"], :N].contiguous()
        if self.world == 1:
            return own.cpu()
        per = self.n // self.world
        chunk = torch.zeros((per + 1, N), dtype=own.dtype)
        chunk[: own.shape[0]] = own.cpu()
        out = [torch.zeros_like(chunk) for _ in range(self.world)]
        if dist.get_backend(self.group) == "nccl":
            dev = own.device
            outd = [o.to(dev) for o in out]
            dist.all_gather(outd, chunk.to(dev), group=self.group)
            out = [o.cpu() for o in outd]
        else:
            dist.all_gather(out, chunk, group=self.group)
        full = torch.cat([o[:per] for o in out[:-1]] + [out[-1][: per + 1]], dim=0)
        return full.cpu()

    def _allgather_coarse_f(self):
        """first replicated level: every rank wrote its owned coarse rows into the full array; make it complete"""
        ld = self.part.ld
        if self.world == 1:
            return
        fc = self.ops.coarse_f()  # [1][Nc][pitch]
        nl = self.n // (2 ** ld)
        per = nl // self.world
        body = fc[0, :nl]  # the last row (ring) is zero on every rank
        mine = body[self.rank * per:(self.rank + 1) * per]
        if _staged(fc, self.group):
            outs = [torch.empty(mine.shape, dtype=mine.dtype) for _ in range(self.world)]
            dist.all_gather(outs, mine.cpu(), group=self.group)
            for r, o in enumerate(outs):
                body[r * per:(r + 1) * per].copy_(o)
        else:
            dist.all_gather_into_tensor(body.reshape(-1), mine.clone().reshape(-1), group=self.group) \
                if fc.is_cuda else dist.all_gather([body[r * per:(r + 1) * per] for r in range(self.world)],
                                                   mine.clone(), group=self.group)

    def enable_graph(self, warm=3):
        """capture one whole cycle -- kernels AND the NCCL halo exchanges / all-gather / all-reduce -- in a CUDA graph
        (removes ~25 host-side launches per cycle); falls back to eager launches if capture is not possible"""
        if self._graph is not None or not torch.cuda.is_available():
            return self._graph is not None
        try:
            for _ in range(warm):  # communicators, function attributes and scratch must exist before capture
                self._cycle_eager()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._graph_out = self._cycle_eager()
            self._graph = g
        except Exception as e:  # noqa: BLE001
            self._graph = None
            self._graph_err = repr(e)
            torch.cuda.synchronize()
        return self._graph is not None

    def cycle(self, want_norm=True):
        if self._graph is not None:
            self._graph.replay()
            return self._graph_out
        return self._cycle_eager(want_norm)

    def _cycle_eager(self, want_norm=True):
        p, ops, ld = self.part, self.ops, self.part.ld
        if ld == 0:
            ops.coarse.cycle()
            return ops.coarse.sumsq.clone()
        if self.peer is not None:
            return self._cycle_peer(want_norm)
        # ---- down leg on the slabs: no exchange, every level computes its deep-halo rows itself (halo_depths)
        for l in range(ld):
            fc = self.f[l + 1] if l + 1 < ld else ops.coarse_f()
            ops.down(l, self.u[l] if l == 0 else None, self.u_alt[l], self.f[l], fc)
        # ---- replicated coarse levels
        self._allgather_coarse_f()
        ops.coarse_cycle()
        # ---- up leg
        ss = None
        for l in range(ld - 1, -1, -1):
            vc = self.u[l + 1] if l + 1 < ld else ops.coarse_u()
            ss = ops.up(l, vc, self.u_alt[l], self.u[l], self.f[l], want_norm and l == 0)
        halo_exchange(self.u[0], p.levels[0], self.rank, self.world, self.group)  # the only halo exchange of the cycle
        if want_norm:
            tot = ss.clone()
            if self.world > 1:
                if _staged(tot, self.group):
                    t = tot.cpu()
                    dist.all_reduce(t, group=self.group)
                    tot.copy_(t)
                else:
                    dist.all_reduce(tot, group=self.group)
            return tot
        return None

    def _cycle_peer(self, want_norm=True, zero_guess=False, push_u=True):
        """the same cycle with the two exchanges done by peer stores (PeerSlabMemory.step), no NCCL: the gather of the first
        replicated right-hand side after the last down leg, and after the finest up leg the push of the G[0] boundary rows
        of the new iterate into the neighbours' ghost rows together with the all-reduce of the residual norm.
        Overwriting a neighbour's ghost rows needs no second handshake: u[0]'s ghost rows are read by down(0) only, and a
        rank reaches its next push of u[0] only after the gather of the NEXT cycle, which every rank enters after its
        down(0).  The gathered right-hand side is read by the coarse cycle only, and the next gather follows the push /
        reduce step that every rank enters after its coarse cycle."""
        ops, ld, peer = self.ops, self.part.ld, self.peer
        for l in range(ld):
            last = l == ld - 1
            ops.down(l, self.u[l] if (l == 0 and not zero_guess) else None, self.u_alt[l], self.f[l],
                     ops.coarse_f() if last else self.f[l + 1])
        peer.step((), gather=True)
        ops.coarse_cycle()
        fuse = push_u and self.fused_push
        # device-side stopping rule (Solve): the finest up leg and the final step do nothing once the solve is done, so
        # cycles enqueued past convergence leave the solution alone; all ranks reduce the same total -> same decision
        ctl = (self.ctl.data_ptr(), self.hist.data_ptr(), self.hist.shape[0]) if (want_norm and not zero_guess) else None
        for l in range(ld - 1, -1, -1):
            vc = self.u[l + 1] if l + 1 < ld else ops.coarse_u()
            ops.up(l, vc, self.u_alt[l], self.u[l], self.f[l], want_norm and l == 0,
                   push=self._push0 if (fuse and l == 0) else None, ctl=ctl[0] if (ctl and l == 0) else None)
        if fuse:  # the rows are already on their way (stored by the up leg itself): wait for the neighbours' flags
            peer.step((), reduce=want_norm, fused=True, ctl=ctl)
        elif push_u or want_norm:
            peer.step((("u", 0),) if push_u else (), reduce=want_norm, ctl=ctl)
        return peer.total if want_norm else None

    def _sync_ranks(self):
        """Order every rank's LOCAL writes to its slab arrays (set_problem / fill_local / zero_(), ghost rows included)
        before any peer starts storing into those ghost rows: a neighbour that runs ahead would otherwise have its pushed
        rows wiped by this rank's own initialisation while the flag still counts the push."""
        if self.world > 1:
            if torch.cuda.is_available():
                torch.cuda.current_stream().synchronize()
            dist.barrier(group=self.group)

    def close(self):
        """drop the captured graphs and unmap / free the peer block (IPC mappings do not go away with the Python object)"""
        self._graph = self._graph64 = None
        if getattr(self.ops, "coarse", None) is not None:
            self.ops.coarse._graph = None
        if self.peer is not None:
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier(group=self.group)  # nobody may still be storing into a block that is about to be freed
            self.u = self.u_alt = self.f = []
            self.peer.partial = self.peer.total = self.peer.err = None
            self.ops._sumsq = None
            self.ops.coarse = None
            self.peer.block.close()
            self.peer = None

    def exchange_initial(self):
        """ghost rows of the level-0 iterate and right-hand side (after set_problem / fill_local)"""
        if self.part.ld == 0:
            return
        self._sync_ranks()
        if self.peer is not None:
            self.peer.step((("u", 0), ("f", 0)))
        else:
            halo_exchange(self.u[0], self.part.levels[0], self.rank, self.world, self.group)
            halo_exchange(self.f[0], self.part.levels[0], self.rank, self.world, self.group)

    # ---- fp64 defect correction on the slabs (SURVEY 8f.1 + 8e): iterate / rhs / residual in fp64, cycle in fp32
    def set_problem64(self, u0_full, f_full):
        """every rank takes its rows (owned + ghost) of the fp64 problem from the full host arrays; the ring of the
        iterate is zeroed (the reference's first reset_boundary)"""
        if self.peer is None:
            raise self.ops.mg.MgfeaError("SolveMixed on slabs needs the peer-memory exchange")
        lev = self.part.levels[0]
        N = lev["N"]
        rows = slice(lev["row0"], lev["row0"] + lev["nrows"])
        self.u64, self.f64 = self.peer.array("u64", 0), self.peer.array("f64", 0)
        for dst, src, ring in ((self.u64, u0_full, True), (self.f64, f_full, False)):
            src = torch.as_tensor(src).reshape(N, N).to(torch.float64)
            if ring:
                src = src.clone()
                src[0, :] = 0
                src[-1, :] = 0
                src[:, 0] = 0
                src[:, -1] = 0
            dst.zero_()
            dst[0, :, :N].copy_(src[rows], non_blocking=True)

    def fill_local64(self, fn_f):
        """fp64 problem from a per-rank generator: u0 = 0, f rows [own0, own1) = fn_f(own0, nrows, N) (float64 device
        tensor); the ghost rows of f come from the neighbours"""
        if self.peer is None:
            raise self.ops.mg.MgfeaError("SolveMixed on slabs needs the peer-memory exchange")
        lev = self.part.levels[0]
        N = lev["N"]
        self.u64, self.f64 = self.peer.array("u64", 0), self.peer.array("f64", 0)
        self.u64.zero_()
        self.f64.zero_()
        o0, o1 = lev["own0"] - lev["row0"], lev["own1"] - lev["row0"]
        self.f64[0, o0:o1, :N].copy_(fn_f(lev["own0"], o1 - o0, N))
        self._sync_ranks()  # the zero fill above covers the ghost rows the neighbours are about to store into
        self.peer.step((("f64", 0),))

    def SolveMixed(self, n_iter=None, EPS=None, max_cycles=200, use_graph=True):
        """Multigrid.Solve semantics on the fp64 problem set by set_problem64: every step is one slab V-cycle (fp32, zero
        guess) on the fp64 residual.  Returns the fp64 interior residual 2-norms after each cycle; `self.u64` holds the
        local rows of the solution (gather_solution64 assembles it)."""
        peer = self.peer
        if n_iter is None:
            n_iter = 0
        elif EPS is None:
            EPS = math.inf
        self._sync_ranks()  # set_problem64 / fill_local64 wrote ghost rows locally; peers push into them next
        self._defect64()
        self.r0 = float(torch.sqrt(peer.total.sum()).item())
        res, hist = self.r0, []
        while (res > EPS or len(hist) < n_iter) and len(hist) < max_cycles:
            # the first step runs eagerly (lazy one-time initialisation must not happen inside a capture)
            if use_graph and self._graph64 is None and self._mixed_steps >= 1 and torch.cuda.is_available():
                self._capture_graph64()
            self._mixed_steps += 1
            if self._graph64 is not None:
                self._graph64.replay()
            else:
                self._mixed_iter()
            res = float(torch.sqrt(peer.total.sum()).item())
            hist.append(res)
            peer.check()
        peer.check()
        self.residuals = hist
        return hist

    def _defect64(self):
        """r = f64 - K u64 -> f[0] (owned rows + 3 ghost rows per side), all-rank interior sum of squares -> peer.total"""
        mg, ops, peer = self.ops.mg, self.ops, self.peer
        g, sl = ops._grid(0, self.f[0]), ops._slab(0)
        lev = self.part.levels[0]
        peer.step((("u64", 0),))  # G[0] ghost rows of the iterate
        # the zero-guess down leg of level 0 reads its right-hand side on the deep-halo rows + 2
        mg.check(mg.lib().mgfea_slab_defect_f64_ext(ctypes.byref(g), ctypes.byref(sl), lev["G"] - 2, self.u64.data_ptr(),
                                                    self.f64.data_ptr(), self.f[0].data_ptr(), peer.partial.data_ptr(), 1,
                                                    mg.stream_ptr()))
        peer.step((), reduce=True)

    def _mixed_iter(self):
        """e = V-cycle(0, r) (fp32, slabs); u64 += e; new defect + norm"""
        mg, ops = self.ops.mg, self.ops
        self._cycle_peer(want_norm=False, zero_guess=True, push_u=False)  # the correction is used on the owned rows only
        g, sl = ops._grid(0, self.f[0]), ops._slab(0)
        mg.check(mg.lib().mgfea_slab_correct_f64(ctypes.byref(g), ctypes.byref(sl), self.u64.data_ptr(), self.u[0].data_ptr(),
                                                 1, mg.stream_ptr()))
        self._defect64()

    def _capture_graph64(self):
        """one defect-correction step (slab cycle + exchanges + fp64 kernels) as a CUDA graph; every rank captures after
        the same number of eager steps, so the exchange counters stay aligned"""
        try:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._mixed_iter()
            self._graph64 = g
        except Exception as e:  # noqa: BLE001
            self._graph64, self._graph_err = None, repr(e)
            raise

    def gather_solution64(self):
        """full fp64 solution on every rank (host tensor)"""
        lev = self.part.levels[0]
        N = lev["N"]
        own = self.u64[0, lev["own0"] - lev["row0"]: lev["own1"] - lev["row0"], :N].contiguous()
        per = self.n // self.world
        chunk = torch.zeros((per + 1, N), dtype=own.dtype)
        chunk[: own.shape[0]] = own.cpu()
        out = [torch.zeros_like(chunk) for _ in range(self.world)]
        if dist.get_backend(self.group) == "nccl":
            outd = [o.to(own.device) for o in out]
            dist.all_gather(outd, chunk.to(own.device), group=self.group)
            out = [o.cpu() for o in outd]
        else:
            dist.all_gather(out, chunk, group=self.group)
        return torch.cat([o[:per] for o in out[:-1]] + [out[-1][: per + 1]], dim=0)

    def _ctl_set(self, min_cycles, eps2, max_cycles):
        c = self.ops.mg.Ctl()
        c.cycle, c.done, c.min_cycles, c.max_cycles, c.conv_rule, c.eps2 = 0, 0, int(min_cycles), int(max_cycles), 0, eps2
        self.ctl.copy_(torch.from_numpy(np.frombuffer(bytes(c), dtype=np.int32).copy()))

    def Solve(self, n_iter=None, EPS=None, max_cycles=200, chunk=4):
        """Multigrid.Solve semantics (MM_Model_convergence.ipynb cell 3): cycles while (res > EPS or n < n_iter).
        Peer path: the loop condition is evaluated on the device by the all-reduce step (identical on every rank); the host
        enqueues `chunk` cycles (graph replays once captured) per synchronisation, cycles past convergence are no-ops on the
        solution.  NCCL fallback: one host synchronisation per cycle like the reference's .item()."""
        if n_iter is None:
            n_iter = 0
        elif EPS is None:
            EPS = math.inf
        self.exchange_initial()
        if self.peer is None or self.part.ld == 0:
            res, hist = 1.0, []
            while (res > EPS or len(hist) < n_iter) and len(hist) < max_cycles:
                res = float(torch.sqrt(self.cycle().sum()).item())
                hist.append(res)
            self.residuals = hist
            return hist
        cap = min(max_cycles, self.max_cycles)
        if n_iter > cap:
            raise self.ops.mg.MgfeaError(f"n_iter={n_iter} exceeds the history capacity {cap}")
        self._ctl_set(n_iter, float(EPS) ** 2 if math.isfinite(EPS) else 1.7e308, cap)
        if n_iter > 0 and not math.isfinite(EPS):
            chunk = n_iter  # fixed cycle count: one synchronisation at the end
        done = False
        while not done:
            for _ in range(chunk):
                self.cycle()
            c = self.ctl.cpu()
            done = bool(c[1].item())
            self.peer.check()  # a timed-out wait ends the solve here: nothing runs on with stale ghost rows
        ncyc = int(c[0].item())
        hist = [float(math.sqrt(v)) for v in self.hist[:ncyc].cpu().numpy()]
        self._ctl_set(0, -1.0, 2 ** 31 - 1)  # back to free-running for cycle() users
        self.peer.check()
        self.residuals = hist
        return hist
