"""ctypes binding of libmgfea.so (C ABI in include/mgfea.h) for torch CUDA tensors.

PyTorch is plumbing here: device memory, streams, H2D/D2H copies.  Every numerical operation on the V-cycle path is a
call into the hand-written sm_100a kernels of ``csrc/``.  There is NO CPU / eager fallback: if the shared library is
missing or no CUDA device is present, operations raise ``MgfeaError``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MGFEA_LIB") or os.path.join(_HERE, "libmgfea.so")  # MGFEA_LIB: A/B builds while tuning
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.abspath(os.path.join(_HERE, "..", "..", "include"))

SMOOTH_JACOBI, SMOOTH_HJACOBI = 0, 1
PROLONG_BILINEAR, PROLONG_TABLE = 1, 3
CONV_SUM, CONV_MAX = 0, 1

EXPORTS = [
    "mgfea_version", "mgfea_error_string", "mgfea_set_loader", "mgfea_set_option", "mgfea_launch_count", "mgfea_pack", "mgfea_unpack",
    "mgfea_stiffness_apply", "mgfea_load_vector", "mgfea_split_x", "mgfea_reset_boundary", "mgfea_smooth",
    "mgfea_residual", "mgfea_restrict", "mgfea_smooth_residual_restrict", "mgfea_prolong_correct_smooth",
    "mgfea_residual_norm", "mgfea_vcycle", "mgfea_restrict_channels", "mgfea_prolong_channels",
    "mgfea_slab_smooth_residual_restrict", "mgfea_slab_prolong_correct_smooth",
    "mgfea_peer_alloc", "mgfea_peer_free", "mgfea_peer_export", "mgfea_peer_open", "mgfea_peer_close",
    "mgfea_p2p_exchange", "mgfea_trace", "mgfea_prolong_correct_smooth_norm",
    "mgfea_defect_f64", "mgfea_correct_f64", "mgfea_slab_defect_f64", "mgfea_slab_correct_f64", "mgfea_pattern_keys",
    "mgfea_widen_f64", "mgfea_slab_defect_f64_ext", "mgfea_slab_prolong_correct_smooth_push",
    "mgfea_elem_stiffness_apply", "mgfea_elem_residual", "mgfea_elem_smooth", "mgfea_elem_coarsen", "mgfea_sumsq_interior",
    "mgfea_smooth_pbc", "mgfea_corr9", "mgfea_restrict_adjoint", "mgfea_prolong_adjoint", "mgfea_restrict_wgrad",
    "mgfea_prolong_wgrad",
]


class MgfeaError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile csrc/mgfea.cu for sm_100a into libmgfea.so (in-tree).  nvcc cross-compiles without a GPU."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "mgfea.h")]
    if os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    # -fmad=false: every multiply-add the arithmetic contract allows is written as an explicit fma intrinsic; nothing
    # else may be contracted (NB: ptxas still fuses inline-PTX mul.rn.f32x2 + add.rn.f32x2, see mgfea_stream.cuh)
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-fmad=false", "-lineinfo", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC", os.path.join(CSRC, "mgfea.cu"), "-o", LIB_PATH]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return LIB_PATH


class Grid(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int32), ("pitch", ctypes.c_int32), ("plane", ctypes.c_int64), ("npat", ctypes.c_int32),
                ("key_pitch", ctypes.c_int32), ("keys", ctypes.c_void_p), ("ktab", ctypes.c_void_p),
                ("invd", ctypes.c_void_p), ("bc_idx", ctypes.c_void_p), ("bc_val", ctypes.c_void_p),
                ("bc_plane", ctypes.c_int64)]


class Ctl(ctypes.Structure):
    _fields_ = [("cycle", ctypes.c_int32), ("done", ctypes.c_int32), ("min_cycles", ctypes.c_int32),
                ("max_cycles", ctypes.c_int32), ("conv_rule", ctypes.c_int32), ("pad_", ctypes.c_int32),
                ("eps2", ctypes.c_double)]


class CycleCfg(ctypes.Structure):
    _fields_ = [("nu1", ctypes.c_int32), ("nu2", ctypes.c_int32), ("smoother", ctypes.c_int32),
                ("nlayers", ctypes.c_int32), ("hw", ctypes.c_void_p), ("prolong_mode", ctypes.c_int32),
                ("rtab_n", ctypes.c_int32), ("rtab", ctypes.c_void_p), ("ptab_n", ctypes.c_int32),
                ("r_has_scale", ctypes.c_int32), ("ptab", ctypes.c_void_p), ("r_scale_host", ctypes.c_float),
                ("p_scale_host", ctypes.c_float), ("r_scale_dev", ctypes.c_void_p), ("p_scale_dev", ctypes.c_void_p),
                ("p_has_scale", ctypes.c_int32), ("quirk_level0", ctypes.c_int32), ("tail_max_n", ctypes.c_int32),
                ("compute_norm", ctypes.c_int32), ("zero_guess", ctypes.c_int32), ("pad_", ctypes.c_int32)]


class Slab(ctypes.Structure):
    _fields_ = [("row0", ctypes.c_int32), ("nrows", ctypes.c_int32), ("own0", ctypes.c_int32), ("own1", ctypes.c_int32)]


XCHG_MAX_JOBS, XCHG_MAX_PEERS, XCHG_PUSH, XCHG_WAIT, IPC_HANDLE_BYTES = 16, 8, 1, 2, 64


class Xchg(ctypes.Structure):
    """mgfea_xchg: one exchange step of one rank (include/mgfea.h)"""
    _fields_ = [("njobs", ctypes.c_int32), ("nsignal", ctypes.c_int32), ("nwait", ctypes.c_int32),
                ("mode", ctypes.c_int32), ("src", ctypes.c_void_p * XCHG_MAX_JOBS),
                ("dst", ctypes.c_void_p * XCHG_MAX_JOBS), ("bytes", ctypes.c_uint64 * XCHG_MAX_JOBS),
                ("signal", ctypes.c_void_p * XCHG_MAX_PEERS), ("wait", ctypes.c_void_p * XCHG_MAX_PEERS),
                ("seq", ctypes.c_void_p), ("err", ctypes.c_void_p), ("red_src", ctypes.c_void_p),
                ("red_dst", ctypes.c_void_p), ("nred", ctypes.c_int32), ("red_stride", ctypes.c_int32),
                ("grid", ctypes.c_int32), ("nwait2", ctypes.c_int32), ("wait2", ctypes.c_void_p * 2),
                ("seq2", ctypes.c_void_p), ("ctl", ctypes.c_void_p), ("hist", ctypes.c_void_p),
                ("hist_cap", ctypes.c_int32), ("pad2_", ctypes.c_int32)]


class SlabPush(ctypes.Structure):
    """mgfea_slab_push: fused halo push of the finest slab up leg (include/mgfea.h)"""
    _fields_ = [("up", ctypes.c_void_p), ("dn", ctypes.c_void_p), ("flag_up", ctypes.c_void_p),
                ("flag_dn", ctypes.c_void_p), ("ticket", ctypes.c_void_p), ("rows", ctypes.c_int32),
                ("own0", ctypes.c_int32), ("own1", ctypes.c_int32)]


class LevelBufs(ctypes.Structure):
    _fields_ = [("u", ctypes.c_void_p), ("u_alt", ctypes.c_void_p), ("f", ctypes.c_void_p)]


_lib = None


def lib():
    """The loaded C-ABI library; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MgfeaError(f"{LIB_PATH} not found: run __graft_entry__.build() (nvcc, sm_100a). "
                             "There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        L.mgfea_version.restype = ctypes.c_char_p
        L.mgfea_error_string.restype = ctypes.c_char_p
        L.mgfea_error_string.argtypes = [ctypes.c_int]
        L.mgfea_launch_count.restype = ctypes.c_uint64
        vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
        G = ctypes.POINTER(Grid)
        L.mgfea_set_loader.argtypes = [i32]
        L.mgfea_set_option.argtypes = [ctypes.c_char_p, i32]
        L.mgfea_trace.argtypes = [vp, i32]
        L.mgfea_pack.argtypes = [vp, vp, i32, i32, i64, i32, vp]
        L.mgfea_unpack.argtypes = [vp, vp, i32, i32, i64, i32, vp]
        L.mgfea_stiffness_apply.argtypes = [G, vp, vp, i32, vp]
        L.mgfea_load_vector.argtypes = [vp, vp, vp, i32, i32, i64, i32, vp]
        L.mgfea_split_x.argtypes = [G, vp, vp, i32, vp]
        L.mgfea_reset_boundary.argtypes = [G, vp, vp, i32, vp]
        L.mgfea_smooth.argtypes = [G, vp, vp, vp, i32, i32, vp, i32, i32, vp]
        L.mgfea_residual.argtypes = [G, vp, vp, vp, i32, vp]
        L.mgfea_restrict.argtypes = [G, vp, vp, i32, i64, vp, i32, i32, f32, vp, i32, vp]
        L.mgfea_smooth_residual_restrict.argtypes = [G, vp, vp, vp, i32, i32, vp, i32, vp, i32, i64, vp, i32, i32, f32,
                                                     vp, i32, vp]
        L.mgfea_prolong_correct_smooth.argtypes = [G, G, vp, vp, vp, vp, i32, vp, i32, i32, f32, vp, i32, i32, vp, i32,
                                                   i32, vp]
        L.mgfea_prolong_correct_smooth_norm.argtypes = [G, G, vp, vp, vp, vp, i32, vp, i32, i32, f32, vp, i32, i32, vp,
                                                        i32, vp, i32, vp]
        L.mgfea_defect_f64.argtypes = [G, vp, vp, vp, vp, vp, vp, i32, vp]
        L.mgfea_correct_f64.argtypes = [G, vp, vp, vp, i32, vp]
        L.mgfea_slab_defect_f64.argtypes = [G, ctypes.POINTER(Slab), vp, vp, vp, vp, i32, vp]
        L.mgfea_slab_correct_f64.argtypes = [G, ctypes.POINTER(Slab), vp, vp, i32, vp]
        L.mgfea_slab_defect_f64_ext.argtypes = [G, ctypes.POINTER(Slab), i32, vp, vp, vp, vp, i32, vp]
        L.mgfea_pattern_keys.argtypes = [vp, i32, i32, i32, vp]
        L.mgfea_elem_stiffness_apply.argtypes = [vp, vp, vp, i32, i32, i64, i32, vp]
        L.mgfea_elem_residual.argtypes = [vp, vp, vp, vp, i32, i32, i64, i32, vp]
        L.mgfea_elem_smooth.argtypes = [vp, vp, vp, vp, f32, i32, i32, i64, i32, vp]
        L.mgfea_elem_coarsen.argtypes = [vp, vp, i32, i32, i32, vp]
        L.mgfea_sumsq_interior.argtypes = [vp, vp, i32, i32, i64, i32, vp]
        L.mgfea_restrict_adjoint.argtypes = [G, G, vp, i32, f32, vp, vp, i32, vp]
        L.mgfea_prolong_adjoint.argtypes = [G, G, vp, i32, f32, vp, vp, i32, vp]
        L.mgfea_restrict_wgrad.argtypes = [G, G, i32, f32, vp, vp, vp, i32, vp]
        L.mgfea_prolong_wgrad.argtypes = [G, G, i32, f32, vp, vp, vp, i32, vp]
        L.mgfea_corr9.argtypes = [vp, vp, vp, i32, i32, i64, i32, vp]
        L.mgfea_smooth_pbc.argtypes = [vp, vp, vp, vp, vp, i32, i32, i64, i32, i64, i32, vp]
        L.mgfea_widen_f64.argtypes = [vp, vp, i32, i32, i64, i32, i32, vp]
        L.mgfea_restrict_channels.argtypes = [vp, vp, vp, i32, i32, i32, vp]
        L.mgfea_prolong_channels.argtypes = [vp, vp, vp, i32, i32, i32, vp]
        L.mgfea_residual_norm.argtypes = [G, vp, vp, vp, vp, vp, i32, vp]
        S = ctypes.POINTER(Slab)
        L.mgfea_slab_smooth_residual_restrict.argtypes = [G, S, vp, vp, vp, vp, S, i32, i64, vp, i32, f32, vp, i32, vp]
        L.mgfea_slab_prolong_correct_smooth.argtypes = [G, S, vp, S, i32, i64, vp, vp, vp, vp, i32, vp]
        L.mgfea_slab_prolong_correct_smooth_push.argtypes = [G, S, vp, S, i32, i64, vp, vp, vp, vp,
                                                             ctypes.POINTER(SlabPush), vp, i32, vp]
        L.mgfea_peer_alloc.argtypes = [ctypes.POINTER(vp), ctypes.c_uint64]
        L.mgfea_peer_free.argtypes = [vp]
        L.mgfea_peer_export.argtypes = [vp, vp]
        L.mgfea_peer_open.argtypes = [vp, ctypes.POINTER(vp)]
        L.mgfea_peer_close.argtypes = [vp]
        L.mgfea_p2p_exchange.argtypes = [ctypes.POINTER(Xchg), vp]
        L.mgfea_vcycle.argtypes = [ctypes.POINTER(Grid), ctypes.POINTER(LevelBufs), i32, ctypes.POINTER(CycleCfg), vp,
                                   vp, vp, i32, vp]
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise MgfeaError(f"libmgfea error {rc}: {lib().mgfea_error_string(rc).decode()}")


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise MgfeaError("mgfea needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def bind_to_gpu_numa(device_index: Optional[int] = None) -> Optional[int]:
    """Pin this process (and therefore the pinned host buffers it allocates afterwards: first touch) to the NUMA node
    the GPU hangs off.  One process per GPU on an 8-GPU box otherwise leaves every rank on node 0 and the host<->device
    copies of all ranks share one memory controller / PCIe root (measured: 13.5 GB/s per GPU at 8 ranks vs 31 GB/s at 2).
    Returns the node, or None when the topology cannot be read (nothing is changed then)."""
    try:
        dev = torch.cuda.current_device() if device_index is None else device_index
        pr = torch.cuda.get_device_properties(dev)
        if hasattr(pr, "pci_bus_id"):
            bdf = f"{getattr(pr, 'pci_domain_id', 0):04x}:{pr.pci_bus_id:02x}:{getattr(pr, 'pci_device_id', 0):02x}.0"
        else:  # older torch: ask the driver (match by UUID, CUDA_VISIBLE_DEVICES may have renumbered the devices)
            rows = subprocess.run(["nvidia-smi", "--query-gpu=uuid,pci.bus_id", "--format=csv,noheader"],
                                  capture_output=True, text=True, timeout=10).stdout.strip().splitlines()
            table = [tuple(c.strip() for c in r.split(",")) for r in rows]
            uuid = str(getattr(pr, "uuid", ""))
            hit = [b for u, b in table if uuid and uuid.replace("GPU-", "") in u] or [table[dev][1]]
            dom, rest = hit[0].lower().split(":", 1)
            bdf = f"{dom[-4:]}:{rest}"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:  # noqa: BLE001  (containers may hide sysfs; binding is an optimisation only)
        return None


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(lib().mgfea_launch_count())


def set_loader(use_tma: bool) -> int:
    return int(lib().mgfea_set_loader(1 if use_tma else 0))


def set_option(name: str, value: int) -> int:
    """kernel-selection threshold `name` (see include/mgfea.h mgfea_set_option); returns the previous value.  Engines
    that captured a CUDA graph keep replaying the kernels chosen at capture time."""
    rc = int(lib().mgfea_set_option(name.encode(), int(value)))
    if rc < 0:
        raise MgfeaError(f"mgfea_set_option({name!r}): {rc}")
    return rc


# ----------------------------------------------------------------------------------------------------------
# padded-pitch fields
# ----------------------------------------------------------------------------------------------------------
def pitch_for(N: int) -> int:
    """row pitch in floats: rows start on 128-byte boundaries (N = 2^k + 1 is odd)"""
    return (N + 31) // 32 * 32


class Field:
    """fp32 field [B][N][pitch] in HBM; `.view` is the (B,1,N,N) strided torch view handed to API users."""

    __slots__ = ("store", "B", "N", "pitch")

    def __init__(self, B: int, N: int, device=None, store: Optional[torch.Tensor] = None):
        self.B, self.N, self.pitch = B, N, pitch_for(N)
        if store is None:
            store = torch.zeros((B, N, self.pitch), dtype=torch.float32, device=device or require_cuda())
        self.store = store

    @property
    def plane(self) -> int:
        return self.N * self.pitch

    @property
    def ptr(self) -> int:
        return self.store.data_ptr()

    @property
    def view(self) -> torch.Tensor:
        v = self.store[:, None, :, : self.N]
        v._mgfea_field = self  # lets as_field() recognise our own buffers without a copy
        return v

    def zero_(self):
        self.store.zero_()
        return self


def as_field(x: torch.Tensor, device=None) -> Field:
    """Accept a (B,1,N,N) / (B,N,N) / (N,N) tensor on any device and return a padded device Field.
    Our own strided views are used in place; anything else is copied (H2D if needed) and packed by mgfea_pack."""
    fld = getattr(x, "_mgfea_field", None)
    if fld is not None and x.is_cuda and x.data_ptr() == fld.ptr:
        return fld
    dev = device or require_cuda()
    if x.dim() == 4:
        if x.shape[1] != 1:
            raise MgfeaError(f"expected a single-channel field, got {tuple(x.shape)}")
        x3 = x[:, 0]
    elif x.dim() == 2:
        x3 = x[None]
    else:
        x3 = x
    B, N, N2 = x3.shape
    if N != N2:
        raise MgfeaError("fields must be square")
    xc = x3.detach().to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
    out = Field(B, N, dev, store=torch.empty((B, N, pitch_for(N)), dtype=torch.float32, device=dev))
    check(lib().mgfea_pack(xc.data_ptr(), out.ptr, N, out.pitch, out.plane, B, stream_ptr()))
    return out


def to_contiguous(fld: Field) -> torch.Tensor:
    out = torch.empty((fld.B, 1, fld.N, fld.N), dtype=torch.float32, device=fld.store.device)
    check(lib().mgfea_unpack(fld.ptr, out.data_ptr(), fld.N, fld.pitch, fld.plane, fld.B, stream_ptr()))
    return out


def pack_keys(keys_u8) -> torch.Tensor:
    """uint8 (N,N) numpy/torch -> device tensor [N][key_pitch] with key_pitch % 16 == 0 (zero padded)"""
    dev = require_cuda()
    k = torch.as_tensor(keys_u8, dtype=torch.uint8)
    N = k.shape[0]
    kp = (N + 127) // 128 * 128
    out = torch.zeros((N, kp), dtype=torch.uint8, device=dev)
    out[:, :N] = k.to(dev)
    return out


def device_pattern_keys(N: int, shape: int) -> torch.Tensor:
    """[N][key_pitch] uint8 key map of the two-phase plate generated on the device (mgfea_pattern_keys)"""
    dev = require_cuda()
    kp = (N + 127) // 128 * 128
    out = torch.empty((N, kp), dtype=torch.uint8, device=dev)
    check(lib().mgfea_pattern_keys(out.data_ptr(), N, kp, int(shape), stream_ptr()))
    return out


class _RawCuda:
    """__cuda_array_interface__ carrier: lets torch view memory owned by the C library (mgfea_peer_alloc) without a copy"""

    def __init__(self, ptr, shape, typestr, owner):
        self.owner = owner  # keeps the block alive while a tensor built on it exists
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerBlock:
    """A device block that the other ranks of the node can map (cudaIpc): slab arrays + the exchange mailbox live here.
    `base` is the local address; `tensor()` carves torch views out of it; `handle()` is what travels to the peers."""

    def __init__(self, nbytes: int):
        self.dev = require_cuda()
        self.nbytes = int(nbytes)
        p = ctypes.c_void_p()
        check(lib().mgfea_peer_alloc(ctypes.byref(p), self.nbytes))
        self.base = int(p.value)
        self._opened = []

    def handle(self) -> bytes:
        buf = ctypes.create_string_buffer(IPC_HANDLE_BYTES)
        check(lib().mgfea_peer_export(self.base, buf))
        return bytes(buf.raw)

    def open_peer(self, handle: bytes) -> int:
        """map another rank's block into this process; returns its base address here"""
        p = ctypes.c_void_p()
        check(lib().mgfea_peer_open(ctypes.create_string_buffer(handle, IPC_HANDLE_BYTES), ctypes.byref(p)))
        self._opened.append(int(p.value))
        return int(p.value)

    def tensor(self, offset: int, shape, dtype=torch.float32) -> torch.Tensor:
        typestr = {torch.float32: "<f4", torch.float64: "<f8", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
        nbytes = int(torch.empty((), dtype=dtype).element_size())
        for d in shape:
            nbytes *= int(d)
        if offset < 0 or offset + nbytes > self.nbytes:
            raise MgfeaError("PeerBlock.tensor: region outside the block")
        return torch.as_tensor(_RawCuda(self.base + offset, shape, typestr, self), device=self.dev)

    def close(self):
        for p in self._opened:
            lib().mgfea_peer_close(p)
        self._opened = []
        if self.base:
            lib().mgfea_peer_free(self.base)
            self.base = 0


class DeviceTable:
    """Device copy of a small live parameter tensor, refreshed when the parameter changes (load_state_dict, in-place
    edits bump torch's version counter), so the kernels always see the current weights."""

    def __init__(self):
        self._src_id = None
        self._dev = None

    def get(self, t: torch.Tensor) -> torch.Tensor:
        t = t.detach()
        if t.is_cuda and t.dtype == torch.float32 and t.is_contiguous():
            return t
        sid = (t.data_ptr(), t._version, tuple(t.shape))
        if sid != self._src_id:
            self._dev = t.to(device=require_cuda(), dtype=torch.float32).contiguous()
            self._src_id = sid
        return self._dev
