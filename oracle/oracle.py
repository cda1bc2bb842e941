"""numpy front-end of the C oracle (oracle/mgfea_oracle.c) + restated V-cycle drivers.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (multigrid-feanet_b200/) never imports it.

Parity status: PINNED against the unmodified reference through tests/golden/ (see
tests/golden/make_golden.py and tests/test_oracle_golden.py).

Restated drivers (reference file:line, relative to /root/reference):
  vcycle(...)  MM_Model_convergence.ipynb cell 3 `Multigrid.rec_V_cycle` / `V_cycle`   (variant "A")
               M-FEANet-mg_test.ipynb cell 19 `MultiGrid.Step`, FEANet/multigrid.py:159-185 `iterate` (variant "B")
               MM_Interface_error.ipynb cell 2 `rec_V_cycle` (pre-smooth always on level 0: `quirk_level0=True`)
  solve(...)   MM_Model_convergence.ipynb cell 3 `Multigrid.Solve`
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f32p = ctypes.POINTER(ctypes.c_float)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> str:
    out = os.path.join(_HERE, "_build", "libmgfea_oracle.so")
    src = os.path.join(_HERE, "mgfea_oracle.c")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return out


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _p(a, ty):
    if a is None:
        return ctypes.cast(None, ty)
    return a.ctypes.data_as(ty)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _as3(a):
    a = _f(a)
    if a.ndim == 2:
        a = a[None]
    if a.ndim == 4:
        assert a.shape[1] == 1
        a = a[:, 0]
    return np.ascontiguousarray(a)


def _bc(idx, bval, B, N):
    """returns (idx_arr, bval_arr, batch_stride)"""
    if idx is None:
        return None, None, 0
    idx = _as3(idx)
    bval = _as3(bval if bval is not None else np.zeros_like(idx))
    assert idx.shape == bval.shape and idx.shape[1:] == (N, N)
    stride = N * N if idx.shape[0] > 1 else 0
    assert idx.shape[0] in (1, B)
    return idx, bval, stride


def _keys(keys):
    return None if keys is None else np.ascontiguousarray(keys, dtype=np.uint8)


def stiffness_apply(u, keys, ktab):
    u = _as3(u)
    B, N, _ = u.shape
    out = np.empty_like(u)
    keys = _keys(keys)
    ktab = _f(ktab).reshape(-1, 9)
    lib().orc_stiffness_apply(_p(u, _f32p), _p(out, _f32p), _p(keys, _u8p), _p(ktab, _f32p), N, B)
    return out


def conv3x3(u, w9):
    """FNet / single-table zero-padded 3x3 correlation (FEANet/model.py:49-61)."""
    return stiffness_apply(u, None, _f(w9).reshape(1, 9))


def split_x(x, keys, C):
    x = _as3(x)
    B, N, _ = x.shape
    out = np.empty((B, C, N, N), np.float32)
    keys = _keys(keys)
    lib().orc_split_x(_p(x, _f32p), _p(out, _f32p), _p(keys, _u8p), C, N, B)
    return out


def reset_boundary(u, idx=None, bval=None):
    u = _as3(u)
    B, N, _ = u.shape
    idx, bval, bs = _bc(idx, bval, B, N)
    out = np.empty_like(u)
    lib().orc_reset_boundary(_p(u, _f32p), _p(out, _f32p), _p(idx, _f32p), _p(bval, _f32p), ctypes.c_size_t(bs), N, B)
    return out


def inv_diag(omega, dtab):
    """per-key omega/d exactly as torch evaluates `self.omega/self.d_mat` (jacobi.py:46): Tensor.__rtruediv__ is
    `d_mat.reciprocal() * omega`, i.e. fl32(fl32(1/d) * fl32(omega))."""
    return ((np.float32(1.0) / _f(dtab)).astype(np.float32) * np.float32(omega)).astype(np.float32)


def jacobi(u, f, keys, ktab, invd, idx=None, bval=None, nsweeps=1):
    u, f = _as3(u), _as3(f)
    B, N, _ = u.shape
    idx, bval, bs = _bc(idx, bval, B, N)
    out = np.empty_like(u)
    keys = _keys(keys)
    ktab = _f(ktab).reshape(-1, 9)
    invd = _f(invd).reshape(-1)
    lib().orc_jacobi(_p(u, _f32p), _p(out, _f32p), _p(f, _f32p), _p(keys, _u8p), _p(ktab, _f32p), _p(invd, _f32p),
                     _p(idx, _f32p), _p(bval, _f32p), ctypes.c_size_t(bs), N, B, nsweeps)
    return out


def hjacobi(u, f, keys, ktab, invd, hw, idx=None, bval=None, nsweeps=1):
    u, f = _as3(u), _as3(f)
    B, N, _ = u.shape
    idx, bval, bs = _bc(idx, bval, B, N)
    out = np.empty_like(u)
    keys = _keys(keys)
    ktab = _f(ktab).reshape(-1, 9)
    invd = _f(invd).reshape(-1)
    hw = _f(hw).reshape(-1, 9)
    lib().orc_hjacobi(_p(u, _f32p), _p(out, _f32p), _p(f, _f32p), _p(keys, _u8p), _p(ktab, _f32p), _p(invd, _f32p),
                      _p(idx, _f32p), _p(bval, _f32p), ctypes.c_size_t(bs), _p(hw, _f32p), hw.shape[0], N, B, nsweeps)
    return out


def jacobi_pbc(u, f_pad, w9, invd, nsweeps=1):
    """JacobiBlockPBC.jacobi_convolution (FEANet/jacobi.py:89-97); f_pad (B, N+2, N+2) is the caller-padded load vector"""
    u, f_pad = _as3(u), _as3(f_pad)
    B, N, _ = u.shape
    assert f_pad.shape == (B, N + 2, N + 2)
    out = np.empty_like(u)
    lib().orc_jacobi_pbc(_p(u, _f32p), _p(out, _f32p), _p(f_pad, _f32p), _p(_f(w9).reshape(9), _f32p),
                         ctypes.c_float(float(np.float32(invd))), N, B, nsweeps)
    return out


def residual(u, f, keys, ktab):
    u, f = _as3(u), _as3(f)
    B, N, _ = u.shape
    out = np.empty_like(u)
    keys = _keys(keys)
    ktab = _f(ktab).reshape(-1, 9)
    lib().orc_residual(_p(u, _f32p), _p(f, _f32p), _p(out, _f32p), _p(keys, _u8p), _p(ktab, _f32p), N, B)
    return out


def restrict(r, keys, rtab, scale=None):
    r = _as3(r)
    B, N, _ = r.shape
    Nc = (N - 1) // 2 + 1
    out = np.empty((B, Nc, Nc), np.float32)
    keys = _keys(keys)
    rtab = _f(rtab).reshape(-1, 9)
    lib().orc_restrict(_p(r, _f32p), _p(out, _f32p), _p(keys, _u8p), _p(rtab, _f32p),
                       ctypes.c_float(0.0 if scale is None else float(np.float32(scale))), int(scale is not None), N, B)
    return out


def prolong_bilinear(vc, u, idx=None, bval=None):
    """returns u + BC_fine(bilinear(vc))"""
    vc, u = _as3(vc), _as3(u).copy()
    B, N, _ = u.shape
    idx, bval, bs = _bc(idx, bval, B, N)
    lib().orc_prolong_bilinear(_p(vc, _f32p), _p(u, _f32p), _p(idx, _f32p), _p(bval, _f32p), ctypes.c_size_t(bs), N, B,
                               int(N <= 33))
    return u


def prolong_table(vc, u, keys_c, ptab, scale=None):
    """returns u + scale*convT(P)(vc)"""
    vc, u = _as3(vc), _as3(u).copy()
    B, N, _ = u.shape
    keys_c = _keys(keys_c)
    ptab = _f(ptab).reshape(-1, 9)
    lib().orc_prolong_table(_p(vc, _f32p), _p(u, _f32p), _p(keys_c, _u8p), _p(ptab, _f32p),
                            ctypes.c_float(0.0 if scale is None else float(np.float32(scale))), int(scale is not None), N, B)
    return u


# ----------------------------------------------------------------------------------------------
# general per-element conductivity: see the block comment in mgfea_oracle.c
# ----------------------------------------------------------------------------------------------
KE = (-1.0 / 6.0 * np.array([[-4., 1., 2., 1.], [1., -4., 1., 2.], [2., 1., -4., 1.], [1., 2., 1., -4.]],
                            dtype=np.float32)).astype(np.float32)  # FEANet/mesh.py:28-31, same expression


def _elem(a, N):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.shape == (N - 1, N - 1), (a.shape, N)
    return a


def elem_stiffness_apply(u, a):
    u = _as3(u)
    B, N, _ = u.shape
    out = np.empty_like(u)
    lib().orc_elem_stiffness_apply(_p(u, _f32p), _p(out, _f32p), _p(_elem(a, N), _f32p), _p(KE, _f32p), N, B)
    return out


def elem_residual(u, f, a):
    u, f = _as3(u), _as3(f)
    B, N, _ = u.shape
    out = np.empty_like(u)
    lib().orc_elem_residual(_p(u, _f32p), _p(f, _f32p), _p(out, _f32p), _p(_elem(a, N), _f32p), _p(KE, _f32p), N, B)
    return out


def elem_diag(a):
    N = a.shape[0] + 1
    d = np.empty((N, N), np.float32)
    lib().orc_elem_diag(_p(d, _f32p), _p(_elem(a, N), _f32p), _p(KE, _f32p), N)
    return d


def elem_jacobi(u, f, a, omega=2.0 / 3.0, idx=None, bval=None, nsweeps=1):
    u, f = _as3(u), _as3(f)
    B, N, _ = u.shape
    idx, bval, bs = _bc(idx, bval, B, N)
    out = np.empty_like(u)
    lib().orc_elem_jacobi(_p(u, _f32p), _p(out, _f32p), _p(f, _f32p), _p(_elem(a, N), _f32p), _p(KE, _f32p),
                          ctypes.c_float(float(np.float32(omega))), _p(idx, _f32p), _p(bval, _f32p), ctypes.c_size_t(bs),
                          N, B, nsweeps)
    return out


def coarsen_elements(a):
    """conductivity of a coarse element = mean of its four children, summed in fp32 in row-major order (our convention:
    the reference rediscretises its two-phase inclusion on every level instead, FEANet/multigrid.py:19-29)"""
    a = np.ascontiguousarray(a, dtype=np.float32)
    s = (a[0::2, 0::2] + a[0::2, 1::2]).astype(np.float32)
    s = (s + a[1::2, 0::2]).astype(np.float32)
    s = (s + a[1::2, 1::2]).astype(np.float32)
    return (s * np.float32(0.25)).astype(np.float32)


def phase_map(N, shape=0):
    """element phases (n x n, 0/1) of the reference's centred inclusion, from the bit-pinned pattern keys: the SE
    element of node (i, j) is e4 of its pattern"""
    keys = pattern_keys(N, shape)
    e4 = np.array([REF_PATTERNS[k][3] for k in range(16)], dtype=np.int64)
    ph = e4[keys[:-1, :-1].astype(np.int64)]
    ph[0, :] = 0  # nodes on the ring carry key 0 whatever their elements are; the inclusion never touches the edge
    ph[:, 0] = 0
    return ph


def elem_vcycle(a_levels, u, f, nu1=1, nu2=1, omega=2.0 / 3.0):
    """V(nu1,nu2) with the per-element operator on every level (a_levels[l]: (n_l x n_l)), full weighting x4 and bilinear
    prolongation (MM_Model_convergence.ipynb cell 3 skeleton)"""
    L = len(a_levels)
    us, fs = [None] * L, [None] * L
    us[0], fs[0] = _as3(u), _as3(f)
    B = us[0].shape[0]
    for l in range(L):
        N = a_levels[l].shape[0] + 1
        if l > 0:
            us[l] = np.zeros((B, N, N), np.float32)
        if nu1 > 0:
            us[l] = elem_jacobi(us[l], fs[l], a_levels[l], omega, nsweeps=nu1)
        if l < L - 1:
            fs[l + 1] = restrict(elem_residual(us[l], fs[l], a_levels[l]), None, FW16, 4.0)
    for l in range(L - 1, -1, -1):
        if l < L - 1:
            us[l] = prolong_bilinear(us[l + 1], us[l])
        if nu2 > 0:
            us[l] = elem_jacobi(us[l], fs[l], a_levels[l], omega, nsweeps=nu2)
    return us[0]


def sumsq_interior(r):
    r = _as3(r)
    B, N, _ = r.shape
    out = np.empty(B, np.float64)
    lib().orc_sumsq_interior(_p(r, _f32p), _p(out, _f64p), N, B)
    return out


def pattern_keys(N, shape=0):
    keys = np.empty((N, N), np.uint8)
    lib().orc_pattern_keys(_p(keys, _u8p), N, shape)
    return keys


# ----------------------------------------------------------------------------------------------
# problem setup restated (FEANet/mesh.py:28-31,103-117; FEANet/model.py:54-56)
# ----------------------------------------------------------------------------------------------
REF_PATTERNS = {0: [0, 0, 0, 0], 1: [1, 1, 1, 1], 2: [0, 0, 0, 1], 3: [0, 0, 1, 0], 4: [1, 0, 0, 0], 5: [0, 1, 0, 0],
                6: [0, 0, 1, 1], 7: [1, 1, 0, 0], 8: [0, 1, 1, 0], 9: [1, 0, 0, 1], 10: [0, 1, 0, 1], 11: [1, 0, 1, 0],
                12: [1, 1, 1, 0], 13: [1, 1, 0, 1], 14: [0, 1, 1, 1], 15: [1, 0, 1, 1]}


def kernel_table(prop, npat):
    """(npat, 3, 3) f32 kernels; every product/sum evaluated in fp32 in the reference's order."""
    a = np.array(prop, dtype=np.float32)
    Ke = -1.0 / 6.0 * np.array([[-4., 1., 2., 1.], [1., -4., 1., 2.], [2., 1., -4., 1.], [1., 2., 1., -4.]],
                               dtype=np.float32)
    out = np.zeros((npat, 3, 3), np.float32)
    for k in range(npat):
        p = REF_PATTERNS[k]
        kern = out[k]
        kern[0, 0] = a[p[3]] * Ke[1, 3]
        kern[0, 1] = a[p[3]] * Ke[1, 2] + a[p[2]] * Ke[0, 3]
        kern[0, 2] = a[p[2]] * Ke[0, 2]
        kern[1, 0] = a[p[0]] * Ke[2, 3] + a[p[3]] * Ke[1, 0]
        kern[1, 1] = a[p[2]] * Ke[0, 0] + a[p[3]] * Ke[1, 1] + a[p[0]] * Ke[2, 2] + a[p[1]] * Ke[3, 3]
        kern[1, 2] = a[p[1]] * Ke[3, 2] + a[p[2]] * Ke[0, 1]
        kern[2, 0] = a[p[0]] * Ke[2, 0]
        kern[2, 1] = a[p[0]] * Ke[2, 1] + a[p[1]] * Ke[3, 0]
        kern[2, 2] = a[p[1]] * Ke[3, 1]
    return out


def load_vector_weights(h):
    return np.array([[h * h / 36., h * h / 9., h * h / 36.], [h * h / 9., 4. * h * h / 9., h * h / 9.],
                     [h * h / 36., h * h / 9., h * h / 36.]], dtype=np.float32)


@dataclass
class Level:
    N: int
    keys: Optional[np.ndarray]  # uint8 (N,N) or None
    ktab: np.ndarray  # (C,9)
    invd: np.ndarray  # (C,)
    idx: Optional[np.ndarray] = None  # general BC masks (level 0 only in the reference)
    bval: Optional[np.ndarray] = None


def make_levels(n, L=None, prop=None, shape=0, omega=2.0 / 3.0) -> List[Level]:
    """levels N_l = n/2^l + 1, l < L (L = log2 n by default: MM_Model_convergence.ipynb cell 3 __init__)"""
    if L is None:
        L = int(np.log2(n))
    lv = []
    for l in range(L):
        N = int(n / (2.0 ** l)) + 1
        if prop is None:
            tab = kernel_table([1.0], 1).reshape(1, 9)
            keys = None
        else:
            tab = kernel_table(prop, 16).reshape(16, 9)
            keys = pattern_keys(N, shape)
        lv.append(Level(N, keys, tab, inv_diag(omega, tab[:, 4])))
    return lv


@dataclass
class CycleCfg:
    nu1: int = 1
    nu2: int = 1
    smoother: str = "jac"  # 'jac' | 'hjac'
    hw: Optional[np.ndarray] = None  # (nlayers, 9)
    prolong: str = "bilinear"  # 'bilinear' (variant A) | 'table' (variant B)
    rtab: Optional[np.ndarray] = None  # (C,9) or (1,9); default full weighting /16
    r_scale: Optional[float] = 4.0  # None: no multiply
    ptab: Optional[np.ndarray] = None
    p_scale: Optional[float] = None
    quirk_level0: bool = False  # MM_Interface_error.ipynb cell 2: pre-smooth always applied to level 0
    fields: dict = field(default_factory=dict)


FW16 = (np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=np.float32) / np.float32(16.0)).reshape(1, 9)
LIN4 = (np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=np.float32) / np.float32(4.0)).reshape(1, 9)


def _relax(lv: Level, cfg: CycleCfg, u, f, k):
    if k <= 0:
        return u
    if cfg.smoother == "jac":
        return jacobi(u, f, lv.keys, lv.ktab, lv.invd, lv.idx, lv.bval, k)
    return hjacobi(u, f, lv.keys, lv.ktab, lv.invd, cfg.hw, lv.idx, lv.bval, k)


def _tab_for(tab, lv):
    tab = _f(tab).reshape(-1, 9)
    if lv.keys is not None and tab.shape[0] == 1:
        tab = np.repeat(tab, lv.ktab.shape[0], axis=0)
    return tab


def vcycle(levels: List[Level], cfg: CycleCfg, u, f):
    """one V(nu1,nu2) cycle; u,f (B,N,N) on level 0; returns new u (B,N,N)."""
    L = len(levels)
    us = [None] * L
    fs = [None] * L
    us[0], fs[0] = _as3(u), _as3(f)
    B = us[0].shape[0]
    rtab = FW16 if cfg.rtab is None else cfg.rtab
    ptab = LIN4 if cfg.ptab is None else cfg.ptab
    for l in range(L):
        if l > 0:
            us[l] = np.zeros((B, levels[l].N, levels[l].N), np.float32)
        if cfg.quirk_level0:
            us[0] = _relax(levels[0], cfg, us[0], fs[0], cfg.nu1)
        else:
            us[l] = _relax(levels[l], cfg, us[l], fs[l], cfg.nu1)
        if l < L - 1:
            r = residual(us[l], fs[l], levels[l].keys, levels[l].ktab)
            fs[l + 1] = restrict(r, levels[l].keys, _tab_for(rtab, levels[l]), cfg.r_scale)
    for l in range(L - 1, -1, -1):
        if l < L - 1:
            if cfg.prolong == "bilinear":
                us[l] = prolong_bilinear(us[l + 1], us[l], levels[l].idx, levels[l].bval)
            else:
                us[l] = prolong_table(us[l + 1], us[l], levels[l + 1].keys, _tab_for(ptab, levels[l + 1]), cfg.p_scale)
        us[l] = _relax(levels[l], cfg, us[l], fs[l], cfg.nu2)
    return us[0]


def residual_norm(levels, u, f):
    """per-sample interior 2-norm of f - K u on level 0 (float64 accumulate)"""
    r = residual(u, f, levels[0].keys, levels[0].ktab)
    return np.sqrt(sumsq_interior(r))


def solve(levels, cfg, u0, f, n_iter=None, EPS=None, max_cycles=200):
    """Multigrid.Solve semantics: repeat while (res > EPS or n < n_iter); res = whole-batch interior 2-norm.
    Returns (u, [res per cycle])."""
    if n_iter is None:
        assert EPS is not None
        n_iter = 0
    elif EPS is None:
        EPS = np.inf
    u = _as3(u0)
    f = _as3(f)
    hist = []
    res = 1.0
    n = 0
    while (res > EPS or n < n_iter) and n < max_cycles:
        u = vcycle(levels, cfg, u, f)
        res = float(np.sqrt(np.sum(sumsq_interior(residual(u, f, levels[0].keys, levels[0].ktab)))))
        hist.append(res)
        n += 1
    return u, hist
