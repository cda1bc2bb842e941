"""Structured quad meshes of the reference in closed form (reference: FEANet/mesh.py).

The reference builds the two-phase pattern map with Python loops that are O(N^4) (mesh.py:78-101: one
``np.where(cells == pid)`` per node), which takes seconds at 129^2 and cannot produce 4097^2.  Here the same integer
masks come from closed-form integer arithmetic (bit-identical to the reference for every size it can run; pinned by
tests/golden/mesh.npz), in O(N^2) vectorised numpy.  Attribute contract kept: ``nnode_edge``, ``kernel_dict``
{key: (3,3) float32}, ``global_pattern_center`` {key: int[N*N]}, ``a``, ``Ke``, ``ref_pattern_dict``, ``phase``.
"""
from collections.abc import Mapping

import numpy as np

_REF_PATTERNS = {0: [0, 0, 0, 0], 1: [1, 1, 1, 1], 2: [0, 0, 0, 1], 3: [0, 0, 1, 0],
                 4: [1, 0, 0, 0], 5: [0, 1, 0, 0], 6: [0, 0, 1, 1], 7: [1, 1, 0, 0],
                 8: [0, 1, 1, 0], 9: [1, 0, 0, 1], 10: [0, 1, 0, 1], 11: [1, 0, 1, 0],
                 12: [1, 1, 1, 0], 13: [1, 1, 0, 1], 14: [0, 1, 1, 1], 15: [1, 0, 1, 1]}


def _element_stiffness():
    # Q1 Laplacian on the unit square (mesh.py:28-31), h-independent
    return -1. / 6. * np.array([[-4., 1., 2., 1.], [1., -4., 1., 2.], [2., 1., -4., 1.], [1., 2., 1., -4.]],
                               dtype=np.float32)


def _node_kernel(a, Ke, pat):
    """3x3 nodal stencil of a node whose four elements [e1,e2,e3,e4] have phases `pat` (mesh.py:103-117).
    All products and sums are fp32 and in the reference's order, so the table is bit-identical."""
    k = np.zeros((3, 3), dtype=np.float32)
    e1, e2, e3, e4 = (a[pat[0]], a[pat[1]], a[pat[2]], a[pat[3]])
    k[0, 0] = e4 * Ke[1, 3]
    k[0, 1] = e4 * Ke[1, 2] + e3 * Ke[0, 3]
    k[0, 2] = e3 * Ke[0, 2]
    k[1, 0] = e1 * Ke[2, 3] + e4 * Ke[1, 0]
    k[1, 1] = e3 * Ke[0, 0] + e4 * Ke[1, 1] + e1 * Ke[2, 2] + e2 * Ke[3, 3]
    k[1, 2] = e2 * Ke[3, 2] + e3 * Ke[0, 1]
    k[2, 0] = e1 * Ke[2, 0]
    k[2, 1] = e1 * Ke[2, 1] + e2 * Ke[3, 0]
    k[2, 2] = e2 * Ke[3, 1]
    return k


class _LazyPatternMap(Mapping):
    """``global_pattern_center``: {key: int[N*N] one-hot of nodes with that pattern}, materialised per key on access
    (16 dense int64 maps of a 4097^2 grid would be 2 GB; the kernels use the uint8 key map instead)."""

    def __init__(self, keys_u8, npat):
        # keys_u8: the (N,N) uint8 key map, or an object whose ``pattern_keys`` attribute produces it on demand
        self._src, self._npat = keys_u8, npat

    @property
    def _keys(self):
        return self._src if isinstance(self._src, np.ndarray) else self._src.pattern_keys

    def __getitem__(self, k):
        if not (0 <= k < self._npat):
            raise KeyError(k)
        return (self._keys.reshape(-1) == k).astype(int)

    def __iter__(self):
        return iter(range(self._npat))

    def __len__(self):
        return self._npat


class _MeshContainer:
    """stand-in for the meshio.Mesh object the reference keeps (container only: mesh.py:60,68)"""

    def __init__(self, points, cells):
        self.points, self.cells, self.cell_data = points, cells, {}

    def write(self, *_a, **_k):
        raise NotImplementedError("mesh file output (meshio) is outside the solve() path")


class _StructuredQuad:
    def _lazy_geometry(self):
        n1 = self.nnode_edge
        x = np.linspace(self.size / 2, -self.size / 2, n1, dtype=np.float32)
        y = np.linspace(-self.size / 2, self.size / 2, n1, dtype=np.float32)
        mx, my = np.meshgrid(x, y)
        pts = np.concatenate((mx.reshape(-1, 1), my.reshape(-1, 1), np.zeros((n1 * n1, 1), np.float32)), axis=1)
        nodes = np.arange(n1 * n1).reshape(n1, n1)
        cells = np.stack([nodes[:-1, :-1].ravel(), nodes[:-1, 1:].ravel(), nodes[1:, 1:].ravel(),
                          nodes[1:, :-1].ravel()], axis=1)
        return pts, cells

    @property
    def points(self):
        return self._lazy_geometry()[0]

    @property
    def cells(self):
        return self._lazy_geometry()[1]

    @property
    def mesh(self):
        m = _MeshContainer(*self._lazy_geometry())
        m.cell_data['Phase'] = self.phase
        return m

    def save_mesh(self, outfile=None):
        raise NotImplementedError("mesh file output (meshio) is outside the solve() path")


class MeshSquare(_StructuredQuad):
    """Homogeneous square plate, one pattern (reference: FEANet/mesh.py:122-192)."""

    def __init__(self, size=2, nnode_edge=65, outfile=None):
        self.size, self.nnode_edge = size, nnode_edge
        self.a = np.array([1.], dtype=np.float32)
        self.ref_pattern_dict = {0: [0, 0, 0, 0]}
        self.Ke = _element_stiffness()
        self.kernel_dict = {0: _node_kernel(self.a, self.Ke, [0, 0, 0, 0])}
        self.pattern_keys = None  # single pattern: no key map needed
        self.global_pattern_center = _LazyPatternMap(np.zeros((nnode_edge, nnode_edge), np.uint8), 1)
        if outfile is not None:
            self.save_mesh(outfile)

    @property
    def phase(self):
        return np.zeros(((self.nnode_edge - 1) ** 2,), dtype=int)


class MeshCenterInterface(_StructuredQuad):
    """Square plate with a central inclusion (circle r=0.5, shape=0; square half-width 0.5, shape=1), two phases
    (reference: FEANet/mesh.py:4-120).  ``pattern_keys`` is the uint8 (N,N) key map the kernels consume."""

    def __init__(self, size=2, prop=[1, 20], nnode_edge=65, shape=0, outfile=None):
        if size != 2:
            # the closed-form inclusion test (and mgfea_pattern_keys) is the reference's centroid test for the plate
            # [-1, 1]^2 with radius / half-width 0.5; the reference derives centroids from `size`, so any other size
            # would silently give a DIFFERENT inclusion than the reference does
            raise ValueError("MeshCenterInterface: only size=2 (every driver of the reference) is supported")
        self.size, self.nnode_edge, self.shape = size, nnode_edge, shape
        self.a = np.array(prop, dtype=np.float32)
        self.ref_pattern_dict = {k: list(v) for k, v in _REF_PATTERNS.items()}
        self.Ke = _element_stiffness()
        self.kernel_dict = {k: _node_kernel(self.a, self.Ke, p) for k, p in self.ref_pattern_dict.items()}
        # host copies of the element phases / node keys are built on first access: the kernels take the key map from
        # mgfea_pattern_keys (device_pattern_keys), so a 16385^2 mesh costs no host-side 268M-element passes
        self._phase2d_cache = None
        self._keys_cache = None
        self.global_pattern_center = _LazyPatternMap(self, 16)
        if outfile is not None:
            self.save_mesh(outfile)

    @staticmethod
    def _element_phase(n, shape):
        """phase of element (r,c): centroid strictly inside the inclusion (mesh.py:62-76), integer form.
        circle: 4((2c+1-n)^2 + (2r+1-n)^2) < n^2 ; square: 2|2c+1-n| < n and 2|2r+1-n| < n  """
        if shape not in (0, 1):
            return np.zeros((n, n), np.uint8)
        t = 2 * np.arange(n, dtype=np.int64) + 1 - n
        if shape == 0:
            return (4 * (t[None, :] ** 2 + t[:, None] ** 2) < n * n).astype(np.uint8)
        inside = 2 * np.abs(t) < n
        return (inside[None, :] & inside[:, None]).astype(np.uint8)

    @staticmethod
    def _node_keys(ph):
        """pattern [e1,e2,e3,e4] = phases of elements (i-1,j),(i-1,j-1),(i,j-1),(i,j) of interior node (i,j)
        (mesh.py:78-93 with x decreasing along j); boundary nodes keep key 0 (mesh.py:81-82)."""
        n = ph.shape[0]
        N = n + 1
        e1, e2, e3, e4 = ph[:-1, 1:], ph[:-1, :-1], ph[1:, :-1], ph[1:, 1:]
        code = (e1.astype(np.int64) << 3) | (e2 << 2) | (e3 << 1) | e4
        lut = np.zeros(16, np.uint8)
        for k, p in _REF_PATTERNS.items():
            lut[(p[0] << 3) | (p[1] << 2) | (p[2] << 1) | p[3]] = k
        keys = np.zeros((N, N), np.uint8)
        keys[1:-1, 1:-1] = lut[code]
        return keys

    @property
    def _phase2d(self):
        if self._phase2d_cache is None:
            self._phase2d_cache = self._element_phase(self.nnode_edge - 1, self.shape)
        return self._phase2d_cache

    @property
    def pattern_keys(self):
        if self._keys_cache is None:
            self._keys_cache = self._node_keys(self._phase2d)
        return self._keys_cache

    def device_pattern_keys(self):
        """the uint8 key map as a padded device tensor, generated by the setup kernel (no host pass)"""
        import mgfea

        return mgfea.device_pattern_keys(self.nnode_edge, self.shape)

    @property
    def phase(self):
        return self._phase2d.reshape(-1).astype(int)

    @property
    def pattern(self):
        pat = np.array([_REF_PATTERNS[k] for k in range(16)], dtype=int)
        return pat[self.pattern_keys.reshape(-1)]
