"""Harness that imports the UNMODIFIED reference (/root/reference) in THIS container.

Used only by ``tests/golden/make_golden.py`` to generate the committed golden
vectors.  Nothing under ``tests/`` that runs on the GPU box imports this file
(``/root/reference`` does not exist there).

Shims (none touches reference source):
  * ``meshio``      -- absent here; FEANet/mesh.py:2,60,169 only uses ``meshio.Mesh`` as a container.
  * ``matplotlib``  -- absent; Utils/plot.py:2 imports it, notebooks import ``Utils.plot``.
  * ``h5py``        -- absent; Data/dataset.py:1 (only needed when notebook import cells are exec'd).
  * notebook classes are obtained by ``exec``-ing the code cells of the ``.ipynb`` JSON.
"""
import json
import os
import struct
import sys
import types

import numpy as np

REF = os.environ.get("MGFEA_REFERENCE", "/root/reference")


def _install_stubs():
    if "meshio" not in sys.modules:
        m = types.ModuleType("meshio")

        class Mesh:  # container only
            def __init__(self, points, cells):
                self.points, self.cells, self.cell_data = points, cells, {}

            def write(self, *_a, **_k):
                raise RuntimeError("meshio stub: write unsupported")

        m.Mesh = Mesh
        sys.modules["meshio"] = m
    for name in ("matplotlib", "matplotlib.pyplot", "h5py", "torchvision", "torchvision.transforms"):
        if name in sys.modules:
            continue
        try:
            __import__(name)
            continue
        except Exception:
            pass
        mod = types.ModuleType(name)

        def _noop(*_a, **_k):
            return None

        def _ga(attr, _n=_noop):
            if attr.startswith("__"):
                raise AttributeError(attr)
            return _n

        mod.__getattr__ = _ga  # type: ignore[attr-defined]
        mod.__file__ = f"<stub {name}>"
        sys.modules[name] = mod
    if "matplotlib" in sys.modules and "matplotlib.pyplot" in sys.modules:
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def load_reference():
    """Put the reference on sys.path (first) and return its FEANet package modules."""
    if not os.path.isdir(REF):
        raise RuntimeError(f"reference tree not found at {REF}")
    _install_stubs()
    # the product package is also called FEANet: make sure the REFERENCE one wins here
    for k in [k for k in sys.modules if k == "FEANet" or k.startswith("FEANet.")]:
        del sys.modules[k]
    if REF in sys.path:
        sys.path.remove(REF)
    sys.path.insert(0, REF)
    import FEANet.geo, FEANet.jacobi, FEANet.mesh, FEANet.model  # noqa: E401

    assert FEANet.__file__ is None or REF in os.path.abspath(FEANet.model.__file__)
    return sys.modules["FEANet"]


def notebook_namespace(nb_name, cells, extra=None):
    """exec the given code cells of a reference notebook into a fresh namespace."""
    load_reference()
    nb = json.load(open(os.path.join(REF, nb_name)))
    ns = {"__name__": "__ref_nb__"}
    if extra:
        ns.update(extra)
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        for i in cells:
            c = nb["cells"][i]
            assert c["cell_type"] == "code", (nb_name, i)
            exec(compile("".join(c["source"]), f"{nb_name}[cell {i}]", "exec"), ns)
    finally:
        os.chdir(cwd)
    return ns


def load_reference_multigrid_module():
    """FEANet/multigrid.py with the n_iter shim of SURVEY section 0 (multigrid.py:46 passes
    n_iter to a 2-argument jacobi_convolution): wrap the committed method in a loop."""
    load_reference()
    import FEANet.jacobi as J

    if not getattr(J.JacobiBlock, "_mgfea_shim", False):
        orig = J.JacobiBlock.jacobi_convolution

        def jacobi_convolution(self, initial_u, forcing_term, n_iter=1):
            u = initial_u
            for _ in range(n_iter):
                u = orig(self, u, forcing_term)
            return u

        J.JacobiBlock.jacobi_convolution = jacobi_convolution
        J.JacobiBlock._mgfea_shim = True
    import FEANet.multigrid as M

    return M


def read_h5_contiguous(path, shape, dtype="<f8"):
    """Mini reader for the reference's contiguous, uncompressed little-endian HDF5 payloads
    (SURVEY App. B.2): finds `03 01 <addr:u64> <size:u64>` contiguous-layout messages whose
    size matches `shape` and returns the blocks in file order of their addresses."""
    d = open(path, "rb").read()
    want = int(np.prod(shape)) * np.dtype(dtype).itemsize
    found = {}
    pos = 0
    while True:
        pos = d.find(b"\x03\x01", pos)
        if pos < 0:
            break
        if pos + 18 <= len(d):
            addr, size = struct.unpack("<QQ", d[pos + 2 : pos + 18])
            if size == want and 0 < addr and addr + size <= len(d):
                found[addr] = np.frombuffer(d, dtype=dtype, count=int(np.prod(shape)), offset=addr).reshape(shape)
        pos += 1
    return [found[a] for a in sorted(found)]
