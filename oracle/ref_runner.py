"""Runs the UNMODIFIED reference classes from oracle/_ref/ (vendored by oracle/vendor_ref.py) on the host CPU.

TEST INFRASTRUCTURE ONLY: used by `bench.py --impl reference` / bench.py's cpu_baseline leg.  Nothing in the product
imports this.  Shims (none touches reference source): `meshio` (container only, FEANet/mesh.py:2),
`matplotlib` / `h5py` / `torchvision` stubs for the notebooks' import cells, the notebook classes are obtained by
exec-ing the code cells of the .ipynb JSON, and `random_data`'s NumPy-2 float64 promotion is cast back to fp32 by the
caller (MM_Model_convergence.ipynb cell 3).
"""
import contextlib
import io
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF, "MANIFEST.json")) and os.path.exists(os.path.join(REF, "FEANet", "model.py"))


def _install_stubs():
    if "meshio" not in sys.modules:
        m = types.ModuleType("meshio")

        class Mesh:  # container only
            def __init__(self, points, cells):
                self.points, self.cells, self.cell_data = points, cells, {}

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

        def _ga(attr):
            if attr.startswith("__"):
                raise AttributeError(attr)
            return lambda *a, **k: None

        mod.__getattr__ = _ga
        mod.__file__ = f"<stub {name}>"
        sys.modules[name] = mod
    if "matplotlib" in sys.modules and "matplotlib.pyplot" in sys.modules:
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def load():
    """put oracle/_ref first on sys.path so that `FEANet` is the REFERENCE package (the product package has the same
    name: this process must not have imported it)"""
    if not available():
        raise RuntimeError("oracle/_ref is empty: run oracle/vendor_ref.py where /root/reference exists")
    _install_stubs()
    loaded = sys.modules.get("FEANet")
    if loaded is not None and REF not in os.path.abspath(getattr(loaded, "__file__", "") or ""):
        raise RuntimeError("the product FEANet package is already imported in this process")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import FEANet.model  # noqa: F401

    assert REF in os.path.abspath(sys.modules["FEANet.model"].__file__)


def notebook_namespace(nb_name, cells):
    load()
    nb = json.load(open(os.path.join(REF, nb_name)))
    ns = {"__name__": "__ref_nb__"}
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


def model_problem(n, u0):
    """the reference's own benchmark: MM_Model_convergence.ipynb cell 3 `Multigrid(n)`, f = 0, initial_v = u0 (fp32)"""
    import torch

    ns = notebook_namespace("MM_Model_convergence.ipynb", [1, 2, 3, 4])
    with contextlib.redirect_stdout(io.StringIO()):
        prob = ns["Multigrid"](n)
    prob.initial_v = torch.as_tensor(u0, dtype=torch.float32)
    return prob
