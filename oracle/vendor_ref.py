#!/usr/bin/env python
"""Recipe that vendors the UNMODIFIED reference hot path into oracle/_ref/ (git-ignored, NOT gpurun-ignored, so it
travels to the GPU box like a built .so) -- test infrastructure, never product source.

    python oracle/vendor_ref.py            # needs /root/reference (or $MGFEA_REFERENCE), read-only

The reference is a Python/torch program: "building" it means copying the files the path consists of, byte for byte,
from where they lie: FEANet/*.py, the three driver notebooks whose code cells are exec'd
(MM_Model_convergence, MM_Interface_error, M-FEANet-mg_test), Utils/plot.py (imported by the notebooks' first cell) and
the shipped HNet weights.  `bench.py --impl reference` then times the reference's own classes on the box's host cores
(cpu_baseline.kind = "reference"); when oracle/_ref/ is absent it falls back to the call-for-call port
(oracle/feanet_torch.py, kind = "port").  A MANIFEST with sha256 sums records what was copied.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MGFEA_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["FEANet/__init__.py", "FEANet/geo.py", "FEANet/jacobi.py", "FEANet/mesh.py", "FEANet/model.py",
         "FEANet/multigrid.py", "Utils/plot.py", "MM_Model_convergence.ipynb", "MM_Interface_error.ipynb",
         "M-FEANet-mg_test.ipynb", "Model/learn_iterator/iso_poisson/iso_poisson_33x33.pth"]


def vendor(force=False):
    if not os.path.isdir(SRC):
        return None
    man_path = os.path.join(DST, "MANIFEST.json")
    if not force and os.path.exists(man_path):
        return DST
    man = {}
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        os.chmod(d, 0o644)
        man[rel] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    json.dump({"source": SRC, "sha256": man}, open(man_path, "w"), indent=1)
    return DST


if __name__ == "__main__":
    out = vendor(force="--force" in sys.argv)
    print(out or f"reference tree not found at {SRC}: nothing vendored")
