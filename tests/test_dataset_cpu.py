"""Product-side HDF5 reader / dataset wrappers (FEANet/h5lite.py, FEANet/dataset.py; reference: Data/dataset.py:1-104).
CPU only.  The reference's .h5 files do not travel to the GPU box: fixtures are rebuilt from the committed golden arrays
with the package's own writer; where /root/reference exists the real files are parsed and compared with the manifest."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ARR = np.load(os.path.join(G, "solve_arrays.npz"))
TP = np.load(os.path.join(G, "testpoisson.npz"))
REF = os.environ.get("MGFEA_REFERENCE", "/root/reference")


def iso_fixture(path):
    from FEANet.h5lite import write_h5

    return write_h5(str(path), {"boundary_index": ARR["iso33_bidx"].astype(np.float64),
                                "boundary_value": ARR["iso33_bval"].astype(np.float64),
                                "rhs": ARR["iso33_rhs64"], "u": ARR["iso33_u64"]})


def test_h5_roundtrip(tmp_path):
    from FEANet.h5lite import H5Error, H5File, write_h5

    rs = np.random.RandomState(0)
    arrs = {"a": rs.standard_normal((3, 33, 33)), "b32": rs.standard_normal((5, 7)).astype(np.float32),
            "i": np.arange(24, dtype=np.int32).reshape(2, 3, 4), "u8": np.arange(9, dtype=np.uint8),
            "c": rs.standard_normal((2, 4, 4, 1))}
    f = H5File(write_h5(str(tmp_path / "t.h5"), arrs))
    assert sorted(f.keys()) == sorted(arrs)
    for k, v in arrs.items():
        assert f[k].dtype == v.dtype and np.array_equal(f[k], v) and f.shape(k) == v.shape
    with pytest.raises(KeyError):
        f["nope"]
    (tmp_path / "bad.h5").write_bytes(b"not hdf5 at all")
    with pytest.raises(H5Error):
        H5File(str(tmp_path / "bad.h5"))


def test_isopoisson_dataset(tmp_path):
    """IsoPoissonDataSet (Data/dataset.py:26-51): items (u, f, bc_value, bc_index), each ToTensor()-shaped (1, 33, 33) fp32"""
    from FEANet.dataset import IsoPoissonDataSet, IsoPoissonPBCDataSet

    ds = IsoPoissonDataSet(iso_fixture(tmp_path / "iso.h5"))
    assert len(ds) == 3
    for k in range(3):
        u, f, bv, bi = ds[k]
        for t, ref in ((u, ARR["iso33_u"]), (f, ARR["iso33_rhs"]), (bv, ARR["iso33_bval"]), (bi, ARR["iso33_bidx"])):
            assert tuple(t.shape) == (1, 33, 33) and t.dtype == torch.float32
            assert np.array_equal(t[0].numpy(), ref[k])
    assert torch.equal(IsoPoissonPBCDataSet(str(tmp_path / "iso.h5"))[1], ds[1][1])
    doubled = IsoPoissonDataSet(str(tmp_path / "iso.h5"), transform=lambda t: 2 * t)
    assert torch.equal(doubled[0][1], 2 * ds[0][1])


def test_testpoisson_dataset(tmp_path):
    """TestPoissonDataSet (Data/dataset.py:71-104): seven fp64 fields, `material` per ELEMENT (32 x 32); the real file keeps
    a trailing singleton axis on five of them, which ToTensor() turns into the channel axis"""
    from FEANet.dataset import TestPoissonDataSet
    from FEANet.h5lite import write_h5

    path = write_h5(str(tmp_path / "tp.h5"), {"dirich_idx": TP["dirich_idx"][..., None], "dirich_value": TP["dirich_value"][..., None],
                                              "neumann_idx": TP["neumann_idx"][..., None], "neumann_value": TP["neumann_value"][..., None],
                                              "material": TP["material"][..., None], "source": TP["source"], "solution": TP["solution"]})
    ds = TestPoissonDataSet(path)
    assert len(ds) == 3
    di, dv, ti, tv, mat, src, sol = ds[2]
    assert tuple(mat.shape) == (1, 32, 32) and mat.dtype == torch.float64 and tuple(sol.shape) == (1, 33, 33)
    assert np.array_equal(mat[0].numpy(), TP["material"][2]) and np.array_equal(sol[0].numpy(), TP["solution"][2])
    assert np.array_equal(di[0].numpy(), TP["dirich_idx"][2]) and np.array_equal(src[0].numpy(), TP["source"][2])


def test_reference_files_match_manifest():
    """the real reference files through the product's reader: names, shapes, dtypes and data hashes recorded by
    tests/golden/make_golden.py (which also cross-checks them against the harness' byte-scanning reader)"""
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present (GPU box)")
    from FEANet.dataset import IsoPoissonDataSet
    from FEANet.h5lite import H5File

    man = json.load(open(os.path.join(G, "h5_manifest.json")))
    for rel, sets in man.items():
        f = H5File(os.path.join(REF, rel))
        assert sorted(f.keys()) == sorted(sets)
        for k, d in sets.items():
            a = f[k]
            assert list(a.shape) == d["shape"] and str(a.dtype) == d["dtype"]
            assert hashlib.sha256(a.tobytes()).hexdigest() == d["sha256"]
    ds = IsoPoissonDataSet(os.path.join(REF, "Data/IsoPoisson/poisson2d_33x33.h5"))
    assert len(ds) == 100 and np.array_equal(ds[2][1][0].numpy(), ARR["iso33_rhs"][2])
