"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/mgfea.h declares,
host logic (closed-form meshes, masks, tables) is bit-exact against the reference fixtures.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol():
    import mgfea

    mgfea.build()
    hdr = open(os.path.join(ROOT, "include", "mgfea.h")).read()
    declared = sorted(set(re.findall(r"\b(mgfea_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 19
    L = ctypes.CDLL(mgfea.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/mgfea.h but not exported"
    assert sorted(mgfea.EXPORTS) == declared
    assert b"sm_100a" in mgfea.lib().mgfea_version()


def test_struct_layouts_match_header():
    import mgfea

    assert ctypes.sizeof(mgfea.Grid) == 72
    assert ctypes.sizeof(mgfea.Ctl) == 32
    assert ctypes.sizeof(mgfea.LevelBufs) == 24
    # sizes printed by gcc for include/mgfea.h (sizeof(mgfea_cycle_cfg), sizeof(mgfea_xchg), sizeof(mgfea_slab))
    assert ctypes.sizeof(mgfea.CycleCfg) == 104
    assert ctypes.sizeof(mgfea.Xchg) == 624
    assert ctypes.sizeof(mgfea.Slab) == 16 and ctypes.sizeof(mgfea.SlabPush) == 56
    assert mgfea.CycleCfg.zero_guess.offset == 96 and mgfea.Xchg.grid.offset == 568 and mgfea.Xchg.seq.offset == 528
    assert mgfea.Xchg.nwait2.offset == 572 and mgfea.Xchg.seq2.offset == 592


def test_no_cpu_fallback():
    import mgfea

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mgfea.MgfeaError):
        mgfea.require_cuda()
    from FEANet.mesh import MeshSquare
    from FEANet.model import KNet

    with pytest.raises(mgfea.MgfeaError):
        KNet(MeshSquare(2, 9))(torch.zeros(1, 1, 9, 9))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "multigrid-feanet_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace(
                    "oracle/mgfea_oracle.c", ""), fn


@pytest.mark.parametrize("n", [4, 8, 16, 32, 64])
@pytest.mark.parametrize("shape", [0, 1])
def test_closed_form_pattern_keys_bit_exact(n, shape):
    from FEANet.mesh import MeshCenterInterface

    M = np.load(os.path.join(G, "mesh.npz"))
    m = MeshCenterInterface(2, [1, 20], n + 1, shape=shape)
    assert np.array_equal(m.pattern_keys, M[f"keys_n{n}_s{shape}"])
    # reference attribute contract
    assert set(m.global_pattern_center.keys()) == set(range(16))
    tot = sum(np.asarray(m.global_pattern_center[k]) for k in range(16))
    assert (tot == 1).all()
    assert m.phase.shape == (n * n,) and m.pattern.shape == ((n + 1) ** 2, 4)


def test_kernel_tables_bit_exact():
    from FEANet.mesh import MeshCenterInterface, MeshSquare

    M = np.load(os.path.join(G, "mesh.npz"))
    for prop in ([1, 20], [1, 100], [3, 0.5]):
        m = MeshCenterInterface(2, prop, 9)
        got = np.stack([m.kernel_dict[k] for k in range(16)])
        assert np.array_equal(got.view(np.uint32), M[f"ktab_{prop[0]}_{prop[1]}"].view(np.uint32))
    assert np.array_equal(MeshSquare(2, 9).kernel_dict[0].view(np.uint32), M["ktab_iso"].view(np.uint32))


def test_closed_form_vs_float_centroid_restatement_large():
    """beyond n=64 the reference cannot run; check the integer closed form against a vectorised fp32 restatement of the
    reference's centroid test (mesh.py:46-48,62-76) at n=512, where fp32 centroids are still exact dyadics"""
    from FEANet.mesh import MeshCenterInterface

    n = 512
    x = np.linspace(1, -1, n + 1, dtype=np.float32)
    y = np.linspace(-1, 1, n + 1, dtype=np.float32)
    cx = ((x[:-1] + x[1:] + x[1:] + x[:-1]) / np.float32(4)).astype(np.float32)
    cy = ((y[:-1] + y[:-1] + y[1:] + y[1:]) / np.float32(4)).astype(np.float32)
    ph = ((cx[None, :] ** 2 + cy[:, None] ** 2) < 0.25).astype(np.uint8)
    assert np.array_equal(MeshCenterInterface._element_phase(n, 0), ph)
    ph2 = ((np.abs(cx)[None, :] < 0.5) & (np.abs(cy)[:, None] < 0.5)).astype(np.uint8)
    assert np.array_equal(MeshCenterInterface._element_phase(n, 1), ph2)


def test_geometry_masks_bit_exact():
    from FEANet.geo import Geometry

    OPS = np.load(os.path.join(G, "ops.npz"))
    for N in (9, 17, 33):
        g = Geometry(N)
        assert np.array_equal(g.geometry_idx.numpy()[0], OPS[f"bidx_{N}"][0])
        assert not g.boundary_value.any()
    with pytest.raises(TypeError):
        Geometry(9, l_shape=True)


def test_module_parameter_contract():
    """state_dict layout of the reference modules is preserved (so Model/*.pth files load)"""
    from FEANet.drivers import HNet
    from FEANet.mesh import MeshCenterInterface
    from FEANet.model import FNet, KNet
    from FEANet.multigrid import ProlongationNet, RestrictionNet

    k = KNet(MeshCenterInterface(2, [1, 20], 9))
    assert k.net1.weight.shape == (16, 1, 3, 3) and k.net2.weight.shape == (1, 16, 3, 3)
    assert k.global_pattern.shape == (1, 16, 9, 9) and float(k.global_pattern.sum()) == 81.0
    assert FNet(0.25).net.weight.shape == (1, 1, 3, 3)
    P = torch.ones(3, 3)
    assert RestrictionNet(P).net.weight.shape == (1, 16, 3, 3)
    assert ProlongationNet(P).net.weight.shape == (16, 1, 3, 3)
    h = HNet(3)
    assert sorted(h.state_dict().keys()) == [f"convLayers.{i}.weight" for i in range(3)]
    ARR = np.load(os.path.join(G, "ops.npz"))
    h.load_state_dict({f"convLayers.{i}.weight": torch.from_numpy(ARR["hnet_w"][i]).reshape(1, 1, 3, 3) for i in range(3)})


def test_training_entry_points_exist():
    """the reference's training entry points: MultiGrid.forward / qm are differentiable through the last
    cycle (iterate_grad), HJacIterator trains the HNet (HRelaxGrad); without a GPU they fail loudly, never silently"""
    import mgfea
    from FEANet.drivers import HJacIterator
    from FEANet.multigrid import MultiGrid

    P4 = torch.tensor([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=torch.float32) / 4.0
    mg = MultiGrid(8, P4 / 4, P4, torch.tensor([4.0, 1.0]))
    assert mg._training() and callable(mg.iterate_grad) and callable(HJacIterator(n=8).TrainSingleEpoch)
    if not torch.cuda.is_available():
        with pytest.raises(mgfea.MgfeaError):
            mg(torch.zeros(1, 1, 9, 9))


def test_set_option_names_documented_in_header():
    """every kernel-selection option named in include/mgfea.h (mgfea_set_option) is accepted by the library, returns the
    previous value and can be restored; an unknown name is MGFEA_EINVAL.  Pure host code: no GPU needed."""
    import re

    import mgfea

    hdr = open(os.path.join(ROOT, "include", "mgfea.h")).read()
    block = hdr[hdr.index("kernel-selection thresholds"):hdr.index("int mgfea_set_option")]
    names = re.findall(r'"([a-z_0-9]+)"', block)
    assert len(names) >= 12 and "hstream_min_n" in names and "stream_one_variant" in names
    L = mgfea.lib()
    for n in names:
        prev = L.mgfea_set_option(n.encode(), 7)
        assert prev >= 0, n
        assert L.mgfea_set_option(n.encode(), prev) == 7, n
    assert L.mgfea_set_option(b"no_such_option", 1) == -1
    assert L.mgfea_set_option(None, 1) == -1
