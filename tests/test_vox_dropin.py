"""The voxeliser drop-ins on the reference's OWN objects (build container only: needs /root/reference).

`pffdtd_b200.vox_accel.fill(vox_grid)` and `.calc_adj(vox_scene)` replace `VoxGrid.fill` and `VoxScene.calc_adj` in the reference's
`sim_setup` (python/sim_setup.py:105-113).  Here they run inside that very sequence -- RoomGeo, CartGrid, VoxGrid, VoxScene of the
unmodified reference -- and the reference's own consumers carry on with what they leave behind: `check_adj_full()` (the stability
pre-requisite, numba) and `save()` -> vox_out.h5, which must hold exactly what the all-reference run holds.  No GPU in this
container: the two compute calls are served by the host restatement over the arithmetic shared with the CUDA kernels
(oracle/libvoxhost.so); the glue under test (what is read from and set on the reference's objects) is the product's.
The same drop-ins on the GPU: tests/test_vox.py -m gpu."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
import refshim  # noqa: E402

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference tree not present")


@pytest.fixture()
def host_served(monkeypatch):
    import oracle
    from pffdtd_b200 import vox_accel as va
    oracle.build()
    L = C.CDLL(str(ROOT / "oracle" / "libvoxhost.so"))
    real_fill = va.fill_lists
    monkeypatch.setattr(va, "fill_lists", lambda inp, device=0, host_lib=None: real_fill(inp, host_lib=L))

    def ray_stage(inp, device=0):
        d, keep = va.make_desc(inp)
        return va._run(L, "voxhost", d, int(inp["NN"]), device=None)
    monkeypatch.setattr(va, "ray_stage", ray_stage)
    return va


@pytest.mark.parametrize("name,model,h,fcc", [("ctk_h045_fcc", "CTK_Church", 0.45, True), ("ctk_h030", "CTK_Church", 0.30, False)])
def test_drop_ins_inside_the_reference_setup_sequence(host_served, tmp_path, name, model, h, fcc):
    va = host_served
    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    import make_vox_fixtures  # noqa: F401  (shims: h5py served by h5lite, shared-memory close under Python 3.12)
    from common.room_geo import RoomGeo
    from voxelizer.cart_grid import CartGrid
    from voxelizer.vox_grid import VoxGrid
    from voxelizer.vox_scene import VoxScene
    import voxelizer.vox_scene as VS
    cwd = os.getcwd()
    try:
        os.chdir("/root/reference/python")
        rg = RoomGeo(f"../data/models/{model}/model_export.json", az_el=[0., 0.])
        cg = CartGrid(h=h, offset=3.5, bmin=rg.bmin, bmax=rg.bmax, fcc=fcc)
        ref_vg = VoxGrid(rg, cg)
        os.chdir(tmp_path)  # (scratch files of the reference land here)
        ref_vg.fill(Nprocs=1)
        vg = VoxGrid(rg, cg)
        assert va.fill(vg) is vg
        assert [int(i) for i in vg.nonempty_idx] == [int(i) for i in ref_vg.nonempty_idx] and len(vg.nonempty_idx) > 100
        for a, b in zip(vg.voxels, ref_vg.voxels):
            assert np.array_equal(np.asarray(a.tri_idxs), np.asarray(b.tri_idxs))
            if len(b.tri_idxs):
                assert a.tris_pre.dtype == b.tris_pre.dtype and a.tris_pre.tobytes() == b.tris_pre.tobytes() and np.array_equal(a.tris_mat, b.tris_mat)
        vs = VoxScene(rg, cg, vg, fcc=fcc)
        va.calc_adj(vs)
        z = np.load(ROOT / "tests" / "golden" / f"vox_{name}.npz")  # what the reference's calc_adj produced for this scene
        for k in ("bn_ixyz", "adj_bn", "mat_bn", "saf_bn"):
            got, want = getattr(vs, k), z[f"out_{k}"]
            assert got.dtype == want.dtype and np.array_equal(got, want), k
        # the reference carries on: full adjacency check, then the file the engine reads
        VS.yes_or_no = lambda q: True
        (tmp_path / VS.DAT_FOLDER).mkdir(exist_ok=True)  # (the reference's calc_adj would have made it)
        vs.check_adj_full()
        vs.save(tmp_path / "ours")
        ref_vs = VoxScene(rg, cg, ref_vg, fcc=fcc)
        for k in ("bn_ixyz", "adj_bn", "mat_bn", "saf_bn"):
            setattr(ref_vs, k, z[f"out_{k}"])
        ref_vs.save(tmp_path / "ref")
        assert (tmp_path / "ours" / "vox_out.h5").read_bytes() == (tmp_path / "ref" / "vox_out.h5").read_bytes()
    finally:
        os.chdir(cwd)
