"""Golden vectors for the voxel-grid fill (VoxGridBase.fill, python/voxelizer/vox_grid_base.py:67-176): the UNMODIFIED reference
run here on the models of tests/golden/make_vox_fixtures.py.  Stored in tests/golden/voxfill_<case>.npz: the boxes of ALL voxels,
the triangles' area-scaled normals (the other triangle tables and the expected lists of the non-empty voxels are already in
vox_<case>.npz: in_v / in_cent / in_bmin / in_bmax and in_vox_tri_off / in_vox_tri) and the indices of the non-empty voxels.

    python tests/golden/make_voxfill_fixtures.py        (build container only: needs /root/reference)
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))
import make_vox_fixtures as mv  # noqa: E402  (installs the shims)


def main():
    from common.room_geo import RoomGeo
    from voxelizer.cart_grid import CartGrid
    from voxelizer.vox_grid import VoxGrid
    from pffdtd_b200.vox_accel import fill_inputs_from_grid
    for name, (model, h, fcc) in mv.CASES.items():
        os.chdir("/root/reference/python")
        rg = RoomGeo(f"../data/models/{model}/model_export.json", az_el=[0., 0.])
        cg = CartGrid(h=h, offset=3.5, bmin=rg.bmin, bmax=rg.bmax, fcc=fcc)
        vg = VoxGrid(rg, cg)
        vg.fill(Nprocs=1)
        inp = fill_inputs_from_grid(vg)
        z = np.load(ROOT / "tests" / "golden" / f"vox_{name}.npz")
        for k in ("v", "cent", "bmin", "bmax"):
            assert np.array_equal(inp[k], z[f"in_{k}"]), k
        nonempty = np.array(vg.nonempty_idx, np.int64)
        lists = [np.asarray(vg.voxels[i].tri_idxs, np.int32) for i in nonempty]
        assert np.array_equal(np.concatenate(lists), z["in_vox_tri"])
        dst = ROOT / "tests" / "golden" / f"voxfill_{name}.npz"
        np.savez_compressed(dst, vbmin=inp["vbmin"], vbmax=inp["vbmax"], nor=inp["nor"], nonempty_idx=nonempty)
        print(f"{name}: Nvox {vg.Nvox} nonempty {nonempty.size} Ntris {vg.Ntris} -> {dst} ({dst.stat().st_size / 1e6:.2f} MB)", flush=True)


if __name__ == "__main__":
    main()
