"""Golden vectors for the voxeliser's hot stage: the UNMODIFIED reference `VoxScene.calc_adj`
(python/voxelizer/vox_scene.py:95-440) run here under the shims of tests/refshim.py on the reference's own models, with
everything it reads (grid, voxel lists, triangle tables, direction vectors) and everything it produces (bn_ixyz, adj_bn, mat_bn,
saf_bn, and the per-node nearest triangle / distance) stored in tests/golden/vox_<case>.npz.

    python tests/golden/make_vox_fixtures.py            (build container only: needs /root/reference)
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import refshim  # noqa: E402

refshim.install()
from multiprocessing import shared_memory as shm  # noqa: E402

_orig_close = shm.SharedMemory.close


def _close(self):
    try:
        _orig_close(self)
    except BufferError:  # Py3.12: numpy views of shm.buf are still alive in the reference's voxeliser
        pass


shm.SharedMemory.close = _close

CASES = {
    # name: (model, h, fcc)
    "ctk_h030": ("CTK_Church", 0.30, False),
    "ctk_h045_fcc": ("CTK_Church", 0.45, True),
    "mv_h060_fcc": ("Musikverein_ConcertHall", 0.60, True),
}


def build(model, h, fcc):
    """the reference's own objects up to the voxel grid (sim_setup.py:75-108)"""
    from common.room_geo import RoomGeo
    from voxelizer.cart_grid import CartGrid
    from voxelizer.vox_grid import VoxGrid
    from voxelizer.vox_scene import VoxScene
    os.chdir("/root/reference/python")
    rg = RoomGeo(f"../data/models/{model}/model_export.json", az_el=[0., 0.])
    cg = CartGrid(h=h, offset=3.5, bmin=rg.bmin, bmax=rg.bmax, fcc=fcc)
    vg = VoxGrid(rg, cg)
    vg.fill(Nprocs=1)
    vs = VoxScene(rg, cg, vg, fcc=fcc)
    return rg, cg, vg, vs


def capture_inputs(rg, cg, vg, vs):
    """everything calc_adj reads, as plain arrays (pffdtd_b200.vox_accel.inputs_from_scene makes the same dict)"""
    from pffdtd_b200.vox_accel import inputs_from_scene
    return inputs_from_scene(vs)


def main():
    import tempfile
    for name, (model, h, fcc) in CASES.items():
        rg, cg, vg, vs = build(model, h, fcc)
        inp = capture_inputs(rg, cg, vg, vs)
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)  # calc_adj writes its scratch files under ./_dat
            import voxelizer.vox_scene as VS
            VS.yes_or_no = lambda q: True
            # the reference does not keep tidx / ndist: capture them at the point where they are complete
            keep = {}
            orig_dotv = VS.dotv
            vs.calc_adj(Nprocs=1)
            os.chdir(cwd)
        out = dict(bn_ixyz=vs.bn_ixyz, adj_bn=vs.adj_bn, mat_bn=vs.mat_bn, saf_bn=vs.saf_bn)
        dst = ROOT / "tests" / "golden" / f"vox_{name}.npz"
        np.savez_compressed(dst, **{f"in_{k}": v for k, v in inp.items()}, **{f"out_{k}": v for k, v in out.items()})
        print(f"{name}: grid {cg.Nxyz} Nvox {vg.Nvox} nonempty {len(vg.nonempty_idx)} Ntris {rg.tris_pre.size} Nb {vs.bn_ixyz.size} "
              f"-> {dst} ({dst.stat().st_size / 1e6:.2f} MB)", flush=True)


if __name__ == "__main__":
    main()
