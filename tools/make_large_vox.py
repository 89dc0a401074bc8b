"""A production-size voxelisation as a parity / timing case for the GPU voxeliser (too large to commit): the Musikverein on the FCC grid
at h = 0.03 m (1700 x 662 x 508 points), voxelised by the UNMODIFIED reference (VoxScene.calc_adj on 7 processes) with everything it
read and wrote stored in data_large/vox_mv_h003.npz (git-ignored; travels to the GPU box with the repo snapshot).
tests/test_vox.py::test_cuda_voxeliser_at_production_size compares.

    python tools/make_large_vox.py [h]        (build container only: needs /root/reference)
"""
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))
import make_vox_fixtures as F  # noqa: E402  (installs the shims)


def main():
    h = float(sys.argv[1]) if len(sys.argv) > 1 else 0.03
    t0 = time.time()
    rg, cg, vg, vs = F.build("Musikverein_ConcertHall", h, True)
    inp = F.capture_inputs(rg, cg, vg, vs)
    t1 = time.time()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        import voxelizer.vox_scene as VS
        VS.yes_or_no = lambda q: True
        vs.calc_adj(Nprocs=7)
        os.chdir(cwd)
    t2 = time.time()
    dst = ROOT / "data_large" / f"vox_mv_h{int(round(h * 100)):03d}.npz"
    dst.parent.mkdir(exist_ok=True)
    np.savez_compressed(dst, **{f"in_{k}": v for k, v in inp.items()}, out_bn_ixyz=vs.bn_ixyz, out_adj_bn=vs.adj_bn, out_mat_bn=vs.mat_bn,
                        out_saf_bn=vs.saf_bn, ref_seconds=np.float64(t2 - t1), ref_procs=np.int64(7))
    print(f"grid {cg.Nxyz} voxels {len(vg.nonempty_idx)} tris {rg.tris_pre.size} Nb {vs.bn_ixyz.size}: set-up {t1 - t0:.0f} s, reference calc_adj "
          f"(7 processes) {t2 - t1:.0f} s -> {dst} ({dst.stat().st_size / 1e6:.1f} MB)", flush=True)


def fill_case(h):
    """the inputs of VoxGridBase.fill for the same grid (all voxels' boxes, the triangles' normals, which voxels the reference found
    non-empty) -> data_large/voxfill_mv_h<h>.npz; the expected lists are vox_mv_h<h>.npz's in_vox_tri_off / in_vox_tri"""
    from common.room_geo import RoomGeo
    from voxelizer.cart_grid import CartGrid
    from voxelizer.vox_grid import VoxGrid
    from pffdtd_b200.vox_accel import fill_inputs_from_grid
    tag = f"h{int(round(h * 100)):03d}"
    z = np.load(ROOT / "data_large" / f"vox_mv_{tag}.npz")
    os.chdir("/root/reference/python")
    rg = RoomGeo("../data/models/Musikverein_ConcertHall/model_export.json", az_el=[0., 0.])
    cg = CartGrid(h=h, offset=3.5, bmin=rg.bmin, bmax=rg.bmax, fcc=True)
    vg = VoxGrid(rg, cg)  # (not filled: that is what the reference spent minutes on when vox_mv_<tag>.npz was made)
    inp = fill_inputs_from_grid(vg)
    for k in ("v", "cent", "bmin", "bmax"):
        assert np.array_equal(inp[k], z[f"in_{k}"]), k
    starts = {tuple(int(x) for x in v.ixyz_start): i for i, v in enumerate(vg.voxels)}
    nonempty = np.array([starts[tuple(int(x) for x in s)] for s in z["in_vox_start"]], np.int64)
    assert np.all(np.diff(nonempty) > 0)
    dst = ROOT / "data_large" / f"voxfill_mv_{tag}.npz"
    np.savez_compressed(dst, vbmin=inp["vbmin"], vbmax=inp["vbmax"], nor=inp["nor"], nonempty_idx=nonempty)
    print(f"fill case: Nvox {vg.Nvox} nonempty {nonempty.size} Ntris {vg.Ntris} -> {dst} ({dst.stat().st_size / 1e6:.1f} MB)", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[2] == "fill":
        fill_case(float(sys.argv[1]))
    else:
        main()
        fill_case(float(sys.argv[1]) if len(sys.argv) > 1 else 0.03)
