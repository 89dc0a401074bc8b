"""Full-size real-room inputs for BASELINE configs[1]/[2] (too large to commit): the CTK church voxelised at
h = 0.041 m (about 513x333x179 nodes, 7-point Cartesian) -- and optionally the Musikverein on the FCC grid -- by the
UNMODIFIED reference tool chain (python/sim_setup.py) under the shims of tests/refshim.py, written as "gpu folders"
(rotated, sorted, FCC folded) under data_large/ (git-ignored; travels to the GPU box with the repo snapshot).
tests/test_large_models.py runs them through the CUDA engine and the unmodified reference CPU engine.

    python tools/make_large_models.py ctk [h] [duration]        (build container only: needs /root/reference)
    python tools/make_large_models.py mv  [h] [duration] [folder name under data_large/]
    python tools/make_large_models.py mv 0.03 0.01 mv_fcc_gpu_big        (bench workload "mv_big": about 3e8 stored nodes)
"""
import os
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
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

from pffdtd_b200 import folder_prep, h5lite  # noqa: E402

REF = Path("/root/reference")
CTK_MATS = {'AcousticPanel': 'ctk_acoustic_panel.h5', 'Altar': 'ctk_altar.h5', 'Carpet': 'ctk_carpet.h5', 'Ceiling': 'ctk_ceiling.h5',
            'Glass': 'ctk_window.h5', 'PlushChair': 'ctk_chair.h5', 'Tile': 'ctk_tile.h5', 'Walls': 'ctk_walls.h5'}
MV_MATS = {'Floor': 'mv_floor.h5', 'Chairs': 'mv_chairs.h5', 'Plasterboard': 'mv_plasterboard.h5', 'Window': 'mv_window.h5', 'Wood': 'mv_wood.h5'}


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "ctk"
    from sim_setup import sim_setup
    from fdtd import rotate_sim_data as R
    os.chdir(REF / "python")
    tmp = Path(tempfile.mkdtemp(prefix="large_"))
    t0 = time.time()
    if which == "ctk":
        h = float(sys.argv[2]) if len(sys.argv) > 2 else 0.041
        dur = float(sys.argv[3]) if len(sys.argv) > 3 else 0.04
        sim_setup(model_json_file='../data/models/CTK_Church/model_export.json', mat_folder='../data/materials', source_num=1,
                  insig_type='impulse', diff_source=True, mat_files_dict=CTK_MATS, duration=dur, Tc=20, rh=50, fcc_flag=False, PPW=1.0,
                  fmax=343.2 / h, save_folder=str(tmp / "cpu"), save_folder_gpu=str(tmp / "gpu"), compress=0, Nprocs=7)
        name = "ctk_cart_gpu"
    else:
        h = float(sys.argv[2]) if len(sys.argv) > 2 else 0.08
        dur = float(sys.argv[3]) if len(sys.argv) > 3 else 0.03
        try:
            sim_setup(model_json_file='../data/models/Musikverein_ConcertHall/model_export.json', mat_folder='../data/materials', source_num=3,
                      insig_type='impulse', diff_source=True, mat_files_dict=MV_MATS, duration=dur, Tc=20, rh=50, fcc_flag=True, PPW=1.0,
                      fmax=343.2 / h, save_folder=str(tmp / "cpu"), compress=0, Nprocs=7)
        except AssertionError as ex:
            print("reference clash check:", repr(ex)[:100])
        R.copy_sim_data(tmp / "cpu", tmp / "gpu")
        R.rotate_sim_data(tmp / "gpu")
        R.fold_fcc_sim_data(tmp / "gpu")
        R.sort_sim_data(tmp / "gpu")
        name = sys.argv[4] if len(sys.argv) > 4 else "mv_fcc_gpu"
    files = folder_prep.load_folder(tmp / "gpu")
    dst = ROOT / "data_large" / name
    dst.mkdir(parents=True, exist_ok=True)
    for stem, ds in files.items():
        h5lite.write_all(dst / f"{stem}.h5", ds, compression=4)
    v = files["vox_out"]
    print(f"{name}: h={h} grid {int(v['Nx'])}x{int(v['Ny'])}x{int(v['Nz'])} Nb={int(v['Nb'])} Nt={int(files['comms_out']['Nt'])} "
          f"in {time.time() - t0:.0f} s -> {dst}", flush=True)


if __name__ == "__main__":
    main()
