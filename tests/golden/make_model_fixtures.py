"""Generates the real-room fixtures under tests/golden/ with the UNMODIFIED reference tool chain
(python/sim_setup.py: RoomGeo -> voxeliser -> SimComms -> ... and python/fdtd/rotate_sim_data.py), run under the
shims of tests/refshim.py (h5py served by h5lite; numpy-2 / Python-3.12 compatibility patches), and the golden
receiver traces for them with the unmodified reference C CPU engine (oracle/_ref):

    ctk_h030      CTK church, 7-point Cartesian, h = 0.30 m (77x53x32 grid), 8 materials x 11 branches,
                  BASELINE.json configs[0]; 'cpu' folder (unsorted) and 'gpu' folder (rotated + sorted)
    mv_h040_fcc   Musikverein, 13-point FCC, h = 0.40 m, 5 materials, 'gpu' folder: rotated, FOLDED (fcc_flag 2), sorted

    python tests/golden/make_model_fixtures.py          (build container only: needs /root/reference)
"""
import os
import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))
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

from oracle import Reference  # noqa: E402
from pffdtd_b200 import folder_prep, h5lite  # noqa: E402

REF = Path("/root/reference")
CTK_MATS = {'AcousticPanel': 'ctk_acoustic_panel.h5', 'Altar': 'ctk_altar.h5', 'Carpet': 'ctk_carpet.h5', 'Ceiling': 'ctk_ceiling.h5',
            'Glass': 'ctk_window.h5', 'PlushChair': 'ctk_chair.h5', 'Tile': 'ctk_tile.h5', 'Walls': 'ctk_walls.h5'}
MV_MATS = {'Floor': 'mv_floor.h5', 'Chairs': 'mv_chairs.h5', 'Plasterboard': 'mv_plasterboard.h5', 'Window': 'mv_window.h5', 'Wood': 'mv_wood.h5'}


def write_compact(files, dst):
    dst.mkdir(parents=True, exist_ok=True)
    for stem, ds in files.items():
        h5lite.write_all(dst / f"{stem}.h5", ds, compression=9)


def golden(files, folder):
    out = {}
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    for prec in (1, 2):
        sys.stdout.flush()
        os.dup2(devnull, 1)
        try:
            u, _ = Reference(prec, files, folder).run()
        finally:
            os.dup2(saved, 1)
        out[f"p{prec}"] = u
    return out


def main():
    from sim_setup import sim_setup
    from fdtd import rotate_sim_data as R
    os.chdir(REF / "python")
    tmp = Path(tempfile.mkdtemp(prefix="fixtures_"))
    traces = {}

    # ---- CTK, Cartesian
    h = 0.30
    sim_setup(model_json_file='../data/models/CTK_Church/model_export.json', mat_folder='../data/materials', source_num=1,
              insig_type='impulse', diff_source=True, mat_files_dict=CTK_MATS, duration=0.11, Tc=20, rh=50, fcc_flag=False, PPW=1.0,
              fmax=343.2 / h, save_folder=str(tmp / "ctk_cpu"), save_folder_gpu=str(tmp / "ctk_gpu"), compress=0, Nprocs=4)
    for kind in ("cpu", "gpu"):
        files = folder_prep.load_folder(tmp / f"ctk_{kind}")
        write_compact(files, HERE / f"ctk_h030_{kind}")
        for k, v in golden(files, HERE / f"ctk_h030_{kind}").items():
            traces[f"ctk_h030_{kind}_{k}"] = v

    # ---- Musikverein, FCC: on this coarse grid the reference's clash check fires after the files are written
    # (SURVEY.md App. D-8); the gpu-folder functions are then called directly, as sim_setup.py:119-125 would
    h = 0.40
    try:
        sim_setup(model_json_file='../data/models/Musikverein_ConcertHall/model_export.json', mat_folder='../data/materials', source_num=3,
                  insig_type='impulse', diff_source=True, mat_files_dict=MV_MATS, duration=0.1, Tc=20, rh=50, fcc_flag=True, PPW=1.0,
                  fmax=343.2 / h, save_folder=str(tmp / "mv_cpu"), compress=0, Nprocs=4)
    except AssertionError as ex:
        print("reference clash check:", repr(ex)[:100])
    R.copy_sim_data(tmp / "mv_cpu", tmp / "mv_gpu")
    R.rotate_sim_data(tmp / "mv_gpu")
    R.fold_fcc_sim_data(tmp / "mv_gpu")
    R.sort_sim_data(tmp / "mv_gpu")
    files = folder_prep.load_folder(tmp / "mv_gpu")
    write_compact(files, HERE / "mv_h040_fcc_gpu")
    for k, v in golden(files, HERE / "mv_h040_fcc_gpu").items():
        traces[f"mv_h040_fcc_gpu_{k}"] = v
    np.savez_compressed(HERE / "traces_models_ref_cpu_engine.npz", **traces)
    for k, v in traces.items():
        print(k, v.shape, float(np.abs(v).max()))
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
