"""'gpu folder' transforms (pffdtd_b200/folder_prep.py) vs the reference's own functions
(python/fdtd/rotate_sim_data.py: rotate_sim_data, fold_fcc_sim_data, sort_sim_data), run unmodified on real
.h5 folders through the h5py shim of tests/refshim.py; and invariants that hold without the reference."""
import shutil
import tempfile
from pathlib import Path

import numpy as np
import pytest

import refshim
from cases import make_files
from oracle import Oracle
from pffdtd_b200 import folder_prep, shoebox

needs_ref = pytest.mark.skipif(not refshim.available(), reason="/root/reference absent")


def _same(a: dict, b: dict):
    for stem in folder_prep.STEMS:
        assert set(a[stem]) == set(b[stem]), stem
        for k in a[stem]:
            x, y = np.asarray(a[stem][k]), np.asarray(b[stem][k])
            assert x.shape == y.shape and np.array_equal(x.astype(np.float64), y.astype(np.float64)), f"{stem}/{k}"


@needs_ref
@pytest.mark.parametrize("fcc,dims", ((False, (16, 24, 20)), (True, (16, 24, 20)), (False, (30, 16, 18)), (True, (16, 18, 26))))
def test_gpu_folder_equals_reference_pipeline(fcc, dims, capsys):
    refshim.install()
    from fdtd import rotate_sim_data as R
    files = shoebox.make_shoebox(*dims, 20, fcc=fcc, nmat=2, mb=3)
    src = Path(tempfile.mkdtemp(prefix="fp_src_"))
    dst = Path(tempfile.mkdtemp(prefix="fp_dst_"))
    shoebox.write_folder(files, src, compress=3)  # chunked + deflate, as sim_setup writes them
    for f in src.glob("*.h5"):
        shutil.copy(f, dst / f.name)
    R.rotate_sim_data(dst)
    if fcc:
        R.fold_fcc_sim_data(dst)
    R.sort_sim_data(dst)
    capsys.readouterr()
    _same(folder_prep.gpu_folder(folder_prep.load_folder(src)), folder_prep.load_folder(dst))


def test_rotation_sorts_dims_descending_and_keeps_the_physics():
    files = shoebox.make_shoebox(16, 24, 20, 30, nmat=1, mb=2)
    g = folder_prep.gpu_folder(files)
    v = g["vox_out"]
    assert (int(v["Nx"]), int(v["Ny"]), int(v["Nz"])) == (24, 20, 16)
    assert np.all(np.diff(v["bn_ixyz"]) > 0) and np.all(np.diff(g["comms_out"]["out_ixyz"]) >= 0)
    # a rotated + sorted folder describes the same room: identical traces, in the original receiver order
    a = shoebox.sim_data_from_files(files, 2).scale_input()
    b = shoebox.sim_data_from_files(g, 2).scale_input()
    ua = a.reorder_output(Oracle(a).run_all())
    ub = b.reorder_output(Oracle(b).run_all())
    assert np.allclose(ua, ub, rtol=0, atol=1e-12 * np.abs(ua).max())  # summation order differs with the axes


def test_fold_halves_y_and_keeps_the_physics():
    files = shoebox.make_shoebox(16, 20, 16, 30, fcc=True, nmat=1, mb=2)
    f = folder_prep.sort(folder_prep.fold_fcc(files))
    assert int(f["vox_out"]["Ny"]) == 11 and int(f["sim_consts"]["fcc_flag"]) == 2
    a = shoebox.sim_data_from_files(files, 2).scale_input()
    b = shoebox.sim_data_from_files(f, 2).scale_input()
    ua = a.reorder_output(Oracle(a).run_all())
    ub = b.reorder_output(Oracle(b).run_all())
    assert np.allclose(ua, ub, rtol=0, atol=1e-12 * np.abs(ua).max())
    with pytest.raises(ValueError):
        folder_prep.fold_fcc(f)
