"""the stage AFTER the simulation step.  The reference's post-processing (python/fdtd/process_outputs.py,
UNMODIFIED) reads comms_out.h5 / sim_consts.h5 / sim_outs.h5, re-opens sim_outs.h5 in 'r+' to (re)write `r_out`, and saves
sim_outs_processed.h5.  Here it runs on a sim_outs.h5 written by this repo's host loop (h5py served by h5lite, matplotlib and
resampy stubbed: they are absent from the image and only used for plots / resampling)."""
import sys
import types

import numpy as np
import pytest

import refshim
from cases import make_files
from pffdtd_b200 import h5lite, shoebox, sim_fdtd

pytestmark = pytest.mark.skipif(not refshim.available(), reason="/root/reference absent")


def test_reference_post_processing_accepts_our_sim_outs(tmp_path, monkeypatch, capsys):
    import test_host_loop
    monkeypatch.setattr(sim_fdtd, "Engine", test_host_loop.OracleEngine)  # CPU stand-in for the CUDA engine (test oracle)
    refshim.install()
    for name in ("matplotlib", "matplotlib.pyplot", "resampy"):
        monkeypatch.setitem(sys.modules, name, sys.modules.get(name) or types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(sys.modules["resampy"], "resample"):
        sys.modules["resampy"].resample = None
    from fdtd.process_outputs import ProcessOutputs
    shoebox.write_folder(make_files("cart_lossy_mb11"), tmp_path)
    u = sim_fdtd.run_folder(tmp_path, precision=2)
    for _ in range(2):  # the second pass deletes and re-creates r_out in the file we wrote ('r+')
        po = ProcessOutputs(tmp_path)
        po.initial_process(fcut=10.0, N_order=4)
        po.apply_lowpass(fcut=2000.0, N_order=8, symmetric=True)
        po.save_h5()
    capsys.readouterr()
    f = h5lite.File(tmp_path / "sim_outs.h5")
    assert sorted(f.keys()) == ["r_out", "u_out"] and np.array_equal(f["u_out"][...], u)
    r = f["r_out"][...]
    assert r.shape == (u.shape[0] // 8, u.shape[1]) and np.isfinite(r).all() and np.abs(r).max() > 0
    g = h5lite.File(tmp_path / "sim_outs_processed.h5")
    assert g["r_out_f"][...].shape == r.shape and float(g["Fs_f"][()]) > 0
