"""BASELINE configs[1] / [2] at full size with the REAL rooms: the CTK church voxelised at h = 0.041 m (about 513x333x179,
7-point Cartesian) and the Musikverein on the folded FCC grid, produced by the unmodified reference tool chain
(tools/make_large_models.py -> data_large/, git-ignored, shipped with the repo snapshot).  The CUDA engine's
sim_outs must equal, bit for bit, what the UNMODIFIED reference CPU engine (oracle/_ref, all host cores) computes from
the same folder.  Skipped where the folders or the reference build are absent."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import Reference
from pffdtd_b200 import folder_prep
from pffdtd_b200.sim_data import SimData

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("folder,precision", (("ctk_cart_gpu", 1), ("mv_fcc_gpu", 1)))
def test_full_size_room_equals_the_reference_cpu_engine(folder, precision, capfd):
    d = ROOT / "data_large" / folder
    if not (d / "vox_out.h5").exists():
        pytest.skip(f"{d} not generated (tools/make_large_models.py)")
    if not (ROOT / "oracle" / "_ref" / "libpffdtd_ref_f32.so").exists():
        pytest.skip("oracle/_ref not built")
    from pffdtd_b200.engine import Engine
    sd = SimData.load(d, precision).scale_input()
    if sd.fcc_flag == 2:
        sd = sd.sorted()
    assert sd.Npts > 2.5e7
    with Engine(sd) as e:
        e.run_steps(0, sd.Nt)
        u = sd.reorder_output(sd.rescale_output(e.read_outputs()))
    ref, _ = Reference(precision, folder_prep.load_folder(d), d, threads=os.cpu_count()).run()
    capfd.readouterr()
    assert np.abs(ref).max() > 0
    assert np.array_equal(u, ref), f"max|d| = {np.abs(u - ref).max():.3e} of peak {np.abs(ref).max():.3e}"
