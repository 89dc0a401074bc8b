"""Generates tests/golden/energy_ref_python_engine.npz with the UNMODIFIED reference Python engine
(python/fdtd/sim_fdtd.py, energy_on=True; h5py served by h5lite, tests/refshim.py): H_tot, E_lost, E_in and u_out
for two synthetic cases (Cartesian lossy, checkerboard FCC lossy) and the CTK church folder.  Run in the build
container (needs /root/reference and numba).

    python tests/golden/make_energy_golden.py
"""
import contextlib
import io
import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))
import refshim  # noqa: E402
from cases import make_files  # noqa: E402
from pffdtd_b200 import shoebox  # noqa: E402

ENERGY_CASES = ("cart_lossy", "cart_lossy_mb11", "cart_hann", "fcc1_lossy", "cart_blobs", "fcc1_blobs")  # walls clear of the absorbing shell, as the reference assumes (sim_fdtd.py:152)
ENERGY_FOLDERS = ("ctk_h030_cpu",)


def run_reference(data_dir):
    from fdtd.sim_fdtd import SimEngine as RefEngine
    eng = RefEngine(Path(data_dir), energy_on=True, nthreads=2)
    with contextlib.redirect_stdout(io.StringIO()):
        eng.load_h5_data(); eng.setup_mask(); eng.allocate_mem(); eng.set_coeffs(); eng.checks()
        eng.run_all(1)  # the energy sums are only meaningful with the default --nsteps 1 (sim_fdtd.py:587: u2 = self.u0 is re-read per run_steps call)
    return dict(H_tot=np.array(eng.H_tot), E_lost=np.array(eng.E_lost), E_in=np.array(eng.E_in), u_out=np.array(eng.u_out))


def main():
    refshim.install()
    out = {}
    for name in ENERGY_CASES:
        d = tempfile.mkdtemp(prefix="energy_")
        shoebox.write_folder(make_files(name), d)
        for k, v in run_reference(d).items():
            out[f"{name}/{k}"] = v
    for folder in ENERGY_FOLDERS:
        d = tempfile.mkdtemp(prefix="energy_")
        for f in (HERE / folder).glob("*.h5"):
            shutil.copy(f, Path(d) / f.name)
        for k, v in run_reference(d).items():
            out[f"{folder}/{k}"] = v
    for k in sorted(out):
        print(k, out[k].shape, float(np.abs(out[k]).max()))
    np.savez_compressed(HERE / "energy_ref_python_engine.npz", **out)


if __name__ == "__main__":
    main()
