"""The host run loop (pffdtd_b200/sim_fdtd.py) end to end on the CPU: folder in -> SimEngine method sequence -> sim_outs.h5 out,
with the CUDA engine replaced by a stand-in that steps the test oracle.  What is checked is the HOST logic the GPU tests
share -- loading, scaling, sorting for folded FCC, batching, receiver reordering, rescaling, the HDF5 writer, the log lines,
the energy plumbing and the reference's CLI flags -- against the golden traces of the unmodified reference CPU engine."""
import shutil
from pathlib import Path

import numpy as np
import pytest

from cases import make_files
from oracle import Oracle
from pffdtd_b200 import h5lite, shoebox
from pffdtd_b200 import sim_fdtd

GOLD_DIR = Path(__file__).parent / "golden"
GOLD = np.load(GOLD_DIR / "traces_ref_cpu_engine.npz")
GOLD_MODELS = np.load(GOLD_DIR / "traces_models_ref_cpu_engine.npz")


class OracleEngine:
    """the calls SimEngine makes on pffdtd_b200.engine.Engine, served by the CPU oracle"""

    def __init__(self, sd, device=0):
        self.sd, self.o, self.opts, self.air_calls = sd, Oracle(sd), {}, 0

    def set_option(self, k, v):
        self.opts[k] = v

    def reset_stats(self):
        pass

    def stat(self, k):
        return {"air_ms": 1.0}[k]

    def run_steps(self, n0, k):
        self.o.run_steps(n0, k)

    def sync(self):
        pass

    def read_outputs(self, n0=0, n1=None):
        return self.o.u_out[:, n0:self.sd.Nt if n1 is None else n1].copy()

    def energy_enable(self):
        self.energy = True

    def read_energy(self):
        Nt = self.sd.Nt
        return np.arange(Nt, dtype=float), np.zeros(Nt + 1), np.zeros(Nt + 1)

    def close(self):
        pass


@pytest.fixture
def stub(monkeypatch):
    monkeypatch.setattr(sim_fdtd, "Engine", OracleEngine)


@pytest.mark.parametrize("name,precision", (("cart_lossy_mb11", 1), ("cart_ragged", 2), ("fcc2_lossy", 1), ("fcc1_lossy", 2), ("cart_empty", 1)))
def test_folder_to_sim_outs_through_the_host_loop(tmp_path, stub, name, precision):
    shoebox.write_folder(make_files(name), tmp_path, compress=2)
    u = sim_fdtd.run_folder(tmp_path, precision=precision, nsteps=7)
    gold = GOLD[f"{name}_p{precision}"]
    assert np.array_equal(u, gold)
    assert np.array_equal(h5lite.File(tmp_path / "sim_outs.h5")["u_out"][...], gold)


@pytest.mark.parametrize("precision", (1, 2))
def test_plain_fcc_folder_run_as_its_gpu_folder(tmp_path, stub, precision):
    """--gpu_folder: the checkerboard folder (fcc_flag 1) is rotated, folded and sorted in memory; the result is what the reference
    computes from the gpu folder sim_setup would have written (golden fcc2_lossy), in the ORIGINAL receiver order"""
    shoebox.write_folder(make_files("fcc1_lossy"), tmp_path)
    u = sim_fdtd.run_folder(tmp_path, precision=precision, gpu_folder=True)
    assert np.array_equal(u, GOLD[f"fcc2_lossy_p{precision}"])


def test_plain_cartesian_room_run_as_its_gpu_folder(tmp_path, stub):
    for f in (GOLD_DIR / "ctk_h030_cpu").glob("*.h5"):
        shutil.copy(f, tmp_path / f.name)
    u = sim_fdtd.run_folder(tmp_path, precision=1, gpu_folder=True)
    assert np.array_equal(u, GOLD_MODELS["ctk_h030_gpu_p1"])


def test_real_room_unsorted_folder_through_the_cli(tmp_path, stub, capsys):
    for f in (GOLD_DIR / "ctk_h030_cpu").glob("*.h5"):
        shutil.copy(f, tmp_path / f.name)
    sim_fdtd.main(["--data_dir", str(tmp_path), "--precision", "2", "--nsteps", "50", "--nthreads", "4", "--timing", "--abc", "--draw_backend", "mayavi"])
    out = capsys.readouterr().out
    assert np.array_equal(h5lite.File(tmp_path / "sim_outs.h5")["u_out"][...], GOLD_MODELS["ctk_h030_cpu_p2"])
    # the reference's log vocabulary: --ENGINE: prefix, the closing three lines of the C engines, the last-samples block
    for needle in ("--ENGINE: loading data..", "--ENGINE: running..", "--ENGINE: Air update: ", "--ENGINE: Boundary loop: ",
                   "--ENGINE: Combined (total): ", "--ENGINE: GRID OUTPUTS", "--ENGINE: out 47", "--ENGINE: sample 218: "):
        assert needle in out, needle


def test_energy_flag_and_plot_flag(tmp_path, stub, capsys):
    shoebox.write_folder(make_files("cart_lossy"), tmp_path)
    sim_fdtd.main(["--data_dir", str(tmp_path), "--energy"])
    out = capsys.readouterr().out
    assert out.count("normalised energy balance:") == 5
    with pytest.raises(SystemExit):
        sim_fdtd.main(["--data_dir", str(tmp_path), "--plot"])
    assert "gather_slice" in capsys.readouterr().err
