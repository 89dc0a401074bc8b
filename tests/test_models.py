"""Real rooms: the CTK church (7-point Cartesian, BASELINE configs[0] size) and the Musikverein (13-point FCC,
folded gpu folder), voxelised by the UNMODIFIED reference tool chain (tests/golden/make_model_fixtures.py), with
golden traces from the unmodified reference C CPU engine.  CPU part: the oracle reproduces them; the reference
PYTHON engine agrees to rounding (it is a different operation order).  GPU part: the CUDA
path reproduces them bit for bit, fp64 and fp32, from the .h5 folder to sim_outs.h5."""
import shutil
import sys
from pathlib import Path

import numpy as np
import pytest

import refshim
from oracle import Oracle
from pffdtd_b200 import folder_prep, h5lite
from pffdtd_b200.sim_data import SimData

GOLD_DIR = Path(__file__).parent / "golden"
GOLD = np.load(GOLD_DIR / "traces_models_ref_cpu_engine.npz")
FOLDERS = ("ctk_h030_cpu", "ctk_h030_gpu", "mv_h040_fcc_gpu")


@pytest.mark.parametrize("precision", (1, 2))
@pytest.mark.parametrize("folder", FOLDERS)
def test_oracle_reproduces_reference_engine_on_real_rooms(folder, precision):
    sd = SimData.load(GOLD_DIR / folder, precision).scale_input()
    if sd.fcc_flag == 2:
        sd = sd.sorted()
    u = sd.reorder_output(sd.rescale_output(Oracle(sd).run_all()))
    assert np.array_equal(u, GOLD[f"{folder}_p{precision}"])


def test_real_rooms_are_not_shoeboxes():
    """what the synthetic cases cannot offer: eight materials, isolated K=0 nodes, unsorted lists, folded adjacency"""
    ctk = SimData.load(GOLD_DIR / "ctk_h030_cpu", 2)
    assert (ctk.Nx, ctk.Ny, ctk.Nz) == (77, 53, 32) and ctk.Nm == 8 and set(ctk.Mb.tolist()) == {11}
    assert len(set(ctk.mat_bnl.tolist())) >= 6 and not ctk.is_sorted() and ctk.Nr == 48 and ctk.Ns == 8
    mv = SimData.load(GOLD_DIR / "mv_h040_fcc_gpu", 2)
    assert mv.fcc_flag == 2 and mv.is_sorted() and (mv.K_bn == 0).any() and mv.Nm == 5


def test_own_gpu_folder_prep_equals_the_reference_on_the_church():
    mine = folder_prep.gpu_folder(folder_prep.load_folder(GOLD_DIR / "ctk_h030_cpu"))
    ref = folder_prep.load_folder(GOLD_DIR / "ctk_h030_gpu")
    for stem in folder_prep.STEMS:
        for k in ref[stem]:
            assert np.array_equal(np.asarray(mine[stem][k], np.float64), np.asarray(ref[stem][k], np.float64)), f"{stem}/{k}"


@pytest.mark.skipif(not refshim.available(), reason="/root/reference absent")
def test_reference_python_engine_agrees_to_rounding(tmp_path, capsys):
    """BASELINE configs[0]: the reference Python/numba engine on the church, ~200 steps.  Not bit-identical to the C
    engine (Laplacian form, no input scaling) -- the budget of SURVEY.md App. C is 1e-12 of the trace peak."""
    pytest.importorskip("numba")
    refshim.install()
    from fdtd.sim_fdtd import SimEngine as RefEngine
    for f in (GOLD_DIR / "ctk_h030_cpu").glob("*.h5"):
        shutil.copy(f, tmp_path / f.name)
    eng = RefEngine(tmp_path, energy_on=False, nthreads=2)
    eng.load_h5_data(); eng.setup_mask(); eng.allocate_mem(); eng.set_coeffs(); eng.checks()
    eng.run_all(50)
    eng.save_outputs()
    capsys.readouterr()
    py = h5lite.File(tmp_path / "sim_outs.h5")["u_out"][...]
    gold = GOLD["ctk_h030_cpu_p2"]
    assert py.shape == gold.shape
    assert np.abs(py - gold).max() <= 1e-12 * np.abs(gold).max()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", (1, 2))
@pytest.mark.parametrize("folder", FOLDERS)
def test_cuda_engine_on_real_rooms_folder_to_file(tmp_path, folder, precision):
    from pffdtd_b200.sim_fdtd import run_folder
    for f in (GOLD_DIR / folder).glob("*.h5"):
        shutil.copy(f, tmp_path / f.name)
    u = run_folder(tmp_path, precision=precision)
    gold = GOLD[f"{folder}_p{precision}"]
    assert np.array_equal(u, gold), f"max|d| = {np.abs(u - gold).max():.3e} of peak {np.abs(gold).max():.3e}"
    assert np.array_equal(h5lite.File(tmp_path / "sim_outs.h5")["u_out"][...], gold)


@pytest.mark.gpu
@pytest.mark.parametrize("ak,fuse", ((0, 0), (1, 0), (1, 1)))
def test_cuda_engine_kernel_variants_on_the_church(ak, fuse):
    from pffdtd_b200.engine import Engine
    sd = SimData.load(GOLD_DIR / "ctk_h030_gpu", 1).scale_input()
    with Engine(sd) as e:
        e.set_option("air_kernel", ak)
        e.set_option("fuse", fuse)
        e.run_steps(0, sd.Nt)
        u = sd.reorder_output(sd.rescale_output(e.read_outputs()))
    assert np.array_equal(u, GOLD["ctk_h030_gpu_p1"])
