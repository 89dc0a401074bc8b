"""The energy balance (reference: python/fdtd/sim_fdtd.py:587-620 with --energy).

Golden vectors: H_tot / E_lost / E_in written by the UNMODIFIED reference Python engine (energy_on=True, default
--nsteps 1) for three synthetic Cartesian cases, one checkerboard-FCC case and the CTK church
(tests/golden/make_energy_golden.py).  CPU part: the numpy restatement (oracle/energy.py) reproduces them.  GPU
part: the device-side sums (pffdtd_energy_enable / pffdtd_read_energy) reproduce them, in fp64 to rounding
(tolerance 1e-11 of each series' peak: different summation order of ~1e5 terms) and the invariant
H_tot + E_lost == E_in holds to 1e-13; the full-size BASELINE configs[3] case (1024^3 fp64 rigid box) checks
that invariant where no CPU engine can reach.
"""
from pathlib import Path

import numpy as np
import pytest

from cases import make_files
from pffdtd_b200 import shoebox
from pffdtd_b200.sim_data import SimData
from pffdtd_b200.sim_fdtd import energy_balance

GOLD_DIR = Path(__file__).parent / "golden"
GOLD = np.load(GOLD_DIR / "energy_ref_python_engine.npz")
CASES = ("cart_lossy", "cart_lossy_mb11", "cart_hann", "fcc1_lossy", "ctk_h030_cpu")
CPU_CASES = CASES + ("cart_blobs", "fcc1_blobs")  # rooms with solid blocks inside: pinned on the CPU this round, on the GPU next
TOL = 1e-11


def _sd(name, precision=2, scale=False):
    """unscaled by default: the reference Python engine does not apply the C binaries' input scaling"""
    if name.startswith("ctk"):
        sd = SimData.load(GOLD_DIR / name, precision)
    else:
        sd = shoebox.sim_data_from_files(make_files(name), precision)
    return sd.scale_input() if scale else sd


def _close(a, b, tol=TOL):
    return np.abs(a - b).max() <= tol * np.abs(b).max()


@pytest.mark.parametrize("name", CPU_CASES)
def test_reference_balance_holds_in_the_golden_vectors(name):
    b = energy_balance(GOLD[f"{name}/H_tot"], GOLD[f"{name}/E_lost"], GOLD[f"{name}/E_in"])
    assert np.abs(b[2:]).max() < 1e-13


@pytest.mark.parametrize("name", CPU_CASES)
def test_restatement_reproduces_the_reference_python_engine(name):
    from oracle.energy import energy_trace
    H, lost, ein, u = energy_trace(_sd(name))
    assert _close(H, GOLD[f"{name}/H_tot"]) and _close(lost, GOLD[f"{name}/E_lost"]) and _close(ein, GOLD[f"{name}/E_in"])
    assert _close(u, GOLD[f"{name}/u_out"], 1e-12)


def test_energy_needs_h_and_c_and_refuses_folded_grids():
    from pffdtd_b200.sim_fdtd import SimEngine
    eng = SimEngine(GOLD_DIR / "mv_h040_fcc_gpu", energy_on=True)
    with pytest.raises(ValueError):
        eng.load_h5_data()


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_energy_reproduces_the_reference_python_engine(name):
    from pffdtd_b200.engine import Engine
    sd = _sd(name)
    with Engine(sd) as e:
        e.energy_enable()
        assert e.stat("energy") == 1 and e.stat("fused") == 0
        e.run_steps(0, sd.Nt)
        H, lost, ein = e.read_energy()
        u = e.read_outputs()
    assert _close(H, GOLD[f"{name}/H_tot"]) and _close(lost, GOLD[f"{name}/E_lost"]) and _close(ein, GOLD[f"{name}/E_in"])
    assert _close(u, GOLD[f"{name}/u_out"], 1e-12)
    assert np.abs(energy_balance(H, lost, ein)[2:]).max() < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("name,precision", (("cart_lossy_mb11", 2), ("cart_lossy_mb11", 1), ("fcc1_lossy", 2), ("ctk_h030_cpu", 1)))
def test_device_energy_equals_the_restatement_and_leaves_the_traces_alone(name, precision):
    """scaled inputs (what the C binaries run), both precisions: same sums as the restatement evaluated on the CPU
    oracle's states, and receiver traces still bit-identical to the oracle's"""
    from oracle import Oracle
    from oracle.energy import energy_trace
    from pffdtd_b200.engine import Engine
    sd = _sd(name, precision, scale=True)
    Ho, lo, io, _ = energy_trace(sd)
    with Engine(sd) as e:
        e.energy_enable()
        for n in range(0, sd.Nt, 7):  # batches: the sums carry over between calls
            e.run_steps(n, min(7, sd.Nt - n))
        H, lost, ein = e.read_energy()
        u = e.read_outputs()
    tol = 1e-11 if precision == 2 else 1e-5
    assert _close(H, Ho, tol) and _close(lost, lo, tol) and _close(ein, io, tol)
    assert np.array_equal(u, Oracle(sd).run_all())
    if precision == 2:
        assert np.abs(energy_balance(H, lost, ein)[2:]).max() < 1e-13


@pytest.mark.gpu
def test_device_energy_two_slabs_add_up():
    """per-slab partial sums (manual halo exchange on one device) add up to the whole grid's"""
    from pffdtd_b200.engine import Engine
    full = _sd("cart_lossy_mb11").sorted()
    with Engine(full) as e:
        e.energy_enable()
        e.run_steps(0, full.Nt)
        want = e.read_energy()
    engs = [Engine(full.slab(r, 2)) for r in range(2)]
    try:
        for e in engs:
            e.set_option("manual_halo", 1)
            e.energy_enable()
        for n in range(full.Nt):
            for e in engs:
                e.run_steps(n, 1)
            g = [e.read_grid(1) for e in engs]
            g[0][-1] = g[1][1]
            g[1][0] = g[0][-2]
            for e, a in zip(engs, g):
                e.write_grid(1, a)
        got = [a + b for a, b in zip(engs[0].read_energy(), engs[1].read_energy())]
    finally:
        for e in engs:
            e.close()
    for a, b in zip(got, want):
        assert _close(a, b, 1e-12)


@pytest.mark.gpu
def test_energy_is_conserved_in_the_full_size_rigid_box():
    """BASELINE configs[3]: rigid shoebox 1024^3, 7-point Cartesian, fp64 (three 8.6 GB grids).  No CPU engine reaches
    this size in test time; the size-independent property is the reference's own invariant: H_tot + E_lost == E_in
    at every step, and -- sealed rigid box, nothing reaches the absorbing shell -- H_tot constant once the source is
    silent."""
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs > 60 GB of device memory")
    from pffdtd_b200.engine import Engine
    Nt = 48
    files = shoebox.make_shoebox(1024, 1024, 1024, Nt, rigid=True, diff=False)
    sd = shoebox.sim_data_from_files(files, 2).scale_input()
    del files
    with Engine(sd) as e:
        e.energy_enable()
        e.run_steps(0, Nt)
        H, lost, ein = e.read_energy()
    assert H[4] > 0 and np.all(lost == 0.0)
    b = energy_balance(H, lost, ein)
    assert np.abs(b[2:]).max() < 1e-12, np.abs(b[2:]).max()
    assert np.abs(H[4:] - H[4]).max() <= 1e-12 * H[4]
