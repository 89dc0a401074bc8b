"""The CPU oracle (oracle/pffdtd_oracle.c, a restatement of c_cuda/cpu_engine.h) is pinned two ways:
 * against tests/golden/traces_ref_cpu_engine.npz, receiver traces written by the UNMODIFIED reference CPU
   engine (tests/golden/make_golden.py) -- runs everywhere;
 * against the unmodified reference engine itself (oracle/_ref), executed here -- when it is available.
Both bit for bit, fp32 and fp64, Cartesian / FCC checkerboard / FCC folded, rigid and lossy walls."""
import os
import tempfile
from pathlib import Path

import numpy as np
import pytest

from cases import CASES, OBSTACLE_CASES, make_files, make_sim_data
from oracle import Oracle, Reference
from pffdtd_b200 import shoebox
from pffdtd_b200.sim_data import SimData

GOLD = np.load(Path(__file__).parent / "golden" / "traces_ref_cpu_engine.npz")


def _oracle_file_order(name, precision):
    sd = make_sim_data(name, precision)  # scale_input applied
    u = Oracle(sd).run_all()
    return sd.reorder_output(sd.rescale_output(u))


@pytest.mark.parametrize("precision", (1, 2))
@pytest.mark.parametrize("name", sorted(CASES) + sorted(OBSTACLE_CASES))
def test_oracle_matches_golden_reference_traces(name, precision):
    got = _oracle_file_order(name, precision)
    ref = GOLD[f"{name}_p{precision}"]
    assert got.shape == ref.shape and np.abs(ref).max() > 0
    assert np.array_equal(got, ref), f"max|d| = {np.abs(got - ref).max():.3e}"


@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built and /root/reference absent")
@pytest.mark.parametrize("precision", (1, 2))
@pytest.mark.parametrize("name", ("cart_lossy", "cart_tight", "fcc1_lossy", "fcc2_lossy"))
def test_oracle_matches_reference_engine_run_here(name, precision, capfd):
    files = make_files(name)
    d = tempfile.mkdtemp(prefix="ref_")
    shoebox.write_folder(files, d)
    ref, _ = Reference(precision, files, d).run()
    capfd.readouterr()  # the reference prints a progress bar
    assert np.array_equal(_oracle_file_order(name, precision), ref)


def test_fp32_and_fp64_traces_agree_to_rounding():
    """sanity of the precision plumbing: the two precisions are different numbers, close to each other"""
    a, b = _oracle_file_order("cart_lossy", 1), _oracle_file_order("cart_lossy", 2)
    assert not np.array_equal(a, b)
    assert np.abs(a - b).max() <= 1e-4 * np.abs(b).max()


def test_unsorted_and_sorted_lists_give_the_same_file():
    """sort_sim_data only permutes the node lists; the file rows come back in the original receiver order"""
    sd = make_sim_data("cart_lossy", 2)
    rng = np.random.default_rng(7)
    pb, po = rng.permutation(sd.Nb), rng.permutation(sd.Nr)
    lossy_of = {int(i): k for k, i in enumerate(sd.bnl_ixyz)}
    files = make_files("cart_lossy")
    v, m = dict(files["vox_out"]), dict(files["comms_out"])
    for n in ("bn_ixyz", "adj_bn", "mat_bn", "saf_bn"):
        v[n] = v[n][pb]
    m["out_ixyz"] = m["out_ixyz"][po]
    m["out_reorder"] = np.argsort(po)
    shuffled = dict(files, vox_out=v, comms_out=m)
    sd2 = shoebox.sim_data_from_files(shuffled, 2).scale_input()
    assert not sd2.is_sorted()
    u2 = sd2.reorder_output(sd2.rescale_output(Oracle(sd2).run_all()))
    sd3 = sd2.sorted()
    assert sd3.is_sorted()
    u3 = sd3.reorder_output(sd3.rescale_output(Oracle(sd3).run_all()))
    ref = GOLD["cart_lossy_p2"]
    assert np.array_equal(u2, ref) and np.array_equal(u3, ref)
    del lossy_of


def _random_room(seed, fcc):
    """a shoebox with 2..6 random solid blocks (seeded), re-drawn until no source / receiver node touches a boundary node"""
    rng = np.random.default_rng(seed)
    N = (int(rng.integers(20, 34)) & ~1, int(rng.integers(18, 30)) & ~1, int(rng.integers(16, 40)) & ~1)
    for _ in range(50):
        blocks = []
        for _ in range(int(rng.integers(2, 7))):
            lo = [int(rng.integers(5, n - 8)) for n in N]
            sz = [int(rng.integers(1, 5)) for _ in N]
            blocks.append((lo[0], lo[0] + sz[0] - 1, lo[1], lo[1] + sz[1] - 1, lo[2], lo[2] + sz[2] - 1))
        try:
            return shoebox.make_shoebox(*N, 36, fcc=fcc, nmat=int(rng.integers(1, 4)), mb=int(rng.integers(1, 7)), obstacles=blocks,
                                        wall_offset=int(rng.integers(1, 4)))
        except AssertionError:
            continue
    raise RuntimeError("no admissible room drawn")


@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built and /root/reference absent")
@pytest.mark.parametrize("seed", range(4))
def test_oracle_matches_reference_engine_on_random_rooms(seed, capfd):
    """seeded random geometry (blocks of 1..4 nodes per axis inside the room, random wall offset, 1-3 materials, 1-6 branches),
    Cartesian for even seeds, checkerboard FCC for odd ones, both precisions: restatement == unmodified reference engine, bit for bit"""
    files = _random_room(seed, fcc=bool(seed & 1))
    d = tempfile.mkdtemp(prefix="ref_")
    shoebox.write_folder(files, d)
    for precision in (1, 2):
        ref, _ = Reference(precision, files, d).run()
        capfd.readouterr()
        sd = shoebox.sim_data_from_files(files, precision).scale_input()
        got = sd.reorder_output(sd.rescale_output(Oracle(sd).run_all()))
        assert np.abs(ref).max() > 0 and np.array_equal(got, ref), f"seed {seed} p{precision}: {np.abs(got - ref).max():.3e}"
