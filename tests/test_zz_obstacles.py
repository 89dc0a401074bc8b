"""Rooms with solid blocks inside (tests/cases.py OBSTACLE_CASES: a pillar, a one-node plate, isolated voxels with K = 0, an
L-shaped block): boundary nodes away from the walls with every adjacency pattern a staircase produces.  CUDA path against the
oracle and the reference's golden traces, bit for bit: traces, and whole grids + branch state from noise, every kernel variant.
(Last file of the suite on purpose: these cases were added after the round's GPU budget was spent; the CPU oracle is pinned on them.)"""
import numpy as np
import pytest

from cases import OBSTACLE_CASES, make_sim_data, noise_grids
from oracle import Oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", (2, 1))
@pytest.mark.parametrize("name", sorted(OBSTACLE_CASES))
def test_obstacle_rooms_bit_exact(name, precision):
    from pffdtd_b200.engine import Engine
    sd = make_sim_data(name, precision)
    ref = Oracle(sd).run_all()
    g1, g0 = noise_grids(sd)
    o = Oracle(sd)
    o.write_grid(1, g1)
    o.write_grid(0, g0)
    o.run_steps(0, 20)
    vo, go = o.read_boundary_state()
    for ak, fuse in (((0, 0), (1, 0), (1, 1)) if sd.fcc_flag == 0 else ((0, 0), (1, 0))):
        with Engine(sd) as e:
            e.set_option("air_kernel", ak)
            e.set_option("fuse", fuse)
            e.run_steps(0, sd.Nt)
            assert np.array_equal(e.read_outputs(), ref), f"{name} p{precision} ak={ak} fuse={fuse}: traces"
        with Engine(sd) as e:
            e.set_option("air_kernel", ak)
            e.set_option("fuse", fuse)
            e.write_grid(1, g1)
            e.write_grid(0, g0)
            e.run_steps(0, 20)
            for which in (1, 0):
                a, b = e.read_grid(which)[1:-1, 1:-1, 1:-1], o.read_grid(which)[1:-1, 1:-1, 1:-1]
                assert np.array_equal(a, b), f"{name} p{precision} ak={ak} fuse={fuse} grid{which}: {np.abs(a - b).max():.3e}"
            v, g = e.read_boundary_state()
            assert np.array_equal(v, vo) and np.array_equal(g, go)


def test_full_size_fp64_box_equals_the_reference_cpu_engine(tmp_path, capfd):
    """BASELINE configs[3] at its full size: rigid shoebox 1024^3, 7-point Cartesian, fp64 -- receivers on a lattice of 100 nodes within 7 nodes
    of the source, 12 steps, against the UNMODIFIED reference CPU engine on the box's cores (bit-exact).
    Needs ~20 GB of device memory and ~20 GB of host memory; skipped on smaller machines."""
    import os
    import torch
    from oracle import Reference
    from pffdtd_b200 import shoebox
    from pffdtd_b200.engine import Engine
    if torch.cuda.get_device_properties(0).total_memory < 40e9:
        pytest.skip("needs > 40 GB of device memory")
    try:
        with open("/proc/meminfo") as f:
            avail_kb = int(next(l for l in f if l.startswith("MemAvailable")).split()[1])
        if avail_kb < 48e6:
            pytest.skip("needs > 48 GB of free host memory")
    except (OSError, StopIteration):
        pass
    if not Reference.available(2):
        pytest.skip("oracle/_ref not built")
    N, Nt = 1024, 12
    files = shoebox.make_shoebox(N, N, N, Nt, rigid=True, diff=False)
    cm = files["comms_out"]
    src = int(cm["in_ixyz"][0])
    ring = sorted({src + dx * N * N + dy * N + dz for dx in (-5, -2, 0, 3, 7) for dy in (-4, 0, 2, 6) for dz in (-7, -3, 0, 2, 5)})
    cm.update(out_ixyz=np.array(ring, np.int64), out_reorder=np.arange(len(ring), dtype=np.int64), Nr=np.int64(len(ring)),
              out_alpha=np.full((1, len(ring)), 1.0 / len(ring)))
    for fn in ("sim_consts.h5", "vox_out.h5", "comms_out.h5", "sim_mats.h5"):
        (tmp_path / fn).touch()  # the reference loader stat()s the files; the datasets travel in memory
    sd = shoebox.sim_data_from_files(files, 2).scale_input()
    with Engine(sd) as e:
        e.run_steps(0, Nt)
        u = sd.reorder_output(sd.rescale_output(e.read_outputs()))
    ref, _ = Reference(2, files, tmp_path, threads=os.cpu_count()).run()
    capfd.readouterr()
    assert np.count_nonzero(np.abs(ref).max(axis=1)) > 20  # the wave front has passed the nearer receivers
    assert np.array_equal(u, ref), f"max|d| = {np.abs(u - ref).max():.3e} of peak {np.abs(ref).max():.3e}"


@pytest.mark.parametrize("name,precision", (("cart_lossy_mb11", 1), ("cart_ragged", 2), ("fcc2_lossy", 1), ("cart_blobs", 2)))
def test_reference_main_sequence_with_this_engine_as_run_sim(tmp_path, name, precision):
    """INTEGRATION.md section 2 made real: the reference's own load_sim_data -> scale_input -> run_sim -> rescale_output ->
    write_outputs (unmodified, oracle/ref_driver.c) with run_sim supplied by integration/b200_engine.h over the C ABI; the sim_outs
    dataset it writes must equal the golden traces of the reference's CPU engine.  In a child process: the binding follows the
    reference's error convention (message + exit), which must not take the test run with it."""
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    if not (root / "oracle" / "_ref" / "libpffdtd_refb200_f32.so").exists():
        pytest.skip("oracle/_ref/libpffdtd_refb200_*.so not built")
    gold = np.load(Path(__file__).parent / "golden" / "traces_ref_cpu_engine.npz")[f"{name}_p{precision}"]
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import numpy as np\n"
        "from cases import make_files\n"
        "from oracle import Reference\n"
        "from pffdtd_b200 import shoebox\n"
        "files = make_files(%r); shoebox.write_folder(files, %r)\n"
        "u, seconds = Reference(%d, files, %r, gpu='b200').run()\n"
        "assert seconds > 0\n"
        "np.save(%r, u)\n") % (str(root), str(root / "tests"), name, str(tmp_path), precision, str(tmp_path), str(tmp_path / "u.npy"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert np.array_equal(np.load(tmp_path / "u.npy"), gold)
