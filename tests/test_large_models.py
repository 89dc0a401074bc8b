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


def probe_the_boundary(files, n_clusters=48, Nt=10):
    """A short run that exercises the room's SURFACE at any grid size: `n_clusters` boundary nodes spread evenly over the list, each
    with a source two nodes away (an air node) and the 3 x 3 x 3 block around it as receivers (boundary nodes included: rigid and
    lossy updates show in the traces within a few steps).  Mutates and returns `files` (a load_folder dict)."""
    vo, cm = files["vox_out"], files["comms_out"]
    Nx, Ny, Nz = int(vo["Nx"]), int(vo["Ny"]), int(vo["Nz"])
    bn = np.asarray(vo["bn_ixyz"], np.int64)
    bset = bn if np.all(np.diff(bn) > 0) else np.sort(bn)
    is_bn = lambda q: bool(bset[min(np.searchsorted(bset, q), bset.size - 1)] == q)
    src, rec = [], set()
    for b in bn[np.linspace(0, bn.size - 1, 4 * n_clusters).astype(np.int64)]:
        iz, iy, ix = int(b % Nz), int(b // Nz % Ny), int(b // (Nz * Ny))
        if not (4 <= ix < Nx - 4 and 4 <= iy < Ny - 4 and 4 <= iz < Nz - 4):
            continue
        for dx, dy, dz in ((2, 0, 0), (-2, 0, 0), (0, 2, 0), (0, -2, 0), (0, 0, 2), (0, 0, -2)):
            q = ((ix + dx) * Ny + iy + dy) * Nz + iz + dz
            if not is_bn(q) and q not in src:
                src.append(q)
                rec.update(((ix + a) * Ny + iy + c) * Nz + iz + e for a in (-1, 0, 1) for c in (-1, 0, 1) for e in (-1, 0, 1))
                break
        if len(src) == n_clusters:
            break
    assert len(src) >= n_clusters // 2
    src = np.array(sorted(src), np.int64)
    sig = np.zeros((src.size, Nt))
    sig[:, 0] = 1.0 + 0.01 * np.arange(src.size)
    sig[:, 1] = -0.5
    out = np.array(sorted(rec), np.int64)
    cm.update(in_ixyz=src, in_sigs=sig, Ns=np.int64(src.size), Nt=np.int64(Nt), out_ixyz=out,
              out_reorder=np.arange(out.size, dtype=np.int64), Nr=np.int64(out.size))
    cm.pop("out_alpha", None)
    return files


@pytest.mark.parametrize("folder", ("mv_fcc_gpu_big", "mv_fcc_gpu_full"))
def test_musikverein_at_the_reference_scripts_size_probed_at_the_surface(folder, capfd):
    """mv_fcc_gpu_full: python/test_script_MV_fcc_gpu.py's own grid (fmax 2500 Hz, PPW 7.7: h = 0.0178 m, 2852 x 552 x 850 stored nodes
    = 1.34 G, 10.7 GB of fp32 state); mv_fcc_gpu_big: h = 0.03 m.  The real run's receivers stay silent for the first hundreds of
    steps, so the comparison uses 48 source / receiver clusters placed on the surface (probe_the_boundary) and 10 steps of the
    UNMODIFIED reference CPU engine (~2 s per step at the full size)."""
    d = ROOT / "data_large" / folder
    if not (d / "vox_out.h5").exists():
        pytest.skip(f"{d} not generated (tools/make_large_models.py)")
    if not (ROOT / "oracle" / "_ref" / "libpffdtd_ref_f32.so").exists():
        pytest.skip("oracle/_ref not built")
    import time
    from pffdtd_b200.engine import Engine
    files = probe_the_boundary(folder_prep.load_folder(d))
    from pffdtd_b200 import shoebox
    sd = shoebox.sim_data_from_files(files, 1).scale_input().sorted()
    t0 = time.perf_counter()
    with Engine(sd) as e:
        e.run_steps(0, sd.Nt)
        u = sd.reorder_output(sd.rescale_output(e.read_outputs()))
        fused = e.stat("fused")
    t1 = time.perf_counter()
    ref, _ = Reference(1, files, d, threads=os.cpu_count()).run()
    t2 = time.perf_counter()
    capfd.readouterr()
    bnr = np.isin(files["comms_out"]["out_ixyz"], files["vox_out"]["bn_ixyz"])
    live = np.abs(ref).max(axis=1) > 0  # (receivers behind the surface stay silent)
    assert np.count_nonzero(live & bnr) > 100 and np.count_nonzero(live & ~bnr) > 50
    assert np.array_equal(u, ref), f"max|d| = {np.abs(u - ref).max():.3e} of peak {np.abs(ref).max():.3e}"
    with capfd.disabled():
        print(f"\n[{folder}] {sd.Nx}x{sd.Ny}x{sd.Nz} = {sd.Npts / 1e9:.3f} G nodes, Nb {sd.Nb}: {ref.shape[0]} receivers ({int(np.count_nonzero(live & bnr))} live ones on boundary "
              f"nodes) x {sd.Nt} steps bit-identical; engine create + run {t1 - t0:.1f} s, reference CPU engine load + run {t2 - t1:.1f} s (fused={fused})")
