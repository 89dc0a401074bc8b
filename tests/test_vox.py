"""The voxeliser's hot stage (VoxScene.calc_adj, python/voxelizer/vox_scene.py:95-440).

Golden vectors: tests/golden/vox_*.npz, written by the UNMODIFIED reference voxeliser on its own models (CTK church Cartesian and
FCC, Musikverein FCC; tests/golden/make_vox_fixtures.py) together with everything it read.  CPU part: the host restatement over the
arithmetic shared with the CUDA kernel (oracle/libvoxhost.so over pffdtd_b200/csrc/vox_core.h) reproduces bn_ixyz, adj_bn, mat_bn,
saf_bn bit for bit.  GPU part: pffdtd_vox_run does, including the nearest triangle and its distance for every boundary node."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from pffdtd_b200 import vox_accel as va

ROOT = Path(__file__).resolve().parent.parent
CASES = ("ctk_h030", "ctk_h045_fcc", "mv_h060_fcc")


def _load(name):
    z = np.load(ROOT / "tests" / "golden" / f"vox_{name}.npz")
    return {k[3:]: z[k] for k in z.files if k.startswith("in_")}, {k[4:]: z[k] for k in z.files if k.startswith("out_")}


def _host(inp):
    import oracle
    oracle.build()
    L = C.CDLL(str(ROOT / "oracle" / "libvoxhost.so"))
    d, keep = va.make_desc(inp)
    return va._run(L, "voxhost", d, int(inp["NN"]), device=None)


def _load_fill(name):
    """inputs of VoxGridBase.fill (all voxels' boxes + triangle tables) and the lists the reference made"""
    z, f = np.load(ROOT / "tests" / "golden" / f"vox_{name}.npz"), np.load(ROOT / "tests" / "golden" / f"voxfill_{name}.npz")
    inp = dict(vbmin=f["vbmin"], vbmax=f["vbmax"], nor=f["nor"], v=z["in_v"], cent=z["in_cent"], bmin=z["in_bmin"], bmax=z["in_bmax"])
    return inp, f["nonempty_idx"], z["in_vox_tri_off"], z["in_vox_tri"]


def _check_fill(inp, nonempty, ref_off, ref_tri, off, tri):
    assert off.size == inp["vbmin"].shape[0] + 1 and off[0] == 0 and off[-1] == tri.size
    assert np.array_equal(np.flatnonzero(np.diff(off) > 0), nonempty)  # vox_grid.nonempty_idx
    assert np.array_equal(np.diff(off)[nonempty], np.diff(ref_off)) and np.array_equal(tri, ref_tri)  # every voxel's list, in order


@pytest.mark.parametrize("name", CASES)
def test_host_fill_equals_the_reference_voxel_grid(name):
    inp, nonempty, ref_off, ref_tri = _load_fill(name)
    assert 0 < nonempty.size < inp["vbmin"].shape[0] and ref_tri.size > 5000
    import oracle
    oracle.build()
    off, tri = va.fill_lists(inp, host_lib=C.CDLL(str(ROOT / "oracle" / "libvoxhost.so")))
    _check_fill(inp, nonempty, ref_off, ref_tri, off, tri)
    # the overlap test is more than the bounding boxes: they alone admit more pairs (vox_grid_base.py:110 'candidates')
    k = int(nonempty[nonempty.size // 2])
    cand = np.all((inp["vbmax"][k] >= inp["bmin"]) & (inp["vbmin"][k] <= inp["bmax"]), axis=-1)
    cands = sum(int(np.all((inp["vbmax"][j] >= inp["bmin"]) & (inp["vbmin"][j] <= inp["bmax"]), axis=-1).sum()) for j in nonempty[:200])
    assert cand.sum() >= off[k + 1] - off[k] and cands > int(np.diff(off)[nonempty[:200]].sum())


def _check(inp, out, bn, adj, tidx):
    assert np.array_equal(bn, out["bn_ixyz"]) and adj.dtype == bool and np.array_equal(adj, out["adj_bn"])
    mat, saf = va.finish(inp, bn, adj, tidx)
    assert np.array_equal(mat, out["mat_bn"]) and np.array_equal(saf, out["saf_bn"])


@pytest.mark.parametrize("name", CASES)
def test_host_restatement_equals_the_reference_voxeliser(name):
    inp, out = _load(name)
    assert out["bn_ixyz"].size > 5000 and int(inp["NN"]) == (12 if "fcc" in name else 6)
    bn, adj, tidx, ndist = _host(inp)
    _check(inp, out, bn, adj, tidx)
    assert np.all(tidx >= 0) and np.all(np.isfinite(ndist)) and np.all(ndist >= 0) and np.all(ndist <= float(inp["hf"]) * (1 + 1e-6))
    # nodes lying on the surface have every link cut and are rigid
    on = ~adj.any(axis=1)
    assert on.any() and np.all(out["mat_bn"][on] == -1)


def test_the_voxel_wide_rules_matter():
    """the two early-outs that couple the points of a voxel are part of the contract: a restatement that treats every point on its
    own (no vote on "some point of the voxel has a hit within hf") must NOT be assumed equal -- keep the fixture honest: dropping a
    voxel's triangle list changes the result"""
    inp, out = _load("ctk_h030")
    inp2 = dict(inp)
    off = inp["vox_tri_off"].copy()
    cut = int(off[5] - off[4])
    assert cut > 0
    inp2["vox_tri"] = np.concatenate([inp["vox_tri"][:off[4]], inp["vox_tri"][off[5]:]])
    off[5:] -= cut
    inp2["vox_tri_off"] = off
    bn, adj, _, _ = _host(inp2)
    assert bn.size < out["bn_ixyz"].size


def test_fill_desc_layout_matches_the_header():
    import subprocess
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "pffdtd_b200.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(pffdtd_voxfill_desc), ' \
          'offsetof(pffdtd_voxfill_desc, Ntris), offsetof(pffdtd_voxfill_desc, bmax));return 0;}\n'
    tmp = Path(subprocess.run(["mktemp", "-d"], capture_output=True, text=True).stdout.strip())
    (tmp / "t.c").write_text(src)
    subprocess.run(["/usr/bin/gcc", "-I", str(ROOT / "include"), str(tmp / "t.c"), "-o", str(tmp / "t")], check=True)
    size, o_nt, o_bmax = map(int, subprocess.run([str(tmp / "t")], capture_output=True, text=True).stdout.split())
    d = va.pffdtd_voxfill_desc
    assert C.sizeof(d) == size and d.Ntris.offset == o_nt and d.bmax.offset == o_bmax


def test_desc_layout_matches_the_header():
    import subprocess
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "pffdtd_b200.h"\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(pffdtd_vox_desc), ' \
          'offsetof(pffdtd_vox_desc, hf), offsetof(pffdtd_vox_desc, Nvox), offsetof(pffdtd_vox_desc, eca));return 0;}\n'
    tmp = Path(subprocess.run(["mktemp", "-d"], capture_output=True, text=True).stdout.strip())
    (tmp / "t.c").write_text(src)
    subprocess.run(["/usr/bin/gcc", "-I", str(ROOT / "include"), str(tmp / "t.c"), "-o", str(tmp / "t")], check=True)
    size, o_hf, o_nvox, o_eca = map(int, subprocess.run([str(tmp / "t")], capture_output=True, text=True).stdout.split())
    d = va.pffdtd_vox_desc
    assert C.sizeof(d) == size and d.hf.offset == o_hf and d.Nvox.offset == o_nvox and d.eca.offset == o_eca


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_voxeliser_equals_the_reference_voxeliser(name):
    inp, out = _load(name)
    bn, adj, tidx, ndist = va.ray_stage(inp, device=0)
    _check(inp, out, bn, adj, tidx)
    hb, ha, ht, hn = _host(inp)
    assert np.array_equal(tidx, ht) and np.array_equal(ndist, hn)


@pytest.mark.gpu
def test_cuda_voxeliser_rejects_bad_descriptions():
    from pffdtd_b200.engine import PffdtdError
    inp, _ = _load("ctk_h045_fcc")
    bad = dict(inp)
    bad["vox_start"] = inp["vox_start"].copy()
    bad["vox_start"][0, 0] = int(inp["Nxyz"][0])  # a voxel outside the grid
    with pytest.raises(PffdtdError):
        va.ray_stage(bad, device=0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_fill_equals_the_reference_voxel_grid(name):
    inp, nonempty, ref_off, ref_tri = _load_fill(name)
    off, tri = va.fill_lists(inp, device=0)
    _check_fill(inp, nonempty, ref_off, ref_tri, off, tri)


@pytest.mark.gpu
def test_cuda_fill_drop_in_sets_what_the_reference_fill_sets():
    """`fill(vox_grid)` on an object with the reference VoxGrid's attributes (vox_grid_base.py:43-66, vox_grid.py:31-39)"""
    inp, nonempty, ref_off, ref_tri = _load_fill("ctk_h045_fcc")
    z = np.load(ROOT / "tests" / "golden" / "vox_ctk_h045_fcc.npz")

    class Vox:
        def __init__(self, lo, hi):
            self.bmin, self.bmax, self.tri_idxs, self.tris_pre, self.tris_mat = lo, hi, [], None, None

    class Grid:
        pass
    vg = Grid()
    nt = inp["nor"].shape[0]
    vg.tris_pre = np.zeros(nt, dtype=[(k, np.float64, (3, 3) if k == "v" else (3,)) for k in ("v", "nor", "cent", "bmin", "bmax")])
    for k in ("v", "nor", "cent", "bmin", "bmax"):
        vg.tris_pre[k] = inp[k]
    vg.mats, vg.Ntris, vg.Nvox = z["in_mat_ind"], nt, inp["vbmin"].shape[0]
    vg.voxels = [Vox(lo, hi) for lo, hi in zip(inp["vbmin"], inp["vbmax"])]
    va.fill(vg, device=0)
    assert vg.nonempty_idx == [int(i) for i in nonempty]
    for j, i in enumerate(nonempty):
        want = ref_tri[ref_off[j]:ref_off[j + 1]]
        vox = vg.voxels[int(i)]
        assert np.array_equal(vox.tri_idxs, want) and np.array_equal(vox.tris_mat, vg.mats[want]) and np.array_equal(vox.tris_pre["cent"], inp["cent"][want])
    assert all(len(vg.voxels[i].tri_idxs) == 0 for i in set(range(vg.Nvox)) - set(vg.nonempty_idx))


LARGE = ROOT / "data_large" / "vox_mv_h003.npz"


@pytest.mark.gpu
@pytest.mark.skipif(not LARGE.exists(), reason="data_large/vox_mv_h003.npz not built (tools/make_large_vox.py, build container only)")
def test_cuda_voxeliser_at_production_size(capsys):
    """Musikverein, FCC grid, h = 0.03 m (1700 x 660 x 508 points, 109 098 voxels of which 21 331 meet the surface, 6.6 M boundary
    nodes): both stages against what the unmodified reference produced (tools/make_large_vox.py), with their times"""
    import time
    z, f = np.load(LARGE), np.load(ROOT / "data_large" / "voxfill_mv_h003.npz")
    inp = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    out = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    va.fill_lists(dict(vbmin=f["vbmin"][:64], vbmax=f["vbmax"][:64], nor=f["nor"], v=inp["v"], cent=inp["cent"], bmin=inp["bmin"], bmax=inp["bmax"]), device=0)
    t0 = time.perf_counter()
    off, tri = va.fill_lists(dict(vbmin=f["vbmin"], vbmax=f["vbmax"], nor=f["nor"], v=inp["v"], cent=inp["cent"], bmin=inp["bmin"], bmax=inp["bmax"]), device=0)
    t1 = time.perf_counter()
    ne = f["nonempty_idx"]
    assert np.array_equal(np.flatnonzero(np.diff(off) > 0), ne) and np.array_equal(tri, inp["vox_tri"])
    assert np.array_equal(np.diff(off)[ne], np.diff(inp["vox_tri_off"]))
    t2 = time.perf_counter()
    bn, adj, tidx, ndist = va.ray_stage(inp, device=0)
    t3 = time.perf_counter()
    _check(inp, out, bn, adj, tidx)
    with capsys.disabled():
        print(f"\n[vox production size] fill {t1 - t0:.3f} s (reference: 330 s single process); calc_adj ray stage {t3 - t2:.3f} s for "
              f"{bn.size} boundary nodes (reference: {float(z['ref_seconds']):.0f} s on {int(z['ref_procs'])} processes)")
