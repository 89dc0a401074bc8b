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
