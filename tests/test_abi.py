"""The C-ABI library loads without a GPU and exports every symbol include/pffdtd_b200.h declares; the ctypes
mirror of pffdtd_desc has the C layout; without a CUDA device the product fails loudly (no CPU fallback)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

from cases import make_sim_data
from pffdtd_b200 import engine
from pffdtd_b200.sim_data import pffdtd_desc

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "pffdtd_b200.h").read_text()


def declared_functions():
    names = re.findall(r"^\s*(?:const\s+)?(?:int64_t|int|char|void)\s*\*?\s*(pffdtd_\w+)\s*\(", HEADER, flags=re.M)
    assert len(names) >= 15
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    L = engine.lib()
    for fn in declared_functions():
        assert hasattr(L, fn), f"{fn} declared in the header but not exported"
    assert b"sm_100a" in L.pffdtd_version()


@pytest.mark.parametrize("xc", (4, 16, 64))
def test_air_chunk_plan_covers_any_number_of_planes(xc):
    """the work queue's x-chunks tile [0, n) exactly for every job length (round 1 capped the plan at 96 chunks, which left one
    chunk of hundreds of planes on grids with Nx > 1536), main chunks are xc planes long and the tail shrinks to short chunks"""
    import numpy as np
    L = engine.lib()
    buf = (C.c_int64 * 40000)()
    for n in list(range(0, 200)) + [510, 1022, 1536, 1537, 2046, 2852, 4097, 32768, 100000]:
        nch = L.pffdtd_air_chunk_plan(n, xc, buf, len(buf))
        assert nch >= 0
        b = np.array(buf[:nch + 1])
        assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) > 0) if n else nch == 0
        if n >= 16 * xc:
            d = np.diff(b)
            assert d.max() == xc and d[-1] <= max(4, xc // 2) and d[-1] <= 8 or xc == 4
            assert np.all(d[:-24] == xc)


def test_desc_layout_matches_the_header():
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "pffdtd_b200.h"\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(pffdtd_desc), ' \
          'offsetof(pffdtd_desc, Nx), offsetof(pffdtd_desc, ix0), offsetof(pffdtd_desc, mat_quads));return 0;}\n'
    tmp = Path(subprocess.run(["mktemp", "-d"], capture_output=True, text=True).stdout.strip())
    (tmp / "t.c").write_text(src)
    subprocess.run(["/usr/bin/gcc", "-I", str(ROOT / "include"), str(tmp / "t.c"), "-o", str(tmp / "t")], check=True)
    size, o_nx, o_ix0, o_quads = map(int, subprocess.run([str(tmp / "t")], capture_output=True, text=True).stdout.split())
    assert C.sizeof(pffdtd_desc) == size
    assert pffdtd_desc.Nx.offset == o_nx and pffdtd_desc.ix0.offset == o_ix0 and pffdtd_desc.mat_quads.offset == o_quads


def test_energy_desc_layout_matches_the_header():
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "pffdtd_b200.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(pffdtd_energy_desc), ' \
          'offsetof(pffdtd_energy_desc, h), offsetof(pffdtd_energy_desc, mat_DEF));return 0;}\n'
    tmp = Path(subprocess.run(["mktemp", "-d"], capture_output=True, text=True).stdout.strip())
    (tmp / "t.c").write_text(src)
    subprocess.run(["/usr/bin/gcc", "-I", str(ROOT / "include"), str(tmp / "t.c"), "-o", str(tmp / "t")], check=True)
    size, o_h, o_def = map(int, subprocess.run([str(tmp / "t")], capture_output=True, text=True).stdout.split())
    d = engine.pffdtd_energy_desc
    assert C.sizeof(d) == size and d.h.offset == o_h and d.mat_DEF.offset == o_def


def test_kernels_are_sm_100a_tma_code():
    """the shipped library holds sm_100a SASS with TMA tensor loads (UTMALDG) for the air kernel"""
    out = subprocess.run(["cuobjdump", "-sass", str(engine.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "UTMALDG" in out and "k_air_tma_cart" in out


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU behaviour")
def test_no_gpu_means_an_error_not_a_fallback():
    sd = make_sim_data("cart_rigid", 2)
    with pytest.raises(engine.PffdtdError) as ei:
        engine.Engine(sd)
    assert ei.value.code == engine.ECUDA
    with pytest.raises(engine.PffdtdError):
        engine.run_sim(sd)


def test_bad_descriptions_are_rejected_before_touching_the_device():
    sd = make_sim_data("cart_rigid", 2)
    d = sd.desc()
    d.struct_size = 8
    h = C.c_void_p()
    assert engine.lib().pffdtd_create(C.byref(d), 0, C.byref(h)) == engine.EINVAL
    assert b"size mismatch" in engine.lib().pffdtd_last_error()
    d = sd.desc()
    d.precision = 3
    assert engine.lib().pffdtd_create(C.byref(d), 0, C.byref(h)) == engine.EINVAL


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU behaviour")
def test_reference_side_binding_fails_loudly_without_a_gpu(tmp_path):
    """integration/b200_engine.h inside the reference's main sequence (oracle/_ref/libpffdtd_refb200_*.so): without a CUDA device
    run_sim must end the process with the engine's message -- the reference's error convention -- not compute anything on the CPU"""
    import sys
    lib = ROOT / "oracle" / "_ref" / "libpffdtd_refb200_f64.so"
    if not lib.exists():
        pytest.skip("oracle/_ref/libpffdtd_refb200_f64.so not built")
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from cases import make_files\n"
        "from oracle import Reference\n"
        "from pffdtd_b200 import shoebox\n"
        "files = make_files('cart_rigid'); shoebox.write_folder(files, %r)\n"
        "Reference(2, files, %r, gpu='b200').run()\n"
        "print('COMPUTED WITHOUT A GPU')\n") % (str(ROOT), str(ROOT / "tests"), str(tmp_path), str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "b200 engine:" in r.stderr and "COMPUTED WITHOUT A GPU" not in r.stdout
