"""The HDF5 subset of the drop-in boundary: what sim_setup writes must be readable, what we
write must be readable by h5py-style code.  The image has neither h5py nor libhdf5, so the genuine fixtures are
the reference's shipped material files; everything else is round-tripped through our own writer."""
from pathlib import Path

import numpy as np
import pytest

from pffdtd_b200 import h5lite

MATS = Path("/root/reference/data/materials")


def _datasets():
    rng = np.random.default_rng(3)
    return {
        "f64_scalar": np.float64(0.577), "i64_scalar": np.int64(-123456789012), "i8_scalar": np.int8(-2),
        "f64_vec": rng.standard_normal(1000), "i64_vec": rng.integers(-2**40, 2**40, 777), "i8_vec": rng.integers(-5, 5, 300).astype(np.int8),
        "f64_mat": rng.standard_normal((11, 3)), "bool_mat": rng.random((257, 12)) > 0.5, "empty": np.zeros(0, np.int64),
        "big": rng.standard_normal((300, 700)),
    }


@pytest.mark.parametrize("compression", (None, 0, 3, 9))
def test_round_trip_every_dtype_and_layout(tmp_path, compression):
    d = _datasets()
    p = tmp_path / "t.h5"
    h5lite.write_all(p, d, compression=compression)
    back = h5lite.read_all(p)
    assert set(back) == set(d)
    for k, v in d.items():
        a = np.asarray(v)
        b = back[k]
        assert b.shape == a.shape, k
        if a.dtype == np.bool_:
            assert b.dtype in (np.bool_, np.int8) and np.array_equal(b.astype(bool), a), k
        else:
            assert b.dtype == a.dtype and np.array_equal(b, a), k


def test_many_datasets_span_several_symbol_table_nodes(tmp_path):
    d = {f"mat_{i:02d}_DEF": np.full((11, 3), float(i)) for i in range(40)}
    d.update(Nmat=np.int8(40), Mb=np.full(40, 11, np.int8))
    h5lite.write_all(tmp_path / "m.h5", d)
    back = h5lite.read_all(tmp_path / "m.h5")
    assert len(back) == 42 and all(np.array_equal(back[k], v) for k, v in d.items())


def test_h5py_style_access_pattern_of_the_reference(tmp_path):
    """the calls rotate_sim_data.py / process_outputs.py make: r+, read, delete, re-create, scalar assignment"""
    p = tmp_path / "v.h5"
    f = h5lite.File(p, "w")
    f.create_dataset("Nx", data=np.int64(10))
    f.create_dataset("bn_ixyz", data=np.arange(50, dtype=np.int64), compression="gzip", compression_opts=3)
    f.create_dataset("adj_bn", data=np.ones((50, 6), bool), compression="gzip", compression_opts=3)
    f.close()
    f = h5lite.File(p, "r+")
    assert f["Nx"][()] == 10 and f["bn_ixyz"][...].sum() == 1225 and "adj_bn" in f
    f["Nx"][()] = 12
    f["bn_ixyz"][...] = np.arange(50, dtype=np.int64)[::-1]
    del f["adj_bn"]
    f.create_dataset("adj_bn", data=np.zeros((50, 12), bool), compression="gzip", compression_opts=3)
    f.close()
    g = h5lite.File(p, "r")
    assert g["Nx"][()] == 12 and g["bn_ixyz"][...][0] == 49 and g["adj_bn"][...].shape == (50, 12)
    assert sorted(g.keys()) == ["Nx", "adj_bn", "bn_ixyz"]
    with pytest.raises(KeyError):
        g["nope"]
    g.close()


def test_bad_files_raise(tmp_path):
    (tmp_path / "x.h5").write_bytes(b"not an hdf5 file at all" * 10)
    with pytest.raises(h5lite.H5Error):
        h5lite.File(tmp_path / "x.h5", "r")
    with pytest.raises(FileNotFoundError):
        h5lite.File(tmp_path / "missing.h5", "r")


@pytest.mark.skipif(not MATS.exists(), reason="/root/reference absent")
def test_reads_the_genuine_hdf5_files_shipped_with_the_reference(tmp_path):
    """written by real h5py/libhdf5: 14 material files, one 11x3 float64 dataset `DEF` each"""
    files = sorted(MATS.glob("*.h5"))
    assert len(files) >= 10
    for p in files:
        d = h5lite.read_all(p)
        assert list(d) == ["DEF"] and d["DEF"].shape == (11, 3) and d["DEF"].dtype == np.float64
        assert np.isfinite(d["DEF"]).all() and (d["DEF"] >= 0).all() and d["DEF"].max() > 0
        # our writer reproduces a file our reader (and theirs: same superblock/heap/B-tree/object-header layout) parses back
        h5lite.write_all(tmp_path / p.name, d)
        assert np.array_equal(h5lite.read_all(tmp_path / p.name)["DEF"], d["DEF"])


def test_reads_a_genuine_libhdf5_file_with_a_user_block():
    """scipy ships one real HDF5 file (a MATLAB v7.3 .mat written by libhdf5, with a 512-byte user block in front of the
    superblock and a (9,1) float64 dataset): the only genuine file besides the reference's materials that travels with the image"""
    import scipy.io
    p = Path(scipy.io.__file__).parent / "matlab" / "tests" / "data" / "testhdf5_7.4_GLNX86.mat"
    if not p.exists():
        pytest.skip("scipy test data not installed")
    d = h5lite.read_all(p)
    assert list(d) == ["testdouble"]
    assert d["testdouble"].shape == (9, 1) and np.allclose(d["testdouble"].ravel(), np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)
