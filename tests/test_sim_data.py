"""Host prep (pffdtd_b200/sim_data.py) vs the unmodified reference loader (load_sim_data + scale_input,
c_cuda/fdtd_data.h:99-718, 879-909), field by field and bit for bit; plus the slab partition rules."""
import tempfile

import numpy as np
import pytest

from cases import CASES, make_files, make_sim_data
from oracle import Reference
from pffdtd_b200 import shoebox
from pffdtd_b200.sim_data import SimData, abc_count, abc_nodes

needs_ref = pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built and /root/reference absent")

ARRAYS = ("bn_ixyz", "bnl_ixyz", "bna_ixyz", "Q_bna", "in_ixyz", "out_ixyz", "out_reorder", "adj_bn", "ssaf_bnl", "mat_bnl",
          "Mb", "mat_beta", "in_sigs")
SCALARS = ("Nx", "Ny", "Nz", "Nt", "l", "l2", "a1", "a2", "sl2", "lo2", "infac")


@needs_ref
@pytest.mark.parametrize("precision", (1, 2))
@pytest.mark.parametrize("name", ("cart_lossy_mb11", "cart_tight", "cart_ragged", "fcc1_lossy", "fcc2_lossy", "fcc2_rigid"))
def test_fields_equal_reference_loader(name, precision, capfd):
    files = make_files(name)
    d = tempfile.mkdtemp(prefix="sd_")
    shoebox.write_folder(files, d)
    ref = Reference(precision, files, d)
    ref.L.refdrv_scale_input()
    sd = SimData.load(d, precision).scale_input()  # through the on-disk .h5 files
    capfd.readouterr()
    for f in SCALARS:
        assert float(getattr(sd, f)) == float(ref.field(f)), f
    for f in ARRAYS:
        a, b = np.asarray(getattr(sd, f)).ravel(), ref.field(f).ravel()
        assert a.shape == b.shape and np.array_equal(a.astype(b.dtype), b), f
    q = ref.field("mat_quads").reshape(-1, 12, 4)
    assert np.array_equal(sd.mat_quads.astype(q.dtype), q)
    assert np.array_equal(sd.K_bn, ref.field("K_bn"))
    assert np.array_equal(sd.bn_mask, ref.field("bn_mask"))


@pytest.mark.parametrize("dims", ((9, 8, 7), (20, 18, 16), (5, 5, 5), (12, 31, 6)))
@pytest.mark.parametrize("fcc_flag", (0, 1))
def test_abc_shell_count_and_structure(dims, fcc_flag):
    Nx, Ny, Nz = dims
    if fcc_flag and (Nx % 2 or Ny % 2 or Nz % 2):
        pytest.skip("FCC grids have even dims")
    bna, Q = abc_nodes(Nx, Ny, Nz, fcc_flag)
    assert bna.size == abc_count(Nx, Ny, Nz, fcc_flag) == Q.size
    ix, iy, iz = bna // (Ny * Nz), (bna // Nz) % Ny, bna % Nz
    q = ((ix == 1) | (ix == Nx - 2)).astype(int) + ((iy == 1) | (iy == Ny - 2)) + ((iz == 1) | (iz == Nz - 2))
    assert np.array_equal(q, Q) and Q.min() >= 1 and np.unique(bna).size == bna.size
    lo, hi = abc_nodes(Nx, Ny, Nz, fcc_flag, ix_range=(0, Nx // 2)), abc_nodes(Nx, Ny, Nz, fcc_flag, ix_range=(Nx // 2, Nx))
    assert np.array_equal(np.concatenate([lo[0], hi[0]]), bna) and np.array_equal(np.concatenate([lo[1], hi[1]]), Q)


def test_scale_input_is_a_power_of_two_over_the_peak():
    for p, want in ((1, 4.0), (2, 4.0)):
        sd = make_sim_data("cart_lossy", p, scale=False)
        peak = np.abs(sd.in_sigs).max()
        sd.scale_input()
        assert np.isclose(np.abs(sd.in_sigs).max(), want) and np.isclose(sd.infac, peak / want)


@pytest.mark.parametrize("nranks", (2, 3, 5))
def test_slab_partition_covers_everything_once(nranks):
    sd = make_sim_data("cart_lossy_mb11", 2).sorted()
    starts, sizes = SimData.slab_planes(sd.Nx, nranks)
    assert sum(sizes) == sd.Nx and starts[0] == 0 and max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    P = sd.Ny * sd.Nz
    seen = {k: [] for k in ("bn_ixyz", "bnl_ixyz", "bna_ixyz", "in_ixyz", "out_ixyz")}
    for r in range(nranks):
        s = sd.slab(r, nranks)
        assert s.Nx == sizes[r] + (r > 0) + (r < nranks - 1)
        assert s.x_lo_edge == (r == 0) and s.x_hi_edge == (r == nranks - 1) and s.ix0 == starts[r] - (r > 0)
        for k in seen:
            a = getattr(s, k)
            if a.size:
                assert a.min() >= P and a.max() < (s.Nx - 1) * P or k == "out_ixyz"
            seen[k].append(a + s.ix0 * P)
        assert s.in_sigs.shape == (s.Ns, sd.Nt) and s.adj_bn.size == s.Nb and s.Q_bna.size == s.Nba
    for k, parts in seen.items():
        assert np.array_equal(np.concatenate(parts), getattr(sd, k)), k


def test_slab_needs_sorted_lists_and_enough_planes():
    sd = make_sim_data("cart_lossy", 2)
    shuffled = make_sim_data("cart_lossy", 2)
    shuffled.bn_ixyz = shuffled.bn_ixyz[::-1].copy()
    with pytest.raises(ValueError):
        shuffled.slab(0, 2)
    with pytest.raises(ValueError):
        sd.sorted().slab(0, sd.Nx)


def test_inconsistent_inputs_are_rejected():
    files = make_files("cart_lossy")
    v = dict(files["vox_out"])
    v["adj_bn"] = np.ones_like(v["adj_bn"])  # a "boundary" node with all links open
    with pytest.raises(ValueError):
        shoebox.sim_data_from_files(dict(files, vox_out=v), 2)
    c = dict(files["sim_consts"])
    c["l"], c["l2"] = np.float64(0.9), np.float64(0.81)  # above the Cartesian CFL limit
    with pytest.raises(ValueError):
        shoebox.sim_data_from_files(dict(files, sim_consts=c), 2)
    m = dict(files["comms_out"], diff=np.int8(0))
    with pytest.raises(ValueError):
        shoebox.sim_data_from_files(dict(files, comms_out=m), 1)  # fp32 needs a differentiated source


@pytest.mark.skipif(not __import__("refshim").available(), reason="/root/reference absent")
def test_fcc_plot_hole_fill_equals_the_reference():
    """gather_slice's checkerboard fill against the reference's nb_fcc_fill_plot_holes (sim_fdtd.py:888-894)"""
    pytest.importorskip("numba")
    import refshim
    refshim.install()
    from fdtd.sim_fdtd import nb_fcc_fill_plot_holes
    from pffdtd_b200.sim_fdtd import fcc_fill_plot_holes
    rng = np.random.default_rng(5)
    for i3 in (4, 7):
        a = rng.standard_normal((13, 18))
        i1, i2 = np.meshgrid(np.arange(13), np.arange(18), indexing="ij")
        a[((i1 + i2 + i3) % 2) == 1] = 0.0   # the unused nodes of a checkerboard slice hold zeros
        want = a.copy()
        nb_fcc_fill_plot_holes(want, i3)
        assert np.array_equal(fcc_fill_plot_holes(a.copy(), i3), want)


def _run_slabs_in_process(full, nranks, planes=None):
    """the oracle on every slab of `full` in one process, halo planes copied by hand after every step"""
    from oracle import Oracle
    slabs = [full.slab(r, nranks, planes=planes) for r in range(nranks)]
    orcs = [Oracle(s) for s in slabs]
    for n in range(full.Nt):
        for o in orcs:
            o.run_steps(n, 1)
        for r in range(nranks - 1):
            lo, hi = orcs[r], orcs[r + 1]
            up, down = lo.read_plane(slabs[r].Nx - 2), hi.read_plane(1)
            hi.write_plane(0, up)
            lo.write_plane(slabs[r].Nx - 1, down)
    return np.concatenate([o.u_out[:, :full.Nt] for o in orcs], axis=0)


@pytest.mark.parametrize("seed,nranks", ((3, 2), (4, 3), (5, 4), (6, 5)))
def test_slab_split_of_random_rooms_changes_no_bit(seed, nranks):
    """SimData.slab (gpu_engine.h:516-662 split_data): random rooms with solid blocks, Cartesian and folded FCC, 2..5 slabs, both
    precisions -- receiver rows concatenated in rank order equal the single-domain run"""
    from oracle import Oracle
    from pffdtd_b200 import folder_prep
    from test_oracle import _random_room
    files = _random_room(seed, fcc=bool(seed & 1))
    if seed & 1:
        files = folder_prep.gpu_folder(files)
    for precision in (1, 2):
        full = shoebox.sim_data_from_files(files, precision).scale_input().sorted()
        want = Oracle(full).run_all()
        assert np.abs(want).max() > 0
        assert np.array_equal(_run_slabs_in_process(full, nranks), want)


@pytest.mark.parametrize("nranks", (2, 3, 4))
def test_cost_weighted_slabs_balance_the_walls_and_change_no_bit(nranks):
    """slab_planes(cost=plane_costs()): ranks that hold the walls perpendicular to x get fewer planes, every slab keeps >= 2 planes,
    the costs are closer to equal than with the reference's equal-plane split, and the traces are the same bits"""
    from oracle import Oracle
    sd = make_sim_data("cart_lossy_mb11", 2).sorted()
    cost = sd.plane_costs()
    assert cost.shape == (sd.Nx,) and cost[1:-1].min() >= sd.Ny * sd.Nz
    starts, sizes = SimData.slab_planes(sd.Nx, nranks, cost=cost)
    assert sum(sizes) == sd.Nx and starts[0] == 0 and min(sizes) >= 2 and starts == [sum(sizes[:r]) for r in range(nranks)]
    e_starts, e_sizes = SimData.slab_planes(sd.Nx, nranks)
    per = lambda st, sz: [cost[a:a + n].sum() for a, n in zip(st, sz)]
    assert max(per(starts, sizes)) <= max(per(e_starts, e_sizes))
    # heavy end planes (x walls + x shell of a large room): the end slabs get fewer planes
    heavy = np.ones(64)
    heavy[[1, 3, 4, 59, 60, 62]] = 12.0
    hs = SimData.slab_planes(64, 4, cost=heavy)[1]
    assert hs[0] < hs[1] and hs[3] < hs[2] and sum(hs) == 64
    want = Oracle(sd).run_all()
    assert np.array_equal(_run_slabs_in_process(sd, nranks, planes=(starts, sizes)), want)
    with pytest.raises(ValueError):
        SimData.slab_planes(sd.Nx, nranks, cost=cost[:-1])
    with pytest.raises(ValueError):
        SimData.slab_planes(5, 3, cost=np.ones(5))


def test_duplicate_source_nodes_are_sorted_and_split():
    """two source entries on one grid node (legal: the reference accumulates them in list order, cpu_engine.h:310-313) still count
    as a sorted list and split into slabs (gpu_engine.h:562-661 only walks the list); the traces equal the single-domain run"""
    from dataclasses import replace
    from oracle import Oracle
    sd = make_sim_data("cart_lossy", 2).sorted()
    dup = replace(sd, in_ixyz=np.concatenate([sd.in_ixyz, sd.in_ixyz[:2]]), in_sigs=np.concatenate([sd.in_sigs, 0.5 * sd.in_sigs[:2]]), _keep=[])
    dup = dup.sorted()
    assert dup.is_sorted() and dup.Ns == sd.Ns + 2 and np.any(np.diff(dup.in_ixyz) == 0)
    want = Oracle(dup).run_all()
    assert not np.array_equal(want, Oracle(sd).run_all())
    assert np.array_equal(_run_slabs_in_process(dup, 2), want)


@pytest.mark.parametrize("name", ("cart_lossy_mb11", "cart_tight", "fcc2_lossy"))
def test_library_slab_plan_is_the_python_one(name):
    """pffdtd_multi_create (one host thread, all GPUs) cuts the grid with the C++ twins of SimData.plane_costs / slab_planes: same
    planes for the reference's equal split and for the cost-weighted one, and slabs past the grid's capacity are refused"""
    import ctypes as C
    from pffdtd_b200 import engine
    L = engine.lib()
    sd = make_sim_data(name, 2).sorted()
    d = sd.desc()
    for n in (1, 2, 3, 4):
        for balance in (0, 1):
            st, sz = (C.c_int64 * n)(), (C.c_int64 * n)()
            assert L.pffdtd_slab_plan(C.byref(d), n, balance, st, sz) == 0, L.pffdtd_last_error()
            want = SimData.slab_planes(sd.Nx, n, cost=sd.plane_costs() if balance else None)
            assert (list(st), list(sz)) == (list(want[0]), list(want[1])), (name, n, balance)
    st, sz = (C.c_int64 * 64)(), (C.c_int64 * 64)()
    assert L.pffdtd_slab_plan(C.byref(d), sd.Nx, 1, st, sz) != 0
