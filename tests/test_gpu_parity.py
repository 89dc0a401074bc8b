"""CUDA path vs the CPU oracle, through the C ABI (include/pffdtd_b200.h).

Bar: BIT-EXACT receiver traces and grids in fp64 AND fp32 -- the kernels reproduce the reference CPU
engine's arithmetic (cpu_engine.h:129-325) operation for operation, so no tolerance is needed.
"""
import numpy as np
import pytest

from cases import CASES, make_sim_data, noise_grids
from oracle import Oracle
from pffdtd_b200.engine import Engine, PffdtdError, run_sim

pytestmark = pytest.mark.gpu


def _kernels(sd):
    """(air_kernel, fuse, svc): generic kernel; tiled kernel with separate ABC / mirror kernels; tiled kernel with the
    absorbing shell and the halo mirrors fused in, boundary work by the list kernels (round 1's step); the same with the
    sparse rigid nodes and the shell's z faces done by the air kernel's service warp (the default where the grid allows it)"""
    if sd.fcc_flag == 0:
        return (0, 0, 0), (1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1)
    # FCC: generic and tiled 13-point kernels (unfused step), without / with the service warp; the fused step (mirror-on-write incl. halo
    # edges and the seam row, shell by k_abc_faces + the service warp) where the grid allows it
    return (0, 0, 0), (1, 0, 0), (1, 0, 1), (1, 1, 1)


def _engine(sd, ak, fuse, cfg=None, svc=1):
    e = Engine(sd)
    e.set_option("air_kernel", ak)
    e.set_option("fuse", fuse)
    e.set_option("svc", svc)  # (also resets the tile configuration to the grid's default)
    if cfg is not None:
        e.set_option("air_cfg", cfg)
    return e


def _fused_expected(sd, lz):
    """the fused 13-point step is refused when the shell node z = Nz-2 opens a z tile (tile = lz vectors of 16 bytes); the 7-point kernel
    handles that alignment itself (round 1 fell back to the unfused step for it: cart_nz_e / cart_nz_f)"""
    return sd.fcc_flag == 0 or (sd.Nz - 2) % (lz * (4 if sd.precision == 1 else 2)) != 0


@pytest.mark.parametrize("precision", (2, 1))
@pytest.mark.parametrize("name", sorted(CASES))
def test_traces_bit_exact(name, precision):
    sd = make_sim_data(name, precision)
    ref = Oracle(sd).run_all()
    assert np.abs(ref).max() > 0
    for ak, fuse, svc in _kernels(sd):
        with _engine(sd, ak, fuse, svc=svc) as e:
            e.run_steps(0, sd.Nt)
            got = e.read_outputs()
        assert np.array_equal(got, ref), f"{name} p{precision} air_kernel={ak} fuse={fuse} svc={svc}: max|d|={np.abs(got - ref).max():.3e}"


@pytest.mark.parametrize("precision", (2, 1))
@pytest.mark.parametrize("name", ("cart_lossy", "cart_ragged", "cart_tight", "cart_tight0", "cart_nz_a", "cart_nz_b", "cart_nz_c", "cart_nz_d",
                                  "cart_nz_e", "cart_nz_f", "cart_long", "cart_open_halo",
                                  "fcc1_lossy", "fcc2_lossy", "fcc1_wide", "fcc2_wide"))
def test_full_state_bit_exact_from_noise(name, precision):
    """whole grids + boundary ODE state after 25 steps from a random initial state: exercises every
    interior node, the halo mirrors, the ABC shell and the lossy walls at once"""
    sd = make_sim_data(name, precision)
    g1, g0 = noise_grids(sd)
    o = Oracle(sd)
    o.write_grid(1, g1)
    o.write_grid(0, g0)
    o.run_steps(0, 25)
    for ak, fuse, svc in _kernels(sd):
        with _engine(sd, ak, fuse, svc=svc) as e:
            if fuse and name in ("fcc1_lossy", "fcc2_lossy", "fcc1_wide", "fcc2_wide"):
                assert e.stat("fused") == 1 and e.stat("svc") == 1
            if svc and ak == 1 and name in ("cart_lossy", "cart_ragged", "cart_long", "cart_nz_a", "fcc1_lossy", "fcc2_lossy", "fcc2_wide"):
                # the service warp really is in use on the ordinary rooms (the fused step's lists always hold the shell's z faces)
                assert e.stat("svc") == 1 and (e.stat("svc_entries") > 0 or not fuse)
            e.write_grid(1, g1)
            e.write_grid(0, g0)
            e.run_steps(0, 25)
            for which in (1, 0):
                a, b = e.read_grid(which), o.read_grid(which)
                if fuse:
                    # the outer halo layer is scratch: the fused step mirrors it when a value is written,
                    # the reference before it is read, so only the nodes 1..N-2 are comparable
                    a, b = a[1:-1, 1:-1, 1:-1], b[1:-1, 1:-1, 1:-1]
                elif e.stat("zflip"):
                    # the unfused step whose shell kernel writes the z halos of the NEW state (k_abc `zf`): those two layers are ahead
                    # of the reference's, which mirrors them at the start of the next step; everything else, halo rows and planes
                    # included, is comparable
                    a, b = a[:, :, 1:-1], b[:, :, 1:-1]
                assert np.array_equal(a, b), f"{name} p{precision} ak={ak} fuse={fuse} svc={svc} grid{which}: {np.abs(a - b).max():.3e}"
            v, g = e.read_boundary_state()
            vo, go = o.read_boundary_state()
            assert np.array_equal(v, vo) and np.array_equal(g, go)


# tile configurations of the TMA kernel: (id, lanes along z); 0/8/9 = 7-point defaults, 5/10/11 = FCC defaults (air_tma.cuh)
# 12/13/14 = the 7-point kernel with the service warp
TILE_CFGS = {"cart": ((0, 32), (8, 16), (9, 8), (12, 32), (13, 16), (14, 8)), "fcc": ((5, 32), (10, 16), (11, 8), (15, 32), (16, 16), (17, 8))}


@pytest.mark.parametrize("precision", (2, 1))
@pytest.mark.parametrize("name", ("cart_lossy_mb11", "cart_ragged", "cart_wide", "cart_tight", "cart_tight0", "cart_nz_a", "cart_nz_b", "cart_nz_c",
                                  "cart_nz_d", "cart_nz_e", "cart_nz_f", "fcc1_lossy", "fcc2_lossy", "fcc1_wide", "fcc2_wide"))
def test_every_tile_width_gives_the_same_bits(name, precision):
    """32, 16 and 8 lanes along z (the narrow tiles serve grids whose Nz is not a multiple of 32 vectors), fused and not:
    whole grids + boundary state after 25 steps from noise, against the oracle"""
    sd = make_sim_data(name, precision)
    g1, g0 = noise_grids(sd)
    o = Oracle(sd)
    o.write_grid(1, g1)
    o.write_grid(0, g0)
    o.run_steps(0, 25)
    want = [o.read_grid(1)[1:-1, 1:-1, 1:-1], o.read_grid(0)[1:-1, 1:-1, 1:-1]]
    vo, go = o.read_boundary_state()
    for cfg, lz in TILE_CFGS["cart" if sd.fcc_flag == 0 else "fcc"]:
        for fuse in (0, 1):
            with _engine(sd, 1, fuse, cfg) as e:
                if fuse and (sd.fcc_flag == 0 or cfg >= 12):  # (the 13-point kernel fuses only with the service warp)
                    assert e.stat("fused") == (1 if _fused_expected(sd, lz) else 0)
                    # the lists exist exactly when the kernel has the warp, the step is fused and no boundary / source node is on the shell
                    assert e.stat("svc") == (1 if cfg >= 12 and _fused_expected(sd, lz) and e.stat("abc_disjoint") else 0) or not _fused_expected(sd, lz)
                e.write_grid(1, g1)
                e.write_grid(0, g0)
                e.run_steps(0, 25)
                for which in (1, 0):
                    a = e.read_grid(which)[1:-1, 1:-1, 1:-1]
                    assert np.array_equal(a, want[1 - which]), f"{name} p{precision} cfg={cfg} fuse={fuse} grid{which}: {np.abs(a - want[1 - which]).max():.3e}"
                v, g = e.read_boundary_state()
                assert np.array_equal(v, vo) and np.array_equal(g, go)


@pytest.mark.parametrize("precision", (2, 1))
def test_service_warp_density_threshold(precision):
    """svc_cap decides which tile-planes' boundary nodes the air kernel finishes itself: none (only the shell's z faces), a few, as many
    as a stage holds -- same bits every time; a room with solid blocks and three materials, and the lossy shoebox with whole tiles"""
    for name in ("cart_blobs", "cart_wide"):
        sd = make_sim_data(name, precision)
        g1, g0 = noise_grids(sd)
        o = Oracle(sd)
        o.write_grid(1, g1)
        o.write_grid(0, g0)
        o.run_steps(0, 25)
        vo, go = o.read_boundary_state()
        entries, left = [], []
        for cap in (0, 3, 17, 192):
            with Engine(sd) as e:
                e.set_option("svc_cap", cap)
                assert e.stat("svc") == 1 and e.stat("svc_entries") > 0
                entries.append(e.stat("svc_entries"))
                left.append(e.stat("nb_left"))
                e.write_grid(1, g1)
                e.write_grid(0, g0)
                e.run_steps(0, 25)
                for which in (1, 0):
                    assert np.array_equal(e.read_grid(which)[1:-1, 1:-1, 1:-1], o.read_grid(which)[1:-1, 1:-1, 1:-1]), (name, cap, which)
                v, g = e.read_boundary_state()
                assert np.array_equal(v, vo) and np.array_equal(g, go)
        assert left[0] == sd.Nb and left == sorted(left, reverse=True) and entries == sorted(entries)
        if name == "cart_wide":
            assert left[-1] < sd.Nb  # the z walls of a wide grid are sparse in every tile-plane


def test_step_host_matches_run_steps():
    sd = make_sim_data("cart_lossy", 1)
    ref = Oracle(sd).run_all()
    with Engine(sd) as e:
        cols = [e.step_host(n, sd.in_sigs[:, n]).copy() for n in range(sd.Nt)]
    assert np.array_equal(np.stack(cols, axis=1), ref)


def test_step_host_mixes_with_run_steps_and_uploaded_signals():
    """host-driven steps (plain launches, then the captured copy+step+copy graph) interleaved with device-driven
    batches, with and without host source samples: one trace, bit for bit"""
    sd = make_sim_data("cart_lossy_mb11", 2)
    ref = Oracle(sd).run_all()
    with Engine(sd) as e:
        cols = []
        n = 0
        while n < sd.Nt:
            if n % 10 < 6:
                cols.append(e.step_host(n, sd.in_sigs[:, n] if n % 3 else None).copy())
                n += 1
            else:
                k = min(4, sd.Nt - n)
                e.run_steps(n, k)
                cols.extend(e.read_outputs(n, n + k).T.copy())
                n += k
        whole = e.read_outputs()
    assert np.array_equal(np.stack(cols, axis=1), ref) and np.array_equal(whole, ref)


@pytest.mark.parametrize("precision", (2, 1))
def test_no_receivers_no_sources_no_steps(precision):
    """empty receiver / source lists and a zero-step run are legal descriptions"""
    from dataclasses import replace
    sd = make_sim_data("cart_lossy", precision)
    z = np.zeros(0, np.int64)
    quiet = replace(sd, in_ixyz=z, in_sigs=np.zeros((0, sd.Nt)), _keep=[])
    with Engine(quiet) as e:
        e.run_steps(0, quiet.Nt)
        assert np.array_equal(e.read_outputs(), np.zeros((sd.Nr, sd.Nt)))
    deaf = replace(sd, out_ixyz=z, out_reorder=z, _keep=[])
    g1, g0 = noise_grids(sd)
    o = Oracle(deaf)
    with Engine(deaf) as e:
        for x in (o, e):
            x.write_grid(1, g1)
            x.write_grid(0, g0)
            x.run_steps(0, 20)
        assert e.read_outputs().shape == (0, sd.Nt)
        assert np.array_equal(e.read_grid(1)[1:-1, 1:-1, 1:-1], o.read_grid(1)[1:-1, 1:-1, 1:-1])
    with Engine(sd) as e:
        e.run_steps(0, 0)
        assert not e.read_outputs().any()


def test_run_sim_entry_point():
    sd = make_sim_data("cart_rigid", 2)
    out, t = run_sim(sd)
    assert np.array_equal(out, Oracle(sd).run_all()) and t > 0


def test_batched_equals_single_steps():
    sd = make_sim_data("cart_hann", 2)
    with Engine(sd) as a, Engine(sd) as b:
        a.run_steps(0, sd.Nt)
        for n in range(0, sd.Nt, 7):
            b.run_steps(n, min(7, sd.Nt - n))
        assert np.array_equal(a.read_outputs(), b.read_outputs())


def test_linearity_in_the_source():
    """doubling the input doubles every trace exactly (all operations are linear and scaling by 2 is exact)"""
    sd = make_sim_data("cart_lossy", 2)
    with Engine(sd) as e:
        e.run_steps(0, sd.Nt)
        u = e.read_outputs()
    sd2 = make_sim_data("cart_lossy", 2)
    sd2.in_sigs = sd2.in_sigs * 2.0
    with Engine(sd2) as e:
        e.run_steps(0, sd2.Nt)
        u2 = e.read_outputs()
    assert np.array_equal(u2, 2.0 * u)


def test_errors_are_reported_not_fatal():
    sd = make_sim_data("cart_rigid", 2)
    with Engine(sd) as e:
        with pytest.raises(PffdtdError):
            e.run_steps(0, sd.Nt + 1)
        with pytest.raises(PffdtdError):
            e.set_option("no_such_option", 1)
    bad = make_sim_data("cart_rigid", 2)
    bad.bn_ixyz = bad.bn_ixyz.copy()
    bad.bn_ixyz[0] = 0  # on the halo layer
    with pytest.raises(PffdtdError):
        Engine(bad)
    with pytest.raises(PffdtdError):
        Engine(sd, device=99)


@pytest.mark.parametrize("precision", (1, 2))
def test_constant_divisor_division_is_exact(precision):
    """the fused absorbing shell divides by 1 + l*Q through a reciprocal and one exact-residual step; on 3e8
    pseudo-random numerators it must give the very bits of an IEEE division"""
    import ctypes as C
    from pffdtd_b200.engine import lib, _check
    L = lib()
    L.pffdtd_selftest.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int64, C.POINTER(C.c_int64)]
    bad = C.c_int64(-1)
    for l in (0.999 / np.sqrt(3.0), 0.5, 0.3141592653589793, 0.57):
        _check(L.pffdtd_selftest(0, float(l), precision, 100_000_000, C.byref(bad)))
        assert bad.value == 0, f"l={l}: {bad.value} mismatches"


@pytest.mark.parametrize("name,precision", (("cart_lossy_mb11", 1), ("cart_ragged", 2), ("fcc2_lossy", 1), ("fcc1_lossy", 2)))
def test_folder_in_sim_outs_out_equals_reference_files(tmp_path, name, precision):
    """the drop-in contract end to end: four .h5 files in (chunked + deflate, as sim_setup writes them),
    sim_outs.h5 out, equal to what the unmodified reference CPU engine wrote for the same folder"""
    from pathlib import Path
    from cases import make_files
    from pffdtd_b200 import h5lite, shoebox
    from pffdtd_b200.sim_fdtd import run_folder
    gold = np.load(Path(__file__).parent / "golden" / "traces_ref_cpu_engine.npz")[f"{name}_p{precision}"]
    shoebox.write_folder(make_files(name), tmp_path, compress=3)
    u = run_folder(tmp_path, precision=precision)
    assert np.array_equal(u, gold)
    on_disk = h5lite.File(tmp_path / "sim_outs.h5")["u_out"][...]
    assert on_disk.dtype == np.float64 and np.array_equal(on_disk, gold)
