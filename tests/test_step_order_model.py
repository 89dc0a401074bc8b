"""A numpy model of the one re-ordering the unfused CUDA step makes to the reference's halo handling ("z halos
without a z kernel"): the reference mirrors the halos of the state it is about to read at the START of a step, in the order seam row
(folded FCC) -> z -> y -> x, each over whole faces (c_cuda/cpu_engine.h:135-172); the CUDA step writes the z halos of the rows that
carry absorbing-shell nodes (x in 1..Nx-2, y in 1..Ny-2) at the END of the step that produced the state (k_abc, kernels.cuh), and the
next step runs seam -> y -> x only.  The model proves the two orders leave identical arrays, halo edges and corners included, for
both layouts and for a slab whose x halo planes come from a neighbour -- and that the shortcut really needs the later passes (it is
not equal before them).  The GPU tests check the same thing through the kernels, bit for bit against the oracle."""
import numpy as np
import pytest


def seam(u):
    u[:, -1, :] = u[:, -2, :]


def flip_z(u, rows=None):
    v = u if rows is None else u[rows]
    v[..., 0] = v[..., 2]
    v[..., -1] = v[..., -3]


def flip_y(u, fold):
    u[:, 0, :] = u[:, 2, :]
    if not fold:
        u[:, -1, :] = u[:, -3, :]


def flip_x(u, lo=True, hi=True):
    if lo:
        u[0] = u[2]
    if hi:
        u[-1] = u[-3]


def reference_pass(u, fold, lo=True, hi=True):
    if fold:
        seam(u)
    flip_z(u)
    flip_y(u, fold)
    flip_x(u, lo, hi)
    return u


@pytest.mark.parametrize("fold", (False, True))
@pytest.mark.parametrize("shape", ((7, 8, 9), (12, 7, 7), (9, 10, 33)))
def test_z_halos_written_early_for_the_shell_rows_give_the_reference_halos(fold, shape):
    rng = np.random.default_rng(5)
    u = rng.standard_normal(shape)
    want = reference_pass(u.copy(), fold)
    got = u.copy()
    inner = (slice(1, -1), slice(1, -1))
    flip_z(got, inner)                      # end of the producing step: only rows with shell nodes
    assert not np.array_equal(got, want)    # (the halo rows / planes still hold stale values)
    if fold:
        seam(got)
    flip_y(got, fold)
    flip_x(got)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("fold", (False, True))
def test_a_slab_receives_its_x_halo_planes_with_the_senders_z_halos(fold):
    """two slabs of one grid: each writes the z halos of its owned shell rows, sends its edge planes (whole planes, as the exchange
    does), then mirrors seam / y locally over ALL its planes and x only at the global ends"""
    rng = np.random.default_rng(9)
    full = rng.standard_normal((14, 9, 11))
    want = reference_pass(full.copy(), fold)
    a, b = full[:8].copy(), full[6:].copy()   # planes 0..7 and 6..13: owned 1..6 and 7..12
    for s in (a, b):
        flip_z(s, (slice(1, -1), slice(1, -1)))
    a[-1], b[0] = b[1].copy(), a[-2].copy()    # the exchange of the new state's edge planes
    for s, lo, hi in ((a, True, False), (b, False, True)):
        if fold:
            seam(s)
        flip_y(s, fold)
        flip_x(s, lo, hi)
    assert np.array_equal(a, want[:8]) and np.array_equal(b, want[6:])
