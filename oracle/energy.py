"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference Python engine's energy balance
(python/fdtd/sim_fdtd.py:587-620, stencils :699-770, sums :838-856), evaluated on the
states of the C restatement (oracle.Oracle), one step at a time.  Small grids only.

Pinned by tests/golden/energy_ref_python_engine.npz: H_tot / E_lost / E_in written by the UNMODIFIED reference
Python engine run with energy_on=True (tests/golden/make_energy_golden.py).
"""
from __future__ import annotations

import numpy as np

from . import Oracle


def _offsets(sd):
    sx, sy = sd.Ny * sd.Nz, sd.Nz
    if sd.fcc_flag == 0:
        return [sx, -sx, sy, -sy, 1, -1]                      # sim_fdtd.py:764-770
    return [sx + sy, -sx - sy, sy + 1, -sy - 1, sx + 1, -sx - 1, sx - sy, -sx + sy, sy - 1, -sy + 1, sx - 1, -sx + 1]  # :741-753


def _mirror(u):
    """nb_flip_halos (sim_fdtd.py:772-789): z, then y, then x faces"""
    u[:, :, 0] = u[:, :, 2]
    u[:, :, -1] = u[:, :, -3]
    u[:, 0, :] = u[:, 2, :]
    u[:, -1, :] = u[:, -3, :]
    u[0] = u[2]
    u[-1] = u[-3]
    return u


def laplacian(sd, u, Lu):
    """Lu <- nb_stencil_air_* then nb_stencil_bn_* of the (mirrored) grid u; entries neither touches keep their value"""
    NN = sd.NN
    lfac = 0.25 if sd.fcc_flag else 1.0
    offs = _offsets(sd)
    Nx, Ny, Nz = sd.Nx, sd.Ny, sd.Nz
    ix, iy, iz = np.meshgrid(np.arange(1, Nx - 1), np.arange(1, Ny - 1), np.arange(1, Nz - 1), indexing="ij")
    c = (ix * Ny + iy) * Nz + iz
    air = np.ones(c.shape, bool)
    if sd.fcc_flag:
        air &= ((ix + iy + iz) % 2) == 0
    isbn = np.zeros(sd.Npts, bool)
    isbn[sd.bn_ixyz] = True
    air &= ~isbn[c]
    ca = c[air]
    uf = u.reshape(-1)
    s = -float(NN) * uf[ca]
    for o in offs:
        s = s + uf[ca + o]
    Lf = Lu.reshape(-1)
    Lf[ca] = lfac * s
    if sd.Nb:
        adj = np.stack([((sd.adj_bn >> j) & 1).astype(np.float64) for j in range(NN)], axis=1)
        b = sd.bn_ixyz
        s = -adj.sum(axis=1) * uf[b]
        for j, o in enumerate(offs):
            s = s + adj[:, j] * uf[b + o]
        Lf[b] = lfac * s
    return Lu


def energy_trace(sd, nsteps=None):
    """-> (H_tot[Nt], E_lost[Nt+1], E_in[Nt+1], u_out[Nr,Nt]) for the problem `sd` (fcc_flag 0 or 1)"""
    assert sd.fcc_flag in (0, 1) and sd.h > 0 and sd.c > 0
    Nt = sd.Nt if nsteps is None else nsteps
    o = Oracle(sd)
    V = 2.0 if sd.fcc_flag else 1.0
    h, c, l, l2, Ts = sd.h, sd.c, sd.l, sd.l2, sd.Ts
    bna, inx = sd.bna_ixyz, sd.in_ixyz
    Q = sd.Q_bna.astype(np.float64)
    V_bna = 2.0 ** -Q
    ssaf = sd.ssaf_bnl.astype(np.float64)[:, None]
    D = sd.DEF[sd.mat_bnl.astype(np.int64), :, 0] if sd.Nbl else np.zeros((0, 12))
    E = sd.DEF[sd.mat_bnl.astype(np.int64), :, 1] if sd.Nbl else np.zeros((0, 12))
    F = sd.DEF[sd.mat_bnl.astype(np.int64), :, 2] if sd.Nbl else np.zeros((0, 12))
    in_sigs = sd.in_sigs.astype(sd.real).astype(np.float64)
    H, lost, ein = np.zeros(Nt), np.zeros(Nt + 1), np.zeros(Nt + 1)
    Lu = np.zeros((sd.Nx, sd.Ny, sd.Nz))
    for n in range(Nt):
        u1, u2 = o.read_grid(1), o.read_grid(0)
        vh1, gh1 = o.read_boundary_state()
        term = ((u1 - u2) ** 2) / l2 - u1 * Lu
        H[n] = V * 0.5 * h * np.sum(term[1:-1, 1:-1, 1:-1])
        H[n] -= V * 0.5 * h * np.sum((1.0 - V_bna) * term.reshape(-1)[bna])
        H[n] += V * 0.5 * c / l2 * np.sum(ssaf * ((vh1 ** 2) * D + ((Ts * gh1) ** 2) * F))
        u2in, u2ba = u2.reshape(-1)[inx].copy(), u2.reshape(-1)[bna].copy()
        Lu = laplacian(sd, _mirror(u1.copy()), Lu)
        o.run_steps(n, 1)
        u0 = o.read_grid(1)  # grids are swapped at the end of the step: the new state
        vh0, _ = o.read_boundary_state()
        lost[n + 1] = lost[n] + V * 0.25 * h / l * np.sum(ssaf * (((vh0 + vh1) ** 2) * E))
        lost[n + 1] += 0.5 * V * h / l * np.sum((V_bna * Q) * (u0.reshape(-1)[bna] - u2ba) ** 2)
        ein[n + 1] = ein[n] + (V * h / l2) * 0.5 * np.sum((u0.reshape(-1)[inx] - u2in) * in_sigs[:, n])
    return H, lost, ein, o.u_out[:, :Nt].copy()
