"""Synthetic shoebox rooms in the reference's file format (SURVEY.md §8d "Synthetic inputs").

Produces exactly the datasets `sim_setup` writes for a box-shaped room inside the
grid: walls lie between node planes `w`|`w+1` and `N-2-w`|`N-1-w` on every axis; boundary nodes are
listed on BOTH sides of a wall with the link across it cut, as the voxeliser does
(python/voxelizer/vox_scene.py:230-232); inner-side nodes carry a material id (or -1 = rigid), the
outer side is rigid.  Source: one 8-node trilinear group near the room centre, impulse scaled like
sim_comms.py:95-104 and (optionally) differentiated like sim_comms.py:106-119.  Receivers: 8-node
groups along the room diagonal.  Everything is deterministic (no RNG).

fcc=False -> 7-point Cartesian files (fcc_flag 0); fcc=True -> 13-point FCC on the even-parity
sub-lattice (fcc_flag 1, all dims even).  `pffdtd_b200.folder_prep.fold_fcc` turns the latter into the
folded gpu-folder form (fcc_flag 2).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from . import h5lite

CART_OFFS = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], np.int64)
FCC_OFFS = np.array([[+1, +1, 0], [-1, -1, 0], [0, +1, +1], [0, -1, -1], [+1, 0, +1], [-1, 0, -1],
                     [+1, -1, 0], [-1, +1, 0], [0, +1, -1], [0, -1, +1], [+1, 0, -1], [-1, 0, +1]], np.int64)


def synthetic_materials(nmat: int, mb: int = 11):
    """deterministic, passive RLC triplets (D, E, F > 0), `mb` branches per material"""
    out = []
    for k in range(nmat):
        m = np.arange(mb, dtype=np.float64)
        D = 0.05 + 0.02 * m + 0.013 * k
        E = 20.0 + 7.0 * m + 3.0 * k
        F = 2.0e3 * (1.6 ** m) * (1.0 + 0.25 * k)
        out.append(np.stack([D, E, F], axis=1))
    return out


def _boundary_nodes(N, lo, hi, offs, fcc):
    """all nodes with at least one link crossing a wall of the box [lo,hi] (inclusive, per axis)"""
    Nx, Ny, Nz = N
    cand = []
    rng = [np.arange(max(1, lo[a] - 1), min(N[a] - 2, hi[a] + 1) + 1, dtype=np.int64) for a in range(3)]
    for a in range(3):
        near = np.array(sorted({lo[a] - 1, lo[a], hi[a], hi[a] + 1}), np.int64)
        near = near[(near >= 1) & (near <= N[a] - 2)]
        axes = [rng[0], rng[1], rng[2]]
        axes[a] = near
        g = np.meshgrid(*axes, indexing="ij")
        cand.append(np.stack([x.ravel() for x in g], axis=1))
    c = np.unique(np.concatenate(cand), axis=0)
    if fcc:
        c = c[(c.sum(axis=1) & 1) == 0]
    lo_, hi_ = np.array(lo), np.array(hi)
    inside = lambda p: np.all((p >= lo_) & (p <= hi_), axis=1)
    ins = inside(c)
    cut = np.zeros((c.shape[0], offs.shape[0]), bool)
    for j, o in enumerate(offs):
        cut[:, j] = inside(c + o) != ins
    keep = cut.any(axis=1)
    c, cut, ins = c[keep], cut[keep], ins[keep]
    lin = (c[:, 0] * Ny + c[:, 1]) * Nz + c[:, 2]
    return lin, ~cut, ins, c


def _boundary_nodes_dense(N, lo, hi, offs, fcc, obstacles):
    """like _boundary_nodes, for a room with solid blocks inside (small grids: dense occupancy array).  `obstacles` = inclusive
    index boxes (x0, x1, y0, y1, z0, z1) of solid nodes.  A link is cut where it joins an air node and a solid / outside node;
    both ends of a cut link are boundary nodes (as the voxeliser lists them, vox_scene.py:230-232)."""
    Nx, Ny, Nz = N
    air = np.zeros(N, bool)
    air[lo[0]:hi[0] + 1, lo[1]:hi[1] + 1, lo[2]:hi[2] + 1] = True
    for (x0, x1, y0, y1, z0, z1) in obstacles:
        air[x0:x1 + 1, y0:y1 + 1, z0:z1 + 1] = False
    g = np.meshgrid(np.arange(1, Nx - 1), np.arange(1, Ny - 1), np.arange(1, Nz - 1), indexing="ij")
    c = np.stack([x.ravel() for x in g], axis=1).astype(np.int64)
    if fcc:
        c = c[(c.sum(axis=1) & 1) == 0]
    ins = air[c[:, 0], c[:, 1], c[:, 2]]
    cut = np.zeros((c.shape[0], offs.shape[0]), bool)
    for j, o in enumerate(offs):
        q = c + o
        cut[:, j] = air[q[:, 0], q[:, 1], q[:, 2]] != ins
    keep = cut.any(axis=1)
    c, cut, ins = c[keep], cut[keep], ins[keep]
    lin = (c[:, 0] * Ny + c[:, 1]) * Nz + c[:, 2]
    return lin, ~cut, ins, c


def _cart_boundary_fast(N, lo, hi, x_range=None):
    """same result as _boundary_nodes for the Cartesian scheme, built face by face (large grids);
    `x_range=(x0,x1)` keeps only the nodes of planes x0 <= ix < x1 (one rank's slab of a huge grid)"""
    Nx, Ny, Nz = N
    idx, clr, ins = [], [], []
    x0, x1 = (0, Nx) if x_range is None else x_range
    r = [np.arange(lo[a], hi[a] + 1, dtype=np.int64) for a in range(3)]
    r[0] = r[0][(r[0] >= x0) & (r[0] < x1)]
    strides = (Ny * Nz, Nz, 1)
    for a in range(3):
        o1, o2 = [b for b in range(3) if b != a]
        g1, g2 = np.meshgrid(r[o1], r[o2], indexing="ij")
        face = g1.ravel() * strides[o1] + g2.ravel() * strides[o2]
        for side, (pin, pout) in enumerate(((lo[a], lo[a] - 1), (hi[a], hi[a] + 1))):
            toward_out = 2 * a + (1 if side == 0 else 0)     # bit of the cut link seen from inside
            toward_in = 2 * a + (0 if side == 0 else 1)
            if a != 0 or x0 <= pin < x1:
                idx.append(face + pin * strides[a]); clr.append(np.full(face.size, 1 << toward_out, np.uint16)); ins.append(np.ones(face.size, bool))
            if 1 <= pout <= N[a] - 2 and (a != 0 or x0 <= pout < x1):
                idx.append(face + pout * strides[a]); clr.append(np.full(face.size, 1 << toward_in, np.uint16)); ins.append(np.zeros(face.size, bool))
    if not idx:
        return np.zeros(0, np.int64), np.zeros((0, 6), bool), np.zeros(0, bool)
    idx, clr, ins = np.concatenate(idx), np.concatenate(clr), np.concatenate(ins)
    k = np.argsort(idx, kind="stable")
    idx, clr, ins = idx[k], clr[k], ins[k]
    first = np.flatnonzero(np.r_[True, idx[1:] != idx[:-1]])
    clr = np.bitwise_or.reduceat(clr, first)
    idx, ins = idx[first], ins[first]
    adj16 = (~clr) & np.uint16(0x3F)
    adj = ((adj16[:, None] >> np.arange(6, dtype=np.uint16)[None, :]) & 1).astype(bool)
    return idx, adj, ins


def plane_costs(Nx, Ny, Nz, wall_offset=3, rigid=False, mb=11):
    """per x-plane cost of a Cartesian shoebox for SimData.slab_planes(cost=...), without building the node lists of the whole
    grid: counts the boundary / lossy / shell nodes of the five kinds of planes (outside the room, the two wall planes of either side,
    a plane across the room) with the same builder make_shoebox uses"""
    from .sim_data import SimData
    N, w = (int(Nx), int(Ny), int(Nz)), int(wall_offset)
    lo, hi = [w + 1] * 3, [n - 2 - w for n in N]
    P = float(Ny * Nz)

    def counts(x):
        _, _, ins = _cart_boundary_fast(N, lo, hi, (x, x + 1))
        return float(ins.size), (0.0 if rigid else float(ins.sum()))
    kinds = {x: counts(x) for x in {lo[0] - 1, lo[0], lo[0] + 1, hi[0], hi[0] + 1}}
    cost = np.zeros(Nx)
    lossy = SimData.COST_BNL_BASE + SimData.COST_BNL_BRANCH * mb
    for x in range(1, Nx - 1):
        nb, nbl = kinds[x] if x in kinds else (kinds[lo[0] + 1] if lo[0] < x < hi[0] else (0.0, 0.0))
        nba = (Ny - 2) * (Nz - 2) if x in (1, Nx - 2) else 2 * (Ny - 2) + 2 * (Nz - 2) - 4
        cost[x] = P + SimData.COST_BN * nb + lossy * nbl + SimData.COST_BNA * nba
    return cost


def make_shoebox(Nx, Ny, Nz, Nt, *, fcc=False, wall_offset=3, nmat=1, mb=11, rigid=False, diff=True,
                 h=0.05, c=343.0, nrec=3, sig="impulse", fast=None, x_range=None, obstacles=None):
    """-> dict of the four files' datasets: {'sim_consts': {...}, 'vox_out': {...}, 'comms_out': {...}, 'sim_mats': {...}}"""
    N = (int(Nx), int(Ny), int(Nz))
    w = int(wall_offset)
    if fcc and any(n % 2 for n in N):
        raise ValueError("FCC grids need even dims (cart_grid.py:35-38)")
    lo = [w + 1] * 3
    hi = [n - 2 - w for n in N]
    if any(hi[a] - lo[a] < 6 for a in range(3)):
        raise ValueError("grid too small for this wall offset")
    l = 0.999 if fcc else 0.999 / np.sqrt(3.0)      # sim_consts.py:34-40
    l2 = l * l
    Ts = h * l / c
    offs = FCC_OFFS if fcc else CART_OFFS
    if fast is None:
        fast = not fcc
    if x_range is not None and (fcc or not fast):
        raise ValueError("x_range needs the fast Cartesian builder")
    if obstacles:
        if x_range is not None:
            raise ValueError("obstacles need the dense builder (no x_range)")
        bn, adj, ins, _ = _boundary_nodes_dense(N, lo, hi, offs, fcc, obstacles)
    elif fast and not fcc:
        bn, adj, ins = _cart_boundary_fast(N, lo, hi, x_range)
    else:
        bn, adj, ins, _ = _boundary_nodes(N, lo, hi, offs, fcc)
    k = np.argsort(bn, kind="stable")
    bn, adj, ins = bn[k], adj[k], ins[k]
    ncut = (~adj).sum(axis=1)
    saf = ncut.astype(np.float64)
    if rigid or nmat == 0:
        mat = np.full(bn.size, -1, np.int8)
        nmat_eff = 0
    else:
        # material by position so that several materials are exercised; outer side and isolated nodes rigid
        ix = bn // (N[1] * N[2])
        mat = np.where(ins, (ix * nmat // N[0]).astype(np.int8), np.int8(-1)).astype(np.int8)
        mat[adj.sum(axis=1) == 0] = -1
        nmat_eff = nmat
    DEF = synthetic_materials(nmat_eff, mb)

    # source: 8-node group near the centre (plain air nodes), trilinear weights of an off-node point
    step = 2 if fcc else 1
    cx, cy, cz = [(lo[a] + hi[a]) // 2 for a in range(3)]
    if fcc and (cx + cy + cz) & 1:
        cz += 1
    fx, fy, fz = 0.3, 0.6, 0.2
    in_ixyz, in_alpha = [], []
    for i, wx in ((0, 1 - fx), (1, fx)):
        for j, wy in ((0, 1 - fy), (1, fy)):
            for kk, wz in ((0, 1 - fz), (1, fz)):
                in_ixyz.append(((cx + step * i) * N[1] + cy + step * j) * N[2] + cz + step * kk)
                in_alpha.append(wx * wy * wz)
    in_ixyz = np.array(in_ixyz, np.int64)
    in_alpha = np.array(in_alpha)
    s = np.zeros(Nt)
    if sig == "impulse":
        s[0] = 1.0
    elif sig == "hann10":
        n = np.arange(10)
        s[:10] = 0.5 * (1.0 - np.cos(2 * np.pi * n / 10))
    else:
        raise ValueError(sig)
    in_sigs = in_alpha[:, None] * s[None, :]
    in_sigs *= (0.5 if fcc else 1.0) * l2 / h
    if diff:   # bilinear differentiator b = 2/Ts [1,-1], a = [1,1]
        y = np.zeros_like(in_sigs)
        prev_x = np.zeros(in_sigs.shape[0])
        prev_y = np.zeros(in_sigs.shape[0])
        for n in range(Nt):
            y[:, n] = (2.0 / Ts) * (in_sigs[:, n] - prev_x) - prev_y
            prev_x, prev_y = in_sigs[:, n], y[:, n]
        in_sigs = y
    # receivers: 8-node groups along the diagonal, inside the room, away from walls
    out_ixyz = []
    for r in range(nrec):
        f = (r + 1) / (nrec + 1)
        p = [int(lo[a] + 2 + f * (hi[a] - lo[a] - 5)) for a in range(3)]
        if fcc and sum(p) & 1:
            p[2] += 1
        for i in (0, step):
            for j in (0, step):
                for kk in (0, step):
                    out_ixyz.append(((p[0] + i) * N[1] + p[1] + j) * N[2] + p[2] + kk)
    out_ixyz = np.array(out_ixyz, np.int64)
    assert not np.intersect1d(in_ixyz, bn).size and not np.intersect1d(out_ixyz, bn).size
    files = {
        "sim_consts": dict(c=np.float64(c), h=np.float64(h), Ts=np.float64(Ts), SR=np.float64(1 / Ts), l=np.float64(l),
                           l2=np.float64(l2), Tc=np.float64(20.0), rh=np.float64(50.0), fcc_flag=np.int8(1 if fcc else 0)),
        "vox_out": dict(Nx=np.int64(N[0]), Ny=np.int64(N[1]), Nz=np.int64(N[2]), Nb=np.int64(bn.size), bn_ixyz=bn, adj_bn=adj,
                        mat_bn=mat, saf_bn=saf, xv=np.arange(N[0]) * h, yv=np.arange(N[1]) * h, zv=np.arange(N[2]) * h,
                        h=np.float64(h)),
        "comms_out": dict(in_ixyz=in_ixyz, out_ixyz=out_ixyz, out_alpha=np.full((nrec, 8), 0.125),
                          out_reorder=np.arange(out_ixyz.size, dtype=np.int64), in_sigs=in_sigs, Ns=np.int64(in_ixyz.size),
                          Nr=np.int64(out_ixyz.size), Nt=np.int64(Nt), diff=np.int8(1 if diff else 0)),
        "sim_mats": dict(Nmat=np.int8(nmat_eff), Mb=np.array([d.shape[0] for d in DEF], np.int8),
                         **{f"mat_{i:02d}_DEF": d for i, d in enumerate(DEF)}),
    }
    return files


def write_folder(files: dict, data_dir, compress=None):
    """write the dict from make_shoebox as sim_consts.h5 / vox_out.h5 / comms_out.h5 / sim_mats.h5"""
    d = Path(data_dir)
    d.mkdir(parents=True, exist_ok=True)
    for stem, ds in files.items():
        h5lite.write_all(d / f"{stem}.h5", ds, compression=compress)
    return d


def sim_data_from_files(files: dict, precision: int, abc_x_range=None):
    """SimData straight from the in-memory dict (skips the disk round trip for large benchmarks)"""
    from .sim_data import SimData
    c, v, m, t = files["sim_consts"], files["vox_out"], files["comms_out"], files["sim_mats"]
    nm = int(t["Nmat"])
    return SimData.from_arrays(precision, fcc_flag=c["fcc_flag"], Nx=v["Nx"], Ny=v["Ny"], Nz=v["Nz"], l=c["l"], l2=c["l2"],
                               Ts=c["Ts"], bn_ixyz=v["bn_ixyz"], adj_bn=v["adj_bn"], mat_bn=v["mat_bn"], saf_bn=v["saf_bn"],
                               in_ixyz=m["in_ixyz"], out_ixyz=m["out_ixyz"], out_reorder=m["out_reorder"], in_sigs=m["in_sigs"],
                               Mb=t["Mb"], DEF=[t[f"mat_{i:02d}_DEF"] for i in range(nm)], diff=bool(m["diff"]),
                               abc_x_range=abc_x_range, h=c.get("h", 0.0), c=c.get("c", 0.0))
