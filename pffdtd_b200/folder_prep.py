"""'gpu folder' preparation: permute axes, fold the FCC grid, sort the node lists.

The reference does this to the files on disk (python/fdtd/rotate_sim_data.py:30-262, driven by
sim_setup.py:119-125); its GPU engine refuses un-prepared folders (gpu_engine.h:677,688).  Here the same
transforms work on the in-memory datasets (the dict layout of `shoebox.make_shoebox` / `load_folder`), so
an un-prepared CPU folder can be fed to the engine directly.  Checked against the reference functions
in tests/test_folder_prep.py (golden fixture generated with the reference under an h5py shim).
"""
from __future__ import annotations

import copy
from pathlib import Path

import numpy as np

from . import h5lite
from .shoebox import CART_OFFS, FCC_OFFS

STEMS = ("sim_consts", "vox_out", "comms_out", "sim_mats")


def load_folder(data_dir) -> dict:
    d = Path(data_dir)
    return {s: h5lite.read_all(d / f"{s}.h5") for s in STEMS}


def _sub(idx, Ny, Nz):
    iz = idx % Nz
    iy = (idx // Nz) % Ny
    ix = idx // (Nz * Ny)
    return ix, iy, iz


def rotate(files: dict, tr=None) -> dict:
    """permute the axes so that Nx >= Ny >= Nz (smallest halo plane, rotate_sim_data.py:42-43)"""
    f = copy.copy(files)
    v, m = dict(f["vox_out"]), dict(f["comms_out"])
    N = [int(v["Nx"]), int(v["Ny"]), int(v["Nz"])]
    if tr is None:
        tr = np.argsort(np.array(N))[::-1]
    tr = [int(t) for t in tr]
    if tr == [0, 1, 2]:
        return files
    Nt_ = [N[t] for t in tr]

    def relin(idx):
        s = _sub(np.asarray(idx, np.int64), N[1], N[2])
        return (s[tr[0]] * Nt_[1] + s[tr[1]]) * Nt_[2] + s[tr[2]]
    offs = FCC_OFFS if v["adj_bn"].shape[1] == 12 else CART_OFFS
    # column j of the new adjacency = old column whose direction, permuted, is direction j
    perm = offs[:, tr]
    src = [int(np.flatnonzero((perm == o).all(axis=1))[0]) for o in offs]
    v["adj_bn"] = np.ascontiguousarray(v["adj_bn"][:, src])
    v["bn_ixyz"] = relin(v["bn_ixyz"])
    m["in_ixyz"] = relin(m["in_ixyz"])
    m["out_ixyz"] = relin(m["out_ixyz"])
    v["Nx"], v["Ny"], v["Nz"] = (np.int64(n) for n in Nt_)
    ax = [v["xv"], v["yv"], v["zv"]]
    v["xv"], v["yv"], v["zv"] = (ax[t] for t in tr)
    f["vox_out"], f["comms_out"] = v, m
    return f


def fold_fcc(files: dict) -> dict:
    """fold the upper half of y onto the lower half: fcc_flag 1 -> 2 (rotate_sim_data.py:191-262)"""
    f = copy.copy(files)
    c, v, m = dict(f["sim_consts"]), dict(f["vox_out"]), dict(f["comms_out"])
    if int(c["fcc_flag"]) != 1:
        raise ValueError("fold_fcc needs fcc_flag == 1")
    Nx, Ny, Nz = int(v["Nx"]), int(v["Ny"]), int(v["Nz"])
    if Ny % 2:
        raise ValueError("Ny must be even")
    Nyh = Ny // 2 + 1

    def fold(idx):
        ix, iy, iz = _sub(np.asarray(idx, np.int64), Ny, Nz)
        up = iy >= Ny // 2
        return (ix * Nyh + np.where(up, Ny - 1 - iy, iy)) * Nz + iz, up
    bn, up = fold(v["bn_ixyz"])
    adj = np.array(v["adj_bn"], copy=True)
    for a, b in ((0, 6), (1, 7), (2, 9), (3, 8)):   # the y-mirrored partner of each direction
        tmp = adj[up, a].copy()
        adj[up, a] = adj[up, b]
        adj[up, b] = tmp
    v["bn_ixyz"], v["adj_bn"], v["Ny"] = bn, adj, np.int64(Nyh)
    m["in_ixyz"] = fold(m["in_ixyz"])[0]
    m["out_ixyz"] = fold(m["out_ixyz"])[0]
    c["fcc_flag"] = np.int8(2)
    f["sim_consts"], f["vox_out"], f["comms_out"] = c, v, m
    return f


def sort(files: dict) -> dict:
    """ascending node lists + out_reorder (rotate_sim_data.py:131-189)"""
    f = copy.copy(files)
    v, m = dict(f["vox_out"]), dict(f["comms_out"])
    # the same numpy calls as the reference (default sort kind), so that ties between receiver nodes that share
    # a grid node come out in the same order
    k = np.argsort(v["bn_ixyz"])
    for n in ("bn_ixyz", "adj_bn", "mat_bn", "saf_bn"):
        v[n] = np.ascontiguousarray(v[n][k])
    k = np.argsort(m["in_ixyz"])
    m["in_ixyz"], m["in_sigs"] = m["in_ixyz"][k], np.ascontiguousarray(m["in_sigs"][k])
    k = np.argsort(m["out_ixyz"])
    m["out_ixyz"] = m["out_ixyz"][k]
    m["out_reorder"] = np.argsort(k).astype(np.int64)
    f["vox_out"], f["comms_out"] = v, m
    return f


def gpu_folder(files: dict) -> dict:
    """the full save_folder_gpu pipeline of sim_setup.py:119-125: rotate, fold (if FCC), sort"""
    f = rotate(files)
    if int(f["sim_consts"]["fcc_flag"]) == 1:
        f = fold_fcc(f)
    return sort(f)
