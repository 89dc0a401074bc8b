"""The voxeliser's hot stage on the GPU.

`VoxScene.calc_adj` of the reference (python/voxelizer/vox_scene.py:95-440) finds the boundary nodes of the FDTD grid: for every
grid point near a triangle it casts a ray towards each neighbour and cuts the link where the ray meets the surface; then it assigns
materials and surface-area factors.  The ray casting is its cost (the "ray-tri checks" timer; 53 s for the Musikverein at h = 0.2 m on
6 processes).  `calc_adj(vs)` below does the same work for a reference `VoxScene` object `vs` -- same voxel grid, same
triangle lists -- with the ray casting in `pffdtd_vox_run` (libpffdtd_b200.so, one thread block per voxel) and sets the same four
attributes (`bn_ixyz, adj_bn, mat_bn, saf_bn`), so `vs.check_adj_full()` / `vs.save()` carry on unchanged:

    vox_scene = VoxScene(room_geo, cart_grid, vox_grid, fcc=fcc_flag)
    pffdtd_b200.vox_accel.calc_adj(vox_scene)          # instead of vox_scene.calc_adj(Nprocs=...)      (sim_setup.py:111)

The stage before it, `VoxGridBase.fill` (python/voxelizer/vox_grid_base.py:67-176: which triangles meet which voxel, a triangle /
box overlap test per pair, minutes at production size), has the same treatment: `fill(vox_grid)` instead of
`vox_grid.fill(Nprocs=...)` (sim_setup.py:105) leaves the same `tri_idxs / tris_pre / tris_mat` on every voxel and the same
`nonempty_idx`, from `pffdtd_voxfill_run` (one warp per voxel).

No CPU fallback: without the library or a CUDA device it raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

R_EPS = 1e-6  # vox_scene.py:60


class pffdtd_vox_desc(C.Structure):
    """include/pffdtd_b200.h: pffdtd_vox_desc"""
    _fields_ = [("struct_size", C.c_int32), ("NN", C.c_int32), ("fcc", C.c_int32), ("reserved", C.c_int32),
                ("Nx", C.c_int64), ("Ny", C.c_int64), ("Nz", C.c_int64),
                ("xv", C.c_void_p), ("yv", C.c_void_p), ("zv", C.c_void_p),
                ("hf", C.c_double), ("c_bb", C.c_double), ("c_near", C.c_double), ("c_far", C.c_double),
                ("d_eps", C.c_double), ("cp_eps", C.c_double),
                ("vvh", C.c_void_p), ("ray_un", C.c_void_p),
                ("Nvox", C.c_int64), ("vox_start", C.c_void_p), ("vox_shape", C.c_void_p), ("vox_tri_off", C.c_void_p), ("vox_tri", C.c_void_p),
                ("Ntris", C.c_int64), ("unor", C.c_void_p), ("cent", C.c_void_p), ("bmin", C.c_void_p), ("bmax", C.c_void_p),
                ("v", C.c_void_p), ("eab", C.c_void_p), ("ebc", C.c_void_p), ("eca", C.c_void_p)]


class pffdtd_voxfill_desc(C.Structure):
    """include/pffdtd_b200.h: pffdtd_voxfill_desc"""
    _fields_ = [("struct_size", C.c_int32), ("reserved", C.c_int32), ("Nvox", C.c_int64), ("vbmin", C.c_void_p), ("vbmax", C.c_void_p),
                ("Ntris", C.c_int64), ("v", C.c_void_p), ("nor", C.c_void_p), ("cent", C.c_void_p), ("bmin", C.c_void_p), ("bmax", C.c_void_p)]


def fill_inputs_from_grid(vg) -> dict:
    """everything VoxGridBase.fill reads from a reference VoxGrid (also the layout of tests/golden/voxfill_*.npz)"""
    tp = vg.tris_pre
    return dict(vbmin=np.array([v.bmin for v in vg.voxels], np.float64).reshape(-1, 3),
                vbmax=np.array([v.bmax for v in vg.voxels], np.float64).reshape(-1, 3),
                v=np.ascontiguousarray(tp["v"]), nor=np.ascontiguousarray(tp["nor"]), cent=np.ascontiguousarray(tp["cent"]),
                bmin=np.ascontiguousarray(tp["bmin"]), bmax=np.ascontiguousarray(tp["bmax"]))


def fill_lists(inp: dict, device=0, host_lib=None):
    """-> (off [Nvox+1] int64, tri int32): the triangles meeting every voxel, ascending per voxel.  `host_lib`: the host checker
    (oracle/libvoxhost.so; tests only) instead of the GPU"""
    keep = [np.ascontiguousarray(inp[k], np.float64) for k in ("vbmin", "vbmax", "v", "nor", "cent", "bmin", "bmax")]
    d = pffdtd_voxfill_desc()
    d.struct_size = C.sizeof(pffdtd_voxfill_desc)
    d.Nvox, d.Ntris = int(keep[0].shape[0]), int(keep[3].shape[0])
    for k, a in zip(("vbmin", "vbmax", "v", "nor", "cent", "bmin", "bmax"), keep):
        setattr(d, k, a.ctypes.data if a.size else None)
    h = C.c_void_p()
    if host_lib is not None:
        host_lib.voxhost_fill.restype = C.c_void_p
        h = C.c_void_p(host_lib.voxhost_fill(C.byref(d)))
        if not h:
            raise RuntimeError("voxhost_fill failed")
        count, read, free = host_lib.voxhost_fill_count, host_lib.voxhost_fill_read, host_lib.voxhost_fill_free
    else:
        from .engine import _check, lib
        L = lib()
        L.pffdtd_voxfill_run.argtypes = [C.POINTER(pffdtd_voxfill_desc), C.c_int, C.POINTER(C.c_void_p)]
        _check(L.pffdtd_voxfill_run(C.byref(d), int(device), C.byref(h)))
        count, read, free = L.pffdtd_voxfill_count, L.pffdtd_voxfill_read, L.pffdtd_voxfill_free
    count.restype = C.c_int64
    count.argtypes = free.argtypes = [C.c_void_p]
    read.argtypes = [C.c_void_p] * 3
    off, tri = np.zeros(d.Nvox + 1, np.int64), np.zeros(max(int(count(h)), 0), np.int32)
    rc = read(h, off.ctypes.data, tri.ctypes.data if tri.size else None)
    free(h)
    if rc:
        raise RuntimeError("reading the voxel lists failed")
    return off, tri


def fill(vg, device: int = 0):
    """drop-in for VoxGridBase.fill (vox_grid_base.py:67-176): sets tri_idxs / tris_pre / tris_mat of every non-empty voxel and
    vg.nonempty_idx"""
    if vg.Nvox == 1:  # vox_grid_base.py:85-90
        vox = vg.voxels[0]
        vox.tri_idxs, vox.tris_pre, vox.tris_mat = np.arange(vg.Ntris), vg.tris_pre, vg.mats
        vg.nonempty_idx = [0]
        return vg
    off, tri = fill_lists(fill_inputs_from_grid(vg), device)
    vg.nonempty_idx = []
    for i in np.flatnonzero(np.diff(off) > 0):
        vox = vg.voxels[int(i)]
        vox.tri_idxs = tri[off[i]:off[i + 1]].astype(np.int64)
        vox.tris_pre, vox.tris_mat = vg.tris_pre[vox.tri_idxs], vg.mats[vox.tri_idxs]
        vg.nonempty_idx.append(int(i))
    return vg


def inputs_from_scene(vs) -> dict:
    """everything calc_adj reads from a reference VoxScene, as plain arrays (also the layout of tests/golden/vox_*.npz)"""
    cg, vg, rg = vs.cart_grid, vs.vox_grid, vs.room_geo
    tp = rg.tris_pre
    nonempty = [int(i) for i in vg.nonempty_idx]
    start = np.array([vg.voxels[i].ixyz_start for i in nonempty], np.int64).reshape(-1, 3)
    shape = np.array([vg.voxels[i].Nhxyz for i in nonempty], np.int64).reshape(-1, 3)
    lists = [np.asarray(vg.voxels[i].tri_idxs, np.int32) for i in nonempty]
    off = np.concatenate([[0], np.cumsum([a.size for a in lists])]).astype(np.int64)
    uvv = np.asarray(vs.uvv, np.float64)
    eps = np.finfo(np.float64).eps
    # normalise() of common/myfuncs.py:124-125, as tri_ray_intersection_vec applies it to the ray directions
    ray_un = (uvv.T / (np.sqrt(np.sum(uvv * uvv, axis=-1)) + eps)).T
    hf = float(vs.hf)
    return dict(
        NN=np.int32(vs.NN), fcc=np.int32(1 if vs.fcc else 0), Nxyz=np.asarray(cg.Nxyz, np.int64), h=np.float64(cg.h), hf=np.float64(hf),
        xv=np.asarray(cg.xv, np.float64), yv=np.asarray(cg.yv, np.float64), zv=np.asarray(cg.zv, np.float64),
        vvh=np.asarray(vs.vvh, np.float64), uvv=uvv, ray_un=np.ascontiguousarray(ray_un),
        vox_start=start, vox_shape=shape, vox_tri_off=off, vox_tri=np.concatenate(lists) if lists else np.zeros(0, np.int32),
        unor=np.ascontiguousarray(tp["unor"]), cent=np.ascontiguousarray(tp["cent"]), bmin=np.ascontiguousarray(tp["bmin"]),
        bmax=np.ascontiguousarray(tp["bmax"]), v=np.ascontiguousarray(tp["v"]), eab=np.ascontiguousarray(tp["eab_unor"]),
        ebc=np.ascontiguousarray(tp["ebc_unor"]), eca=np.ascontiguousarray(tp["eca_unor"]),
        mat_ind=np.asarray(rg.mat_ind), mat_side=np.asarray(rg.mat_side), Nmat=np.int64(rg.Nmat))


def make_desc(inp: dict):
    """-> (pffdtd_vox_desc, the arrays it points to)"""
    keep = []

    def p(a, dt):
        a = np.ascontiguousarray(a, dt)
        keep.append(a)
        return a.ctypes.data if a.size else None
    d = pffdtd_vox_desc()
    d.struct_size = C.sizeof(pffdtd_vox_desc)
    d.NN, d.fcc = int(inp["NN"]), int(inp["fcc"])
    d.Nx, d.Ny, d.Nz = (int(x) for x in inp["Nxyz"])
    d.xv, d.yv, d.zv = p(inp["xv"], np.float64), p(inp["yv"], np.float64), p(inp["zv"], np.float64)
    hf, h = float(inp["hf"]), float(inp["h"])
    # the very expressions of vox_scene.py:188, 218-228, 213 (python floats = IEEE doubles)
    d.hf, d.c_bb, d.c_near, d.c_far = hf, hf * (1 + R_EPS), R_EPS * hf, (1 + R_EPS) * hf
    d.d_eps, d.cp_eps = abs(1.0e-3 * h), 1e-6
    d.vvh, d.ray_un = p(inp["vvh"], np.float64), p(inp["ray_un"], np.float64)
    d.Nvox = int(inp["vox_start"].shape[0])
    d.vox_start, d.vox_shape = p(inp["vox_start"], np.int64), p(inp["vox_shape"], np.int64)
    d.vox_tri_off, d.vox_tri = p(inp["vox_tri_off"], np.int64), p(inp["vox_tri"], np.int32)
    d.Ntris = int(inp["unor"].shape[0])
    for k in ("unor", "cent", "bmin", "bmax", "v", "eab", "ebc", "eca"):
        setattr(d, k, p(inp[k], np.float64))
    return d, keep


def _run(L, prefix, d, NN, device=None):
    h = C.c_void_p()
    if device is None:  # host checker (oracle/libvoxhost.so)
        L.voxhost_run.restype = C.c_void_p
        h = C.c_void_p(L.voxhost_run(C.byref(d)))
        if not h:
            raise RuntimeError("voxhost_run failed")
        count, read, free = L.voxhost_count, L.voxhost_read, L.voxhost_free
    else:
        from .engine import _check
        _check(L.pffdtd_vox_run(C.byref(d), int(device), C.byref(h)))
        count, read, free = L.pffdtd_vox_count, L.pffdtd_vox_read, L.pffdtd_vox_free
    count.restype = C.c_int64
    count.argtypes = read.argtypes[:1] if getattr(read, "argtypes", None) else [C.c_void_p]
    free.argtypes = [C.c_void_p]
    read.argtypes = [C.c_void_p] * 5
    n = int(count(h))
    bn, adj = np.zeros(n, np.int64), np.zeros((n, NN), np.uint8)
    tidx, ndist = np.zeros(n, np.int32), np.zeros(n, np.float64)
    rc = read(h, bn.ctypes.data, adj.ctypes.data, tidx.ctypes.data, ndist.ctypes.data)
    free(h)
    if rc:
        raise RuntimeError("reading the voxeliser result failed")
    return bn, adj.astype(bool), tidx, ndist


def ray_stage(inp: dict, device: int = 0):
    """the ray casting on the GPU -> (bn_ixyz, adj_bn [Nb,NN] bool, tidx_bn, ndist_bn), the reference's order"""
    from .engine import lib
    L = lib()
    L.pffdtd_vox_run.argtypes = [C.POINTER(pffdtd_vox_desc), C.c_int, C.POINTER(C.c_void_p)]
    d, keep = make_desc(inp)
    return _run(L, "pffdtd_vox", d, int(inp["NN"]), device=device)


def finish(inp: dict, bn_ixyz, adj_bn, tidx_bn):
    """materials and surface-area factors of the boundary nodes (vox_scene.py:388-418): which side of its nearest triangle a
    node lies on decides whether a one-sided material applies; nodes lying on the surface are rigid; the area factor sums, per axis
    pair, |direction . normal| where either link of the pair is cut"""
    Nx, Ny, Nz = (int(x) for x in inp["Nxyz"])
    iz = bn_ixyz % Nz
    iy = (bn_ixyz - iz) // Nz % Ny
    ix = ((bn_ixyz - iz) // Nz - iy) // Ny
    xyz = np.c_[inp["xv"][ix], inp["yv"][iy], inp["zv"][iz]]
    unor = inp["unor"][tidx_bn]
    dv = np.sum((xyz - inp["cent"][tidx_bn]) * unor, axis=-1)
    side = inp["mat_side"][tidx_bn]
    mat_bn = np.array(inp["mat_ind"][tidx_bn])
    mat_bn[(dv > 0) & (side == 1)] = -1
    mat_bn[(dv < 0) & (side == 2)] = -1
    mat_bn[np.all(~adj_bn, axis=-1)] = -1
    saf_bn = np.zeros(bn_ixyz.size, np.float64)
    for j in range(0, int(inp["NN"]), 2):
        saf = np.abs(np.sum(inp["uvv"][j] * unor, axis=-1))
        saf_bn += (~adj_bn[:, j] | ~adj_bn[:, j + 1]) * saf
    return mat_bn, saf_bn


def calc_adj(vs, device: int = 0):
    """drop-in for VoxScene.calc_adj: sets vs.bn_ixyz / adj_bn / mat_bn / saf_bn"""
    inp = inputs_from_scene(vs)
    bn, adj, tidx, _ = ray_stage(inp, device)
    if np.unique(bn).size != bn.size:
        raise RuntimeError("a boundary node was found in two voxels")  # vox_scene.py:379
    vs.bn_ixyz, vs.adj_bn = bn, adj
    vs.mat_bn, vs.saf_bn = finish(inp, bn, adj, tidx)
    return vs
