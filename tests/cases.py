"""Shared problem builders for the parity tests (deterministic, no RNG)."""
import numpy as np

from pffdtd_b200 import folder_prep, shoebox
from pffdtd_b200.sim_data import SimData

# name -> (make_shoebox kwargs, layout) ; layout: "cart", "fcc1" (checkerboard), "fcc2" (folded gpu folder)
CASES = {
    "cart_rigid": (dict(Nx=20, Ny=18, Nz=16, Nt=60, rigid=True), "cart"),
    "cart_lossy": (dict(Nx=24, Ny=20, Nz=18, Nt=60, nmat=2, mb=3), "cart"),
    "cart_lossy_mb11": (dict(Nx=30, Ny=22, Nz=37, Nt=50, nmat=3, mb=11), "cart"),
    "cart_ragged": (dict(Nx=19, Ny=41, Nz=131, Nt=40, nmat=1, mb=4), "cart"),
    "cart_wide": (dict(Nx=17, Ny=70, Nz=270, Nt=30, nmat=2, mb=2), "cart"),
    "cart_hann": (dict(Nx=22, Ny=22, Nz=22, Nt=60, nmat=1, mb=5, sig="hann10"), "cart"),
    # walls one node from the absorbing shell: boundary nodes sit at index 2 / N-3 (late halo mirrors) and next to the shell
    "cart_tight": (dict(Nx=21, Ny=23, Nz=34, Nt=80, nmat=2, mb=3, wall_offset=1), "cart"),
    "cart_tight0": (dict(Nx=16, Ny=15, Nz=14, Nt=80, nmat=1, mb=2, wall_offset=0), "cart"),
    # every alignment of the high z end inside a 16-byte vector (fp32: Nz mod 4, fp64: Nz mod 2), > 1 z tile
    "cart_nz_a": (dict(Nx=16, Ny=21, Nz=133, Nt=40, nmat=1, mb=2), "cart"),
    "cart_nz_b": (dict(Nx=16, Ny=19, Nz=134, Nt=40, nmat=1, mb=2), "cart"),
    "cart_nz_c": (dict(Nx=16, Ny=18, Nz=135, Nt=40, nmat=1, mb=2), "cart"),
    "cart_nz_d": (dict(Nx=16, Ny=20, Nz=68, Nt=40, nmat=1, mb=2), "cart"),
    # the shell node z = Nz-2 opens a z tile (tile width 128 nodes in fp32, 64 in fp64): its mirror source z = Nz-3 lives in the tile before
    "cart_nz_e": (dict(Nx=16, Ny=19, Nz=130, Nt=40, nmat=1, mb=2), "cart"),
    "cart_nz_f": (dict(Nx=16, Ny=18, Nz=66, Nt=40, nmat=1, mb=2), "cart"),
    "fcc1_lossy": (dict(Nx=24, Ny=20, Nz=18, Nt=50, fcc=True, nmat=2, mb=3), "fcc1"),
    "fcc2_lossy": (dict(Nx=24, Ny=20, Nz=18, Nt=50, fcc=True, nmat=2, mb=3), "fcc2"),
    "fcc2_wide": (dict(Nx=18, Ny=44, Nz=150, Nt=30, fcc=True, nmat=1, mb=2), "fcc2"),
    "fcc1_wide": (dict(Nx=16, Ny=38, Nz=140, Nt=30, fcc=True, nmat=1, mb=2), "fcc1"),
    "fcc2_rigid": (dict(Nx=26, Ny=24, Nz=40, Nt=40, fcc=True, rigid=True), "fcc2"),
    # edge cases: the maximum branch count (MMb = 12, fdtd_data.h:33); a room without any boundary node (empty lists: only the
    # absorbing shell acts); one plane more than the minimum the voxeliser can produce in x
    "cart_mb12": (dict(Nx=22, Ny=20, Nz=26, Nt=50, nmat=2, mb=12), "cart"),
    "cart_empty": (dict(Nx=18, Ny=17, Nz=35, Nt=70, nmat=1, mb=2, _empty=True), "cart"),
    "fcc2_empty": (dict(Nx=20, Ny=18, Nz=34, Nt=50, fcc=True, nmat=1, mb=2, _empty=True), "fcc2"),
}


# rooms with solid blocks inside (a pillar, a one-node-thick plate, single isolated voxels (K = 0), an L-shaped block):
# boundary nodes away from the walls, every adjacency pattern a staircase produces
_BLOBS = [(6, 7, 14, 18, 6, 9), (20, 20, 6, 12, 8, 14), (10, 10, 16, 16, 18, 18), (22, 22, 8, 8, 20, 20), (12, 15, 5, 7, 18, 22), (12, 13, 8, 10, 18, 22)]
OBSTACLE_CASES = {
    "cart_blobs": (dict(Nx=30, Ny=24, Nz=28, Nt=70, nmat=3, mb=4, obstacles=_BLOBS), "cart"),
    "fcc1_blobs": (dict(Nx=30, Ny=24, Nz=28, Nt=60, fcc=True, nmat=2, mb=3, obstacles=_BLOBS), "fcc1"),
    "fcc2_blobs": (dict(Nx=30, Ny=24, Nz=28, Nt=60, fcc=True, nmat=2, mb=3, obstacles=_BLOBS), "fcc2"),
}


def _open_halo_links(files):
    """interior nodes of the walls at z = 1 / z = Nz-2 (wall_offset 0: their only cut link points at the halo): every third one gets that
    link OPEN and its +x link cut instead -- a masked node on the z shell that reads the z halo, next to a plain air node at z = 2 / Nz-3
    (the adjacency is not symmetric any more; the engines do not care)"""
    v = files["vox_out"]
    Nx, Ny, Nz = int(v["Nx"]), int(v["Ny"]), int(v["Nz"])
    bn, adj = v["bn_ixyz"], v["adj_bn"].copy()
    iz, iy, ix = bn % Nz, (bn // Nz) % Ny, bn // (Ny * Nz)
    inner = (ix >= 3) & (ix <= Nx - 4) & (iy >= 3) & (iy <= Ny - 4) & ((ix + iy) % 3 == 0)
    lo, hi = inner & (iz == 1), inner & (iz == Nz - 2)
    assert lo.sum() > 10 and hi.sum() > 10 and not adj[lo, 5].any() and not adj[hi, 4].any()
    adj[lo, 5], adj[hi, 4] = True, True
    adj[lo | hi, 0] = False
    v["adj_bn"] = adj
    v["saf_bn"] = (~adj).sum(axis=1).astype(np.float64)
    return files


# cases without golden traces of the reference engine (checked against the oracle only)
EXTRA_CASES = {
    # more planes than round 1's 96-chunk work plan covered in 16-plane chunks (1536): every x-chunk of the arithmetic plan, the guided tail
    "cart_long": (dict(Nx=1700, Ny=16, Nz=24, Nt=30, nmat=1, mb=2), "cart"),
    # masked nodes ON the z shell with an open link to the z halo (the vector holding the halo is fully masked in fp64)
    "cart_open_halo": (dict(Nx=20, Ny=19, Nz=22, Nt=40, nmat=1, mb=2, wall_offset=0, _hook=_open_halo_links), "cart"),
}


def make_files(name):
    kw, layout = CASES[name] if name in CASES else (OBSTACLE_CASES[name] if name in OBSTACLE_CASES else EXTRA_CASES[name])
    kw = dict(kw)
    empty = kw.pop("_empty", False)
    hook = kw.pop("_hook", None)
    files = shoebox.make_shoebox(kw.pop("Nx"), kw.pop("Ny"), kw.pop("Nz"), kw.pop("Nt"), **kw)
    if hook:
        files = hook(files)
    if empty:  # no walls at all: Nb = Nbl = 0
        nn = 12 if kw.get("fcc") else 6
        files["vox_out"].update(Nb=np.int64(0), bn_ixyz=np.zeros(0, np.int64), adj_bn=np.zeros((0, nn), bool),
                                mat_bn=np.zeros(0, np.int8), saf_bn=np.zeros(0))
    if layout == "fcc2":
        files = folder_prep.gpu_folder(files)
    return files


def make_sim_data(name, precision, scale=True):
    files = make_files(name)
    if precision == 1 and not int(files["comms_out"]["diff"]):
        raise ValueError("fp32 needs diff")
    sd = shoebox.sim_data_from_files(files, precision)
    if scale:
        sd.scale_input()
    return sd


def noise_grids(sd, seed=1234):
    """random initial state on the interior nodes of both grids (energy / stress tests)"""
    rng = np.random.default_rng(seed)
    g = []
    for _ in range(2):
        a = np.zeros((sd.Nx, sd.Ny, sd.Nz))
        a[1:-1, 1:-1, 1:-1] = rng.uniform(-1, 1, (sd.Nx - 2, sd.Ny - 2, sd.Nz - 2))
        g.append(a.astype(sd.real).astype(np.float64))
    return g
