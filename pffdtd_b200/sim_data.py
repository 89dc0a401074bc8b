"""Host-side problem description: the reference's ``struct SimData`` and everything that fills it.

Restates, in numpy, the host prep of the reference C engines so the device receives exactly the
numbers the reference would compute (validated field by field against the unmodified
``load_sim_data`` in tests/test_sim_data.py):

* ``SimData.load``           <- load_sim_data        c_cuda/fdtd_data.h:99-718
* ``SimData.scale_input``    <- scale_input          c_cuda/fdtd_data.h:879-909
* ``SimData.rescale_output`` <- rescale_output       c_cuda/fdtd_data.h:912-925
* ``SimData.write_outputs``  <- write_outputs        c_cuda/fdtd_data.h:928-980 / sim_fdtd.py:688-696
* ``SimData.sorted``         <- sort_sim_data        python/fdtd/rotate_sim_data.py:131-189
* ``SimData.slab``           <- split_data + per-GPU localisation  c_cuda/gpu_engine.h:516-662, 739-830

Index conventions are the reference's: grid Nx*Ny*Nz, z contiguous, ii = ix*Ny*Nz + iy*Nz + iz.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field, replace
from pathlib import Path

import numpy as np

from . import h5lite

MMB = 12  # fdtd_data.h:33
MNM = 64  # fdtd_data.h:35
EPS_F32 = 1.19209289e-07  # fdtd_common.h:66


def real_dtype(precision: int):
    if precision == 1:
        return np.float32
    if precision == 2:
        return np.float64
    raise ValueError("precision must be 1 (fp32) or 2 (fp64)")


class pffdtd_desc(C.Structure):
    """ctypes mirror of include/pffdtd_b200.h `pffdtd_desc`"""
    _fields_ = [
        ("struct_size", C.c_int32), ("precision", C.c_int32), ("fcc_flag", C.c_int32), ("Nm", C.c_int32),
        ("Nx", C.c_int64), ("Ny", C.c_int64), ("Nz", C.c_int64),
        ("Nb", C.c_int64), ("Nbl", C.c_int64), ("Nba", C.c_int64),
        ("Ns", C.c_int64), ("Nr", C.c_int64), ("Nt", C.c_int64),
        ("l", C.c_double), ("l2", C.c_double), ("a1", C.c_double), ("a2", C.c_double),
        ("sl2", C.c_double), ("lo2", C.c_double),
        ("ix0", C.c_int64), ("x_lo_edge", C.c_int32), ("x_hi_edge", C.c_int32),
        ("bn_ixyz", C.c_void_p), ("adj_bn", C.c_void_p), ("bnl_ixyz", C.c_void_p), ("mat_bnl", C.c_void_p),
        ("ssaf_bnl", C.c_void_p), ("bna_ixyz", C.c_void_p), ("Q_bna", C.c_void_p), ("in_ixyz", C.c_void_p),
        ("out_ixyz", C.c_void_p), ("in_sigs", C.c_void_p), ("Mb", C.c_void_p), ("mat_beta", C.c_void_p),
        ("mat_quads", C.c_void_p),
    ]


def abc_nodes(Nx, Ny, Nz, fcc_flag, ix_range=None):
    """ABC shell nodes and their Q (faces=1, edges=2, corners=3): fdtd_data.h:620-675.

    Enumerated in the reference's (ix, iy, iz) order over the *unfolded* y extent, parity-filtered
    for FCC, folded and sorted for fcc_flag==2.  `ix_range=(lo,hi)` restricts to planes lo<=ix<hi
    (used when building slabs of very large grids without materialising the global list)."""
    Nyf = 2 * (Ny - 1) if fcc_flag == 2 else Ny
    lo, hi = (1, Nx - 1) if ix_range is None else (max(1, ix_range[0]), min(Nx - 1, ix_range[1]))
    iy = np.arange(1, Nyf - 1, dtype=np.int64)
    iz = np.arange(1, Nz - 1, dtype=np.int64)
    qy = ((iy == 1) | (iy == Nyf - 2)).astype(np.int8)
    qz = ((iz == 1) | (iz == Nz - 2)).astype(np.int8)
    Qyz = qy[:, None] + qz[None, :]                       # [Nyf-2, Nz-2]
    IY, IZ = np.meshgrid(iy, iz, indexing="ij")
    if fcc_flag == 2:
        iyf = np.where(IY >= Nyf // 2, Nyf - IY - 1, IY)  # index on the folded grid
    else:
        iyf = IY
    lin = iyf * Nz + IZ
    par = (IY + IZ) & 1
    # a plane's selection depends on ix only through "is it an x face" and (FCC) the parity of ix: four patterns, made on demand
    pattern = {}

    def plane(qx, odd):
        if (qx, odd) not in pattern:
            sel = (Qyz + qx) > 0
            if fcc_flag > 0:
                sel &= ((par + odd) & 1) == 0
            pattern[(qx, odd)] = (lin[sel], (Qyz[sel] + qx).astype(np.int8))
        return pattern[(qx, odd)]
    idx_parts, q_parts = [], []
    for ix in range(lo, hi):
        l, q = plane(1 if (ix == 1 or ix == Nx - 2) else 0, ix & 1 if fcc_flag > 0 else 0)
        idx_parts.append(ix * Nz * Ny + l)
        q_parts.append(q)
    if idx_parts:
        bna = np.concatenate(idx_parts)
        Q = np.concatenate(q_parts)
    else:
        bna, Q = np.zeros(0, np.int64), np.zeros(0, np.int8)
    if fcc_flag == 2:
        k = np.argsort(bna, kind="stable")
        bna, Q = bna[k], Q[k]
    return bna, Q


def abc_count(Nx, Ny, Nz, fcc_flag):
    Nyf = 2 * (Ny - 1) if fcc_flag == 2 else Ny
    Nba = 2 * (Nx * Nyf + Nx * Nz + Nyf * Nz) - 12 * (Nx + Nyf + Nz) + 56
    return Nba // 2 if fcc_flag > 0 else Nba


@dataclass
class SimData:
    precision: int
    fcc_flag: int
    Nx: int
    Ny: int
    Nz: int
    Nt: int
    l: float
    l2: float
    Ts: float
    diff: bool
    bn_ixyz: np.ndarray        # int64 [Nb]
    adj_bn: np.ndarray         # uint16 [Nb] bit-packed adjacency
    mat_bnl: np.ndarray        # int8 [Nbl]
    ssaf_bnl: np.ndarray       # Real [Nbl]
    bnl_ixyz: np.ndarray       # int64 [Nbl]
    bna_ixyz: np.ndarray       # int64 [Nba]
    Q_bna: np.ndarray          # int8 [Nba]
    in_ixyz: np.ndarray        # int64 [Ns]
    out_ixyz: np.ndarray       # int64 [Nr]
    out_reorder: np.ndarray    # int64 [Nr]
    in_sigs: np.ndarray        # float64 [Ns, Nt]
    Mb: np.ndarray             # int8 [Nm]
    mat_quads: np.ndarray      # Real [Nm, MMB, 4]  (b, bd, bDh, bFh)
    mat_beta: np.ndarray       # Real [Nm]
    a1: float = 0.0            # Real values held exactly in Python floats
    a2: float = 0.0
    sl2: float = 0.0
    lo2: float = 0.0
    infac: float = 1.0
    # slab placement (whole grid by default)
    ix0: int = 0
    x_lo_edge: bool = True
    x_hi_edge: bool = True
    # only the energy balance needs these (sim_fdtd.py:59-95, 257-259): grid spacing, speed of sound, raw (D,E,F) triplets
    h: float = 0.0
    c: float = 0.0
    DEF: np.ndarray = None     # float64 [Nm, MMB, 3], zero padded
    _keep: list = field(default_factory=list, repr=False)

    # ---- derived
    @property
    def Npts(self): return self.Nx * self.Ny * self.Nz
    @property
    def Nb(self): return int(self.bn_ixyz.size)
    @property
    def Nbl(self): return int(self.bnl_ixyz.size)
    @property
    def Nba(self): return int(self.bna_ixyz.size)
    @property
    def Ns(self): return int(self.in_ixyz.size)
    @property
    def Nr(self): return int(self.out_ixyz.size)
    @property
    def Nm(self): return int(self.Mb.size)
    @property
    def NN(self): return 12 if self.fcc_flag > 0 else 6
    @property
    def real(self): return real_dtype(self.precision)

    @property
    def K_bn(self):
        """number of open links per boundary node (fdtd_data.h:553-560)"""
        a = self.adj_bn.astype(np.uint16)
        k = np.zeros(a.shape, np.int8)
        for j in range(self.NN):
            k += ((a >> j) & 1).astype(np.int8)
        return k

    @property
    def bn_mask(self):
        """bit mask over the grid, LSB first, 1 = boundary node (fdtd_data.h:563-572)"""
        m = np.zeros((self.Npts - 1) // 8 + 1, np.uint8)
        np.bitwise_or.at(m, self.bn_ixyz >> 3, (1 << (self.bn_ixyz & 7)).astype(np.uint8))
        return m

    # ---- construction
    @classmethod
    def from_arrays(cls, precision, *, fcc_flag, Nx, Ny, Nz, l, l2, Ts, bn_ixyz, adj_bn, mat_bn, saf_bn,
                    in_ixyz, out_ixyz, out_reorder, in_sigs, Mb, DEF, diff=True, abc_x_range=None, h=0.0, c=0.0):
        """Everything load_sim_data derives, from the raw file contents.
        `abc_x_range=(x0,x1)` builds the absorbing-shell list for those planes only (a rank that will keep just
        its slab of a very large grid need not enumerate the whole shell)."""
        R = real_dtype(precision)
        fcc_flag = int(fcc_flag)
        if not 0 <= fcc_flag <= 2:
            raise ValueError("fcc_flag must be 0, 1 or 2")
        Nx, Ny, Nz = int(Nx), int(Ny), int(Nz)
        l, l2, Ts = float(l), float(l2), float(Ts)
        if fcc_flag > 0:
            if not (l2 <= 1.0 and l <= 1.0):
                raise ValueError("FCC scheme needs l <= 1")
            NN = 12
        else:
            if not (l2 <= 1.0 / 3.0 and l <= np.sqrt(1.0 / 3.0)):
                raise ValueError("Cartesian scheme needs l <= sqrt(1/3)")
            NN = 6
        if precision == 1 and not diff:
            raise ValueError("single precision requires a differentiated input (fdtd_data.h:392)")
        # coefficients: fdtd_data.h:186-194 (double arithmetic, then one rounding to Real)
        eps = EPS_F32 if precision == 1 else 0.0
        lfac = 0.25 if fcc_flag > 0 else 1.0
        dsl2 = (1.0 + eps) * lfac * l2
        da1 = 2.0 - dsl2 * NN
        da2 = lfac * l2
        a1, a2, sl2, lo2 = float(R(da1)), float(R(da2)), float(R(dsl2)), float(R(0.5 * l))

        bn_ixyz = np.ascontiguousarray(bn_ixyz, np.int64)
        Nb = bn_ixyz.size
        adj_bool = np.asarray(adj_bn).astype(bool).reshape(Nb, NN)
        mat_bn = np.ascontiguousarray(mat_bn, np.int8)
        saf_bn = np.ascontiguousarray(saf_bn, np.float64)
        if mat_bn.size != Nb or saf_bn.size != Nb:
            raise ValueError("vox_out arrays disagree on Nb")
        # check_inside_grid (fdtd_common.h:84-102)
        iz = bn_ixyz % Nz
        iy = (bn_ixyz // Nz) % Ny
        ix = bn_ixyz // (Nz * Ny)
        if Nb and not (iz.min() > 0 and iy.min() > 0 and ix.min() > 0 and iz.max() < Nz - 1 and iy.max() < Ny - 1
                       and ix.max() < Nx - 1):
            raise ValueError("boundary node on the halo layer")
        # fdtd_data.h:516-526
        if Nb:
            if adj_bool.all(axis=1).any():
                raise ValueError("boundary node with all links open")
            if (mat_bn[~adj_bool.any(axis=1)] != -1).any():
                raise ValueError("isolated boundary node must be rigid")
        adj = np.zeros(Nb, np.uint16)
        for j in range(NN):
            adj |= adj_bool[:, j].astype(np.uint16) << np.uint16(j)
        # ssaf: fdtd_data.h:283-289 -- (Real)(0.5/sqrt(2)) * saf in double, then rounded to Real
        if fcc_flag > 0:
            ssaf = (np.float64(R(0.5 / np.sqrt(2.0))) * saf_bn).astype(R)
        else:
            ssaf = saf_bn.astype(R)
        lossy = mat_bn >= 0
        # materials: fdtd_data.h:424-460
        Mb = np.ascontiguousarray(Mb, np.int8).reshape(-1)
        Nm = Mb.size
        if Nm > MNM:
            raise ValueError("too many materials")
        quads = np.zeros((Nm, MMB, 4), R)
        beta = np.zeros(Nm, R)
        DEF_pad = np.zeros((Nm, MMB, 3), np.float64)
        for i in range(Nm):
            d = np.asarray(DEF[i], np.float64).reshape(-1, 3)
            if d.shape[0] != Mb[i] or Mb[i] > MMB:
                raise ValueError("bad DEF shape")
            D, E, F = d[:, 0], d[:, 1], d[:, 2]
            DEF_pad[i, :Mb[i]] = d
            Dh, Eh, Fh = D / Ts, E, F * Ts
            b = 1.0 / (2.0 * Dh + Eh + 0.5 * Fh)
            bd = b * (2.0 * Dh - Eh - 0.5 * Fh)
            if not (np.isfinite(b).all() and np.isfinite(bd).all()):
                raise ValueError("non-finite material coefficient")
            quads[i, :Mb[i], 0] = b
            quads[i, :Mb[i], 1] = bd
            quads[i, :Mb[i], 2] = b * Dh
            quads[i, :Mb[i], 3] = b * Fh
            acc = R(0.0)
            for j in range(Mb[i]):      # accumulated in Real, in branch order
                acc = R(acc + R(b[j]))
            beta[i] = acc
        if lossy.any() and int(mat_bn.max()) >= Nm:
            raise ValueError("material id out of range")
        bna, Q = abc_nodes(Nx, Ny, Nz, fcc_flag, ix_range=abc_x_range)
        assert abc_x_range is not None or bna.size == abc_count(Nx, Ny, Nz, fcc_flag)
        in_sigs = np.ascontiguousarray(in_sigs, np.float64)
        in_ixyz = np.ascontiguousarray(in_ixyz, np.int64)
        Nt = int(in_sigs.shape[1]) if in_sigs.ndim == 2 else 0
        return cls(precision=precision, fcc_flag=fcc_flag, Nx=Nx, Ny=Ny, Nz=Nz, Nt=Nt, l=l, l2=l2, Ts=Ts,
                   diff=bool(diff), bn_ixyz=bn_ixyz, adj_bn=adj, mat_bnl=np.ascontiguousarray(mat_bn[lossy]),
                   ssaf_bnl=np.ascontiguousarray(ssaf[lossy]), bnl_ixyz=np.ascontiguousarray(bn_ixyz[lossy]),
                   bna_ixyz=bna, Q_bna=Q, in_ixyz=in_ixyz, out_ixyz=np.ascontiguousarray(out_ixyz, np.int64),
                   out_reorder=np.ascontiguousarray(out_reorder, np.int64), in_sigs=in_sigs.reshape(in_ixyz.size, Nt),
                   Mb=Mb, mat_quads=quads, mat_beta=beta, a1=a1, a2=a2, sl2=sl2, lo2=lo2, h=float(h), c=float(c), DEF=DEF_pad)

    @classmethod
    def load(cls, data_dir, precision: int) -> "SimData":
        """Read the four sim_setup files from `data_dir` (the drop-in boundary)."""
        d = Path(data_dir)
        for fn in ("sim_consts.h5", "vox_out.h5", "comms_out.h5", "sim_mats.h5"):
            if not (d / fn).exists():
                raise FileNotFoundError(f"{fn} doesn't exist in {d}")
        c = h5lite.File(d / "sim_consts.h5")
        v = h5lite.File(d / "vox_out.h5")
        m = h5lite.File(d / "comms_out.h5")
        t = h5lite.File(d / "sim_mats.h5")
        Nb = int(v["Nb"][()])
        Ns, Nr, Nt = int(m["Ns"][()]), int(m["Nr"][()]), int(m["Nt"][()])
        Nmat = int(t["Nmat"][()])
        Mb = t["Mb"][...]
        sd = cls.from_arrays(
            precision, fcc_flag=c["fcc_flag"][()], Nx=v["Nx"][()], Ny=v["Ny"][()], Nz=v["Nz"][()],
            l=c["l"][()], l2=c["l2"][()], Ts=c["Ts"][()], bn_ixyz=v["bn_ixyz"][...], adj_bn=v["adj_bn"][...],
            mat_bn=v["mat_bn"][...], saf_bn=v["saf_bn"][...], in_ixyz=m["in_ixyz"][...], out_ixyz=m["out_ixyz"][...],
            out_reorder=m["out_reorder"][...], in_sigs=m["in_sigs"][...], Mb=Mb[:Nmat],
            DEF=[t[f"mat_{i:02d}_DEF"][...] for i in range(Nmat)], diff=bool(m["diff"][()]),
            h=c["h"][()] if "h" in c else 0.0, c=c["c"][()] if "c" in c else 0.0)
        if sd.Nb != Nb or sd.Ns != Ns or sd.Nr != Nr or sd.Nt != Nt:
            raise ValueError("dataset sizes disagree with the stored counts")
        return sd

    # ---- scaling and output (fdtd_data.h:879-980)
    def scale_input(self):
        max_in = float(np.max(np.abs(self.in_sigs))) if self.in_sigs.size else 0.0
        if max_in == 0.0:
            self.infac = 1.0
            return self
        fi = np.finfo(self.real)
        # REAL_MAX_EXP / REAL_MIN_EXP are the C <float.h> values (128/-125, 1024/-1021)
        pow2 = int(round(0.5 * (fi.maxexp) + 0.5 * (fi.minexp + 1)))
        norm1 = 2.0 ** pow2
        inv_infac = norm1 / max_in
        self.infac = 1.0 / inv_infac
        self.in_sigs = self.in_sigs * inv_infac
        return self

    def rescale_output(self, u_out):
        return u_out * self.infac

    def reorder_output(self, u_out):
        """rows of the file = internal rows permuted by out_reorder (fdtd_data.h:941-945)"""
        return u_out[self.out_reorder, :]

    def write_outputs(self, data_dir, u_out):
        f = h5lite.File(Path(data_dir) / "sim_outs.h5", "w")
        f.create_dataset("u_out", data=np.ascontiguousarray(self.reorder_output(u_out), np.float64))
        f.close()

    # ---- ordering and partitioning
    def is_sorted(self):
        s = lambda a, strict: a.size < 2 or bool(np.all(np.diff(a) > 0) if strict else np.all(np.diff(a) >= 0))
        # duplicates are legal among sources and receivers (split_data only walks the lists, gpu_engine.h:562-661)
        return (s(self.bn_ixyz, True) and s(self.bnl_ixyz, True) and s(self.bna_ixyz, True)
                and s(self.in_ixyz, False) and s(self.out_ixyz, False))

    def sorted(self) -> "SimData":
        """Ascending node lists, what sort_sim_data does to the files (rotate_sim_data.py:157-170);
        required by the slab split (gpu_engine.h:497-513) and good for locality on one GPU."""
        if self.is_sorted():
            return self
        kb = np.argsort(self.bn_ixyz, kind="stable")
        kl = np.argsort(self.bnl_ixyz, kind="stable")
        ka = np.argsort(self.bna_ixyz, kind="stable")
        ki = np.argsort(self.in_ixyz, kind="stable")
        ko = np.argsort(self.out_ixyz, kind="stable")
        inv = np.empty_like(ko)
        inv[ko] = np.arange(ko.size)
        return replace(self, bn_ixyz=self.bn_ixyz[kb], adj_bn=self.adj_bn[kb], bnl_ixyz=self.bnl_ixyz[kl],
                       mat_bnl=self.mat_bnl[kl], ssaf_bnl=self.ssaf_bnl[kl], bna_ixyz=self.bna_ixyz[ka],
                       Q_bna=self.Q_bna[ka], in_ixyz=self.in_ixyz[ki], in_sigs=self.in_sigs[ki],
                       out_ixyz=self.out_ixyz[ko], out_reorder=inv[self.out_reorder], _keep=[])

    @staticmethod
    def slab_planes(Nx, nranks, cost=None):
        """owned planes per rank.  `cost=None`: the reference's split, Nx//n each, +1 for the first Nx%n ranks
        (gpu_engine.h:532-543).  `cost` = per-plane cost [Nx] (see `plane_costs`): contiguous slabs of about equal total cost,
        so that ranks holding walls perpendicular to x get fewer planes.  Every slab keeps at least 2 planes; the arithmetic of
        a node does not depend on where the cuts are, so any split gives the same bits."""
        if cost is None:
            base, rem = divmod(Nx, nranks)
            sizes = [base + (1 if r < rem else 0) for r in range(nranks)]
        else:
            cost = np.asarray(cost, np.float64)
            if cost.shape != (Nx,) or np.any(cost < 0):
                raise ValueError("cost must be a non-negative array of Nx entries")
            if Nx < 2 * nranks:
                raise ValueError("too many ranks for this grid")
            cum = np.concatenate([[0.0], np.cumsum(cost)])
            cuts = [0]
            for r in range(1, nranks):
                x = int(np.searchsorted(cum, cum[-1] * r / nranks, "left"))
                # the nearer of the two planes around the target; leave >= 2 planes to this slab and to every later one
                if x > 0 and abs(cum[x - 1] - cum[-1] * r / nranks) <= abs(cum[x] - cum[-1] * r / nranks):
                    x -= 1
                x = min(max(x, cuts[-1] + 2), Nx - 2 * (nranks - r))
                cuts.append(x)
            cuts.append(Nx)
            sizes = [cuts[r + 1] - cuts[r] for r in range(nranks)]
        starts = [sum(sizes[:r]) for r in range(nranks)]
        return starts, sizes

    # cost of one node of each kind in units of one air node (12.125 B of HBM traffic, fp32), from the measured kernel
    # times on B200 (profiles/): boundary node (rigid update + its share of the adjacency list), lossy node (11 branches of
    # state read and written), absorbing-shell node
    COST_BN, COST_BNL_BASE, COST_BNL_BRANCH, COST_BNA = 3.0, 4.0, 1.5, 2.0

    def plane_costs(self):
        """per x-plane cost [Nx] for `slab_planes(cost=...)`: air nodes + weighted node-list entries of the plane"""
        P = self.Ny * self.Nz
        cnt = lambda a: np.bincount(np.asarray(a, np.int64) // P, minlength=self.Nx).astype(np.float64)
        mb = float(np.max(self.Mb)) if np.size(self.Mb) else 0.0
        air = np.full(self.Nx, float(P))
        if self.x_lo_edge:
            air[0] = 0.0  # the global halo planes are never updated
        if self.x_hi_edge:
            air[-1] = 0.0
        return (air + self.COST_BN * cnt(self.bn_ixyz) + (self.COST_BNL_BASE + self.COST_BNL_BRANCH * mb) * cnt(self.bnl_ixyz)
                + self.COST_BNA * cnt(self.bna_ixyz))

    def slab(self, rank: int, nranks: int, planes=None) -> "SimData":
        """The part of the problem rank `rank` of `nranks` owns, re-based to slab-local indices, with
        one halo plane towards each neighbour (gpu_engine.h:755-823).  Receiver rows of all ranks
        concatenated in rank order give the internal (sorted) receiver order.  `planes` = (starts, sizes) of a
        split other than the reference's equal one (`slab_planes(cost=...)`)."""
        if nranks == 1:
            return self
        if not self.is_sorted():
            raise ValueError("slab split needs sorted node lists; call .sorted() first")
        starts, sizes = planes if planes is not None else self.slab_planes(self.Nx, nranks)
        if nranks > self.Nx - 2 or min(sizes) < 2 or sum(sizes) != self.Nx or len(sizes) != nranks:
            raise ValueError("too many ranks for this grid")
        P = self.Ny * self.Nz
        lo, hi = starts[rank] * P, (starts[rank] + sizes[rank]) * P
        first = starts[rank] - (1 if rank > 0 else 0)           # global index of local plane 0
        Nxh = sizes[rank] + (1 if rank > 0 else 0) + (1 if rank < nranks - 1 else 0)
        off = first * P

        def own(a):
            return slice(int(np.searchsorted(a, lo, "left")), int(np.searchsorted(a, hi, "left")))
        sb, sl, sa, si, so = own(self.bn_ixyz), own(self.bnl_ixyz), own(self.bna_ixyz), own(self.in_ixyz), own(self.out_ixyz)
        return replace(self, Nx=Nxh, bn_ixyz=self.bn_ixyz[sb] - off, adj_bn=self.adj_bn[sb],
                       bnl_ixyz=self.bnl_ixyz[sl] - off, mat_bnl=self.mat_bnl[sl], ssaf_bnl=self.ssaf_bnl[sl],
                       bna_ixyz=self.bna_ixyz[sa] - off, Q_bna=self.Q_bna[sa], in_ixyz=self.in_ixyz[si] - off,
                       in_sigs=self.in_sigs[si], out_ixyz=self.out_ixyz[so] - off,
                       out_reorder=np.arange(so.stop - so.start, dtype=np.int64), ix0=self.ix0 + first,
                       x_lo_edge=self.x_lo_edge and rank == 0, x_hi_edge=self.x_hi_edge and rank == nranks - 1, _keep=[])

    # ---- C ABI
    def desc(self) -> pffdtd_desc:
        """Fill a pffdtd_desc; the numpy buffers it points to are kept alive on `self`."""
        keep = []

        def p(a, dt):
            a = np.ascontiguousarray(a, dt)
            keep.append(a)
            return a.ctypes.data_as(C.c_void_p) if a.size else None
        d = pffdtd_desc()
        d.struct_size = C.sizeof(pffdtd_desc)
        d.precision, d.fcc_flag, d.Nm = self.precision, self.fcc_flag, self.Nm
        d.Nx, d.Ny, d.Nz = self.Nx, self.Ny, self.Nz
        d.Nb, d.Nbl, d.Nba, d.Ns, d.Nr, d.Nt = self.Nb, self.Nbl, self.Nba, self.Ns, self.Nr, self.Nt
        d.l, d.l2, d.a1, d.a2, d.sl2, d.lo2 = self.l, self.l2, self.a1, self.a2, self.sl2, self.lo2
        d.ix0, d.x_lo_edge, d.x_hi_edge = self.ix0, int(self.x_lo_edge), int(self.x_hi_edge)
        d.bn_ixyz = p(self.bn_ixyz, np.int64)
        d.adj_bn = p(self.adj_bn, np.uint16)
        d.bnl_ixyz = p(self.bnl_ixyz, np.int64)
        d.mat_bnl = p(self.mat_bnl, np.int8)
        d.ssaf_bnl = p(self.ssaf_bnl, np.float64)
        d.bna_ixyz = p(self.bna_ixyz, np.int64)
        d.Q_bna = p(self.Q_bna, np.int8)
        d.in_ixyz = p(self.in_ixyz, np.int64)
        d.out_ixyz = p(self.out_ixyz, np.int64)
        d.in_sigs = p(self.in_sigs, np.float64)
        d.Mb = p(self.Mb, np.int8)
        d.mat_beta = p(self.mat_beta, np.float64)
        d.mat_quads = p(self.mat_quads, np.float64)
        self._keep = keep
        return d
