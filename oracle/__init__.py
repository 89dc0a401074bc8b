"""TEST INFRASTRUCTURE ONLY -- ctypes access to the CPU oracles.

* ``Oracle``     : oracle/liboracle.so, the C restatement of the reference step (pffdtd_oracle.c).
* ``Reference``  : oracle/_ref/libpffdtd_ref_{f32,f64}.so, the UNMODIFIED reference CPU engine
                   (c_cuda/cpu_engine.h + fdtd_data.h) behind ref_driver.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  Nothing under pffdtd_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_SRC = Path("/root/reference/c_cuda")


def build(quiet=True):
    """compile liboracle.so and, when the reference sources are present, oracle/_ref"""
    r = subprocess.run(["make", "-C", str(HERE)], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if not quiet:
        print(r.stdout)


def _need(path: Path):
    if not path.exists():
        build()
    if not path.exists():
        raise FileNotFoundError(f"{path} missing (and could not be built)")
    return str(path)


class Oracle:
    """The restatement, driven with the product's own `pffdtd_desc` (from SimData.desc())."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(_need(HERE / "liboracle.so"))
            L.oracle_create.restype = C.c_void_p
            L.oracle_create.argtypes = [C.c_void_p]
            L.oracle_destroy.argtypes = [C.c_void_p]
            L.oracle_run_steps.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64]
            for fn in ("oracle_read_grid", "oracle_write_grid"):
                getattr(L, fn).argtypes = [C.c_void_p, C.c_int, C.c_void_p]
            for fn in ("oracle_read_plane", "oracle_write_plane"):
                getattr(L, fn).argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
            L.oracle_read_boundary_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            cls._lib = L
        return cls._lib

    def __init__(self, sd):
        self.sd = sd
        self._desc = sd.desc()
        self.h = self.lib().oracle_create(C.byref(self._desc))
        if not self.h:
            raise RuntimeError("oracle_create failed")
        self.u_out = np.zeros((sd.Nr, max(sd.Nt, 1)), np.float64)

    def run_steps(self, nstart, nsteps):
        if nstart < 0 or nsteps < 0 or nstart + nsteps > self.sd.Nt:  # the C side writes u_out[:, n] unchecked
            raise ValueError(f"steps [{nstart},{nstart + nsteps}) outside [0,{self.sd.Nt})")
        self.lib().oracle_run_steps(self.h, nstart, nsteps, self.u_out.ctypes.data, self.u_out.shape[1])

    def run_all(self):
        self.run_steps(0, self.sd.Nt)
        return self.u_out[:, :self.sd.Nt]

    def read_grid(self, which=1):
        g = np.empty(self.sd.Npts, np.float64)
        self.lib().oracle_read_grid(self.h, which, g.ctypes.data)
        return g.reshape(self.sd.Nx, self.sd.Ny, self.sd.Nz)

    def write_grid(self, which, g):
        g = np.ascontiguousarray(g, np.float64)
        assert g.size == self.sd.Npts
        self.lib().oracle_write_grid(self.h, which, g.ctypes.data)

    def read_plane(self, ix):
        p = np.empty(self.sd.Ny * self.sd.Nz, np.float64)
        self.lib().oracle_read_plane(self.h, ix, p.ctypes.data)
        return p

    def write_plane(self, ix, p):
        p = np.ascontiguousarray(p, np.float64)
        self.lib().oracle_write_plane(self.h, ix, p.ctypes.data)

    def read_boundary_state(self):
        n = self.sd.Nbl * 12
        vh1, gh1 = np.empty(n), np.empty(n)
        self.lib().oracle_read_boundary_state(self.h, vh1.ctypes.data, gh1.ctypes.data)
        return vh1.reshape(-1, 12), gh1.reshape(-1, 12)

    def close(self):
        if self.h:
            self.lib().oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_DT = {np.dtype(np.float64): 0, np.dtype(np.int64): 1, np.dtype(np.int8): 2, np.dtype(np.bool_): 2}


class Reference:
    """The unmodified reference CPU engine (load_sim_data -> scale_input -> run_sim -> rescale_output ->
    write_outputs), fed with the datasets of a data folder through the in-memory H5 shim."""
    _libs = {}

    @staticmethod
    def available(precision=2):
        p = HERE / "_ref" / f"libpffdtd_ref_f{32 if precision == 1 else 64}.so"
        return p.exists() or REF_SRC.exists()

    @classmethod
    def lib(cls, precision, gpu=False):
        """gpu=True: the reference's own CUDA engine (c_cuda/gpu_engine.h), a performance comparator
        (tests/diag/compare_reference_gpu_engine.py); gpu="b200": the reference's loader / scaling / output code around THIS repository's
        engine (integration/b200_engine.h in the place of cpu_engine.h), the drop-in of INTEGRATION.md exercised by the GPU tests"""
        key = (precision, gpu)
        if key not in cls._libs:
            tag = "b200" if gpu == "b200" else ("gpu" if gpu else "")
            L = C.CDLL(_need(HERE / "_ref" / f"libpffdtd_ref{tag}_f{32 if precision == 1 else 64}.so"))
            L.refdrv_put.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
            L.refdrv_load.argtypes = [C.c_char_p]
            L.refdrv_run_sim.restype = C.c_double
            L.refdrv_field.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int)]
            L.refdrv_get_dataset.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_uint64]
            assert L.refdrv_precision() == precision
            cls._libs[key] = L
        return cls._libs[key]

    def __init__(self, precision, files: dict, data_dir, threads=None, gpu=False):
        """files: {'sim_consts': {...}, ...} datasets; data_dir: folder holding the four .h5 files
        (the reference loader stat()s them)."""
        self.L = self.lib(precision, gpu)
        self.precision = precision
        self.real = np.float32 if precision == 1 else np.float64
        self.L.refdrv_clear()
        if threads:
            self.L.refdrv_set_threads(int(threads))
        for stem, ds in files.items():
            for name, arr in ds.items():
                a = np.asarray(arr)
                if a.ndim:
                    a = np.ascontiguousarray(a)
                if a.dtype == np.bool_:
                    a = a.astype(np.int8)
                if a.dtype not in _DT:
                    a = a.astype(np.float64 if a.dtype.kind == "f" else np.int64)
                dims = (C.c_int64 * max(a.ndim, 1))(*a.shape)
                rc = self.L.refdrv_put(f"{stem}.h5".encode(), name.encode(), _DT[a.dtype], a.ndim, dims, a.ctypes.data)
                assert rc == 0
        rc = self.L.refdrv_load(str(data_dir).encode())
        if rc != 0:
            raise RuntimeError(f"refdrv_load failed ({rc})")

    def field(self, name):
        ptr, cnt, es = C.c_void_p(), C.c_int64(), C.c_int()
        if self.L.refdrv_field(name.encode(), C.byref(ptr), C.byref(cnt), C.byref(es)) != 0:
            raise KeyError(name)
        real_fields = {"ssaf_bnl", "mat_beta", "mat_quads", "sl2", "lo2", "a1", "a2"}
        if name in real_fields:
            dt = self.real
        elif name in ("in_sigs", "u_out", "l", "l2", "infac"):
            dt = np.float64
        elif name == "adj_bn":
            dt = np.uint16
        elif name == "bn_mask":
            dt = np.uint8
        else:
            dt = {1: np.int8, 8: np.int64}[es.value]
        n = cnt.value
        if n == 0:
            return np.zeros(0, dt)
        buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr.value)
        a = np.frombuffer(buf, dtype=dt, count=n).copy()
        return a[0] if name in ("Ns", "Nr", "Nt", "Npts", "Nx", "Ny", "Nz", "Nb", "Nbl", "Nba", "l", "l2", "fcc_flag",
                                "NN", "Nm", "infac", "sl2", "lo2", "a2", "a1") else a

    def run(self):
        """scale_input + run_sim + rescale_output + write_outputs; returns (u_out as written to the
        file [Nr,Nt], engine-reported seconds)"""
        self.L.refdrv_scale_input()
        t = self.L.refdrv_run_sim()
        self.L.refdrv_rescale_output()
        self.L.refdrv_write_outputs()
        Nr, Nt = int(self.field("Nr")), int(self.field("Nt"))
        out = np.empty((Nr, Nt), np.float64)
        rc = self.L.refdrv_get_dataset(b"sim_outs.h5", b"u_out", out.ctypes.data, out.nbytes)
        assert rc == 0
        return out, t

    def run_sim_only(self):
        """just run_sim (for timing); inputs must have been scaled by a previous scale_input or be fine as is"""
        return self.L.refdrv_run_sim()
