"""ctypes binding of libpffdtd_b200.so (include/pffdtd_b200.h) -- the only way the Python host reaches
the GPU.  There is no CPU fallback: if the library or a CUDA device is missing every call raises.

`Engine` wraps one `pffdtd_engine` (one slab of the grid on one device); `build()` compiles the
library in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from .sim_data import SimData, pffdtd_desc

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libpffdtd_b200.so"
HEADER = HERE.parent / "include" / "pffdtd_b200.h"

OK, EINVAL, ECUDA, ENCCL, ESTATE = 0, -1, -2, -3, -4
PEER_BLOB = 256  # PFFDTD_PEER_BLOB


class pffdtd_energy_desc(C.Structure):
    """include/pffdtd_b200.h: pffdtd_energy_desc"""
    _fields_ = [("struct_size", C.c_int32), ("reserved", C.c_int32), ("h", C.c_double), ("c", C.c_double), ("Ts", C.c_double),
                ("mat_DEF", C.c_void_p)]


class PffdtdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def build(force=False, quiet=True):
    """nvcc -gencode arch=compute_100a,code=sm_100a -> pffdtd_b200/libpffdtd_b200.so"""
    args = ["make", "-C", str(HERE / "csrc")] + (["-B"] if force else [])
    r = subprocess.run(args, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libpffdtd_b200.so failed:\n" + r.stdout + r.stderr)
    if not quiet:
        print(r.stdout)
    return LIB_PATH


_lib = None


def lib():
    """load the shared library (building it first if it is not there) and declare the prototypes"""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        build()
    L = C.CDLL(str(LIB_PATH))
    vp, i64, dp = C.c_void_p, C.c_int64, C.POINTER(C.c_double)
    L.pffdtd_last_error.restype = C.c_char_p
    L.pffdtd_version.restype = C.c_char_p
    L.pffdtd_create.argtypes = [C.POINTER(pffdtd_desc), C.c_int, C.POINTER(vp)]
    L.pffdtd_destroy.argtypes = [vp]
    L.pffdtd_comm_unique_id.argtypes = [vp]
    L.pffdtd_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    L.pffdtd_peer_export.argtypes = [vp, vp]
    L.pffdtd_peer_connect.argtypes = [vp, vp, vp]
    L.pffdtd_set_option.argtypes = [vp, C.c_char_p, i64]
    L.pffdtd_get_stat.argtypes = [vp, C.c_char_p, dp]
    L.pffdtd_reset_stats.argtypes = [vp]
    L.pffdtd_run_steps.argtypes = [vp, i64, i64]
    L.pffdtd_step_host.argtypes = [vp, i64, vp, vp]
    L.pffdtd_sync.argtypes = [vp]
    L.pffdtd_read_outputs.argtypes = [vp, i64, i64, vp]
    L.pffdtd_read_grid.argtypes = [vp, C.c_int, vp]
    L.pffdtd_write_grid.argtypes = [vp, C.c_int, vp]
    L.pffdtd_read_boundary_state.argtypes = [vp, vp, vp]
    L.pffdtd_run_sim.argtypes = [C.POINTER(pffdtd_desc), C.c_int, vp, dp]
    L.pffdtd_energy_enable.argtypes = [vp, C.POINTER(pffdtd_energy_desc)]
    L.pffdtd_read_energy.argtypes = [vp, vp, vp, vp]
    ip = C.POINTER(C.c_int)
    L.pffdtd_multi_create.argtypes = [C.POINTER(pffdtd_desc), C.c_int, ip, C.c_int, C.POINTER(vp)]
    L.pffdtd_multi_destroy.argtypes = [vp]
    L.pffdtd_slab_plan.argtypes = [C.POINTER(pffdtd_desc), C.c_int, C.c_int, C.POINTER(i64), C.POINTER(i64)]
    L.pffdtd_multi_slabs.argtypes = [vp, C.POINTER(i64), C.c_int]
    L.pffdtd_multi_engine.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.pffdtd_multi_run_steps.argtypes = [vp, i64, i64]
    L.pffdtd_multi_sync.argtypes = [vp]
    L.pffdtd_multi_read_outputs.argtypes = [vp, i64, i64, vp]
    L.pffdtd_run_sim_multi.argtypes = [C.POINTER(pffdtd_desc), C.c_int, ip, vp, dp]
    L.pffdtd_air_chunk_plan.argtypes = [i64, C.c_int, C.POINTER(i64), i64]
    L.pffdtd_air_chunk_plan.restype = i64
    _lib = L
    return L


def _check(rc):
    if rc != OK:
        raise PffdtdError(rc, lib().pffdtd_last_error().decode(errors="replace"))


def _prefer_torch_nccl():
    """The library binds NCCL at run time.  When torch is in the process its bundled libnccl must be the one (and
    must be loaded first): a different libnccl.so.2 loaded earlier would break `import torch`."""
    if os.environ.get("PFFDTD_NCCL_LIB"):
        return
    try:
        import torch  # noqa: F401  (loads libtorch_cuda and with it the bundled NCCL)
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            cand = Path(list(spec.submodule_search_locations)[0]) / "lib" / "libnccl.so.2"
            if cand.exists():
                os.environ["PFFDTD_NCCL_LIB"] = str(cand)
    except Exception:  # noqa: BLE001 -- no torch: the system NCCL is used
        pass


def comm_unique_id() -> bytes:
    _prefer_torch_nccl()
    buf = C.create_string_buffer(128)
    _check(lib().pffdtd_comm_unique_id(buf))
    return buf.raw


class Engine:
    """One slab on one device.  `sd` must already be scaled (`SimData.scale_input`) if the caller wants
    the reference binaries' behaviour (fdtd_main.c:41-47)."""

    def __init__(self, sd: SimData, device: int = 0):
        self.sd = sd
        self.L = lib()
        self._desc = sd.desc()
        h = C.c_void_p()
        _check(self.L.pffdtd_create(C.byref(self._desc), int(device), C.byref(h)))
        self.h = h
        self._in = np.zeros(max(sd.Ns, 1), np.float64)
        self._out = np.zeros(max(sd.Nr, 1), np.float64)

    # -- lifetime
    def close(self):
        if getattr(self, "h", None):
            self.L.pffdtd_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration
    def comm_init(self, uid: bytes, rank: int, nranks: int):
        _prefer_torch_nccl()
        buf = C.create_string_buffer(uid, 128)
        _check(self.L.pffdtd_comm_init(self.h, buf, rank, nranks))

    def peer_export(self) -> bytes:
        """this slab's blob for the halo exchange over peer memory (CUDA IPC handles of its grids and flag words)"""
        buf = C.create_string_buffer(PEER_BLOB)
        _check(self.L.pffdtd_peer_export(self.h, buf))
        return buf.raw

    def peer_connect(self, blob_lo, blob_hi):
        """map the neighbours' grids (None at an end of the grid); from now on the step exchanges its halo planes by
        device-to-device copies and device-side flags instead of NCCL"""
        lo = C.create_string_buffer(blob_lo, PEER_BLOB) if blob_lo else None
        hi = C.create_string_buffer(blob_hi, PEER_BLOB) if blob_hi else None
        _check(self.L.pffdtd_peer_connect(self.h, lo, hi))

    def set_option(self, key: str, value: int):
        _check(self.L.pffdtd_set_option(self.h, key.encode(), int(value)))

    def stat(self, key: str) -> float:
        v = C.c_double()
        _check(self.L.pffdtd_get_stat(self.h, key.encode(), C.byref(v)))
        return v.value

    def reset_stats(self):
        _check(self.L.pffdtd_reset_stats(self.h))

    def energy_enable(self):
        """accumulate the reference Python engine's energy balance (sim_fdtd.py:587-620) during the following steps"""
        sd = self.sd
        if not (sd.h > 0 and sd.c > 0):
            raise ValueError("the energy balance needs h and c (sim_consts.h5)")
        d = pffdtd_energy_desc()
        d.struct_size = C.sizeof(pffdtd_energy_desc)
        d.h, d.c, d.Ts = sd.h, sd.c, sd.Ts
        self._def = np.ascontiguousarray(sd.DEF if sd.DEF is not None else np.zeros((max(sd.Nm, 1), 12, 3)), np.float64)
        d.mat_DEF = self._def.ctypes.data
        _check(self.L.pffdtd_energy_enable(self.h, C.byref(d)))

    def read_energy(self):
        """(H_tot[Nt], E_lost[Nt+1], E_in[Nt+1]) summed over this engine's planes"""
        Nt = self.sd.Nt
        H, lost, ein = np.zeros(max(Nt, 1)), np.zeros(Nt + 1), np.zeros(Nt + 1)
        _check(self.L.pffdtd_read_energy(self.h, H.ctypes.data, lost.ctypes.data, ein.ctypes.data))
        return H[:Nt], lost, ein

    # -- stepping
    def run_steps(self, nstart: int, nsteps: int):
        _check(self.L.pffdtd_run_steps(self.h, int(nstart), int(nsteps)))

    def step_host(self, n: int, in_samples=None):
        """one step through host buffers: source samples in, this step's receiver samples out"""
        ip = None
        if in_samples is not None:
            self._in[:self.sd.Ns] = in_samples
            ip = self._in.ctypes.data
        _check(self.L.pffdtd_step_host(self.h, int(n), ip, self._out.ctypes.data))
        return self._out[:self.sd.Nr]

    def sync(self):
        _check(self.L.pffdtd_sync(self.h))

    # -- results
    def read_outputs(self, n0=0, n1=None):
        n1 = self.sd.Nt if n1 is None else n1
        out = np.zeros((self.sd.Nr, n1 - n0), np.float64)
        _check(self.L.pffdtd_read_outputs(self.h, n0, n1, out.ctypes.data))
        return out

    def read_grid(self, which=1):
        g = np.empty(self.sd.Npts, np.float64)
        _check(self.L.pffdtd_read_grid(self.h, which, g.ctypes.data))
        return g.reshape(self.sd.Nx, self.sd.Ny, self.sd.Nz)

    def write_grid(self, which, g):
        g = np.ascontiguousarray(g, np.float64)
        if g.size != self.sd.Npts:
            raise ValueError("grid size mismatch")
        _check(self.L.pffdtd_write_grid(self.h, which, g.ctypes.data))

    def read_boundary_state(self):
        n = self.sd.Nbl * 12
        vh1, gh1 = np.zeros(max(n, 1)), np.zeros(max(n, 1))
        _check(self.L.pffdtd_read_boundary_state(self.h, vh1.ctypes.data, gh1.ctypes.data))
        return vh1[:n].reshape(-1, 12), gh1[:n].reshape(-1, 12)


def run_sim(sd: SimData, device: int = 0):
    """whole-run convenience with the reference's run_sim() shape -> (u_out [Nr,Nt] sorted order, seconds)"""
    d = sd.desc()
    out = np.zeros((sd.Nr, sd.Nt), np.float64)
    t = C.c_double()
    _check(lib().pffdtd_run_sim(C.byref(d), int(device), out.ctypes.data, C.byref(t)))
    return out, t.value


class MultiEngine:
    """All slabs of a grid driven by ONE host thread (pffdtd_multi_*): the reference's single-process multi-GPU model
    (gpu_engine.h:679-691, 994, 1086-1126).  `sd` describes the whole grid with sorted node lists; `nslabs` <= 0 takes one
    slab per visible device; `devices` may name a device several times (slabs sharing a GPU)."""

    def __init__(self, sd: SimData, nslabs: int = 0, devices=None, balance: bool = True):
        self.sd = sd
        self.L = lib()
        self._desc = sd.desc()
        dv = None
        if devices is not None:
            dv = (C.c_int * len(devices))(*[int(d) for d in devices])
            nslabs = len(devices)
        h = C.c_void_p()
        _check(self.L.pffdtd_multi_create(C.byref(self._desc), int(nslabs), dv, int(bool(balance)), C.byref(h)))
        self.h = h
        buf = (C.c_int64 * 64)()
        self.nslabs = self.L.pffdtd_multi_slabs(self.h, buf, 64)
        self.planes = [int(buf[i]) for i in range(min(self.nslabs, 64))]

    def close(self):
        if getattr(self, "h", None):
            self.L.pffdtd_multi_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key: str, value: int):
        """the option on every slab's engine"""
        for r in range(self.nslabs):
            e = C.c_void_p()
            _check(self.L.pffdtd_multi_engine(self.h, r, C.byref(e)))
            _check(self.L.pffdtd_set_option(e, key.encode(), int(value)))

    def stat(self, slab: int, key: str) -> float:
        e, v = C.c_void_p(), C.c_double()
        _check(self.L.pffdtd_multi_engine(self.h, slab, C.byref(e)))
        _check(self.L.pffdtd_get_stat(e, key.encode(), C.byref(v)))
        return v.value

    def run_steps(self, nstart: int, nsteps: int):
        _check(self.L.pffdtd_multi_run_steps(self.h, int(nstart), int(nsteps)))

    def sync(self):
        _check(self.L.pffdtd_multi_sync(self.h))

    def read_outputs(self, n0=0, n1=None):
        n1 = self.sd.Nt if n1 is None else n1
        out = np.zeros((self.sd.Nr, n1 - n0), np.float64)
        _check(self.L.pffdtd_multi_read_outputs(self.h, n0, n1, out.ctypes.data))
        return out


def run_sim_multi(sd: SimData, nslabs: int = 0, devices=None):
    """pffdtd_run_sim_multi: the reference's run_sim() over every visible device -> (u_out [Nr,Nt] sorted order, seconds)"""
    d = sd.desc()
    out = np.zeros((sd.Nr, sd.Nt), np.float64)
    t = C.c_double()
    dv = None
    if devices is not None:
        dv = (C.c_int * len(devices))(*[int(x) for x in devices])
        nslabs = len(devices)
    _check(lib().pffdtd_run_sim_multi(C.byref(d), int(nslabs), dv, out.ctypes.data, C.byref(t)))
    return out, t.value
