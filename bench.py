#!/usr/bin/env python
"""bench.py -- Gvoxel-updates/s of the simulation step on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c2|c4|...] [--impl reference] [--no-also]

A "step" is one FDTD time step over the whole grid (air update + boundary passes + source/receiver io).
EVERY N runs the grid the metric is quoted on, BASELINE.json configs[4]: the 2048x2048x1024 fp32 7-point shoebox with
lossy walls (34 GB of state: it fits one B200).  N>1 (under torchrun, one rank per GPU) splits it into x-slabs with the
halo planes exchanged by NCCL -- fixed total work, "scaling": "strong" -- so the 1/2/4/8 values are one curve.
Prints ONE JSON line on rank 0.

* value     : Npts*K / device time (CUDA events on the engine's stream, max over ranks), all state resident in HBM
* e2e       : the same K steps through pffdtd_step_host: per step the source samples go host->device and the
              receiver samples come back device->host, wall clock
* roofline  : the air kernel: 12.125 B (fp32) / 24.125 B (fp64) per node and launch / mean launch time from CUDA
              events around every air launch inside the timed region, against MEASURED_PEAKS.json hbm_gbs
* parity    : (outside the timed region) SHA-256 of the receiver traces of the run's first PARITY_STEPS steps -- a source sits on
              every interface of the 8-way split and 280 receivers cross them -- against the committed hash of the one-GPU run
              (tests/golden/c5_parity.json), plus a reduced grid with the same N-way split, same code path (captured step graphs,
              NCCL exchange), compared bit for bit with the CPU engine (oracle/: checker only)
* also      : N=1 only: the other single-GPU configurations (c2 = configs[1], c3s = configs[2] stand-in, c4 = configs[3], the real
              rooms when data_large/ is there), each with its own value / e2e / roofline
* cpu_baseline : the UNMODIFIED reference CPU engine (oracle/_ref, c_cuda/cpu_engine.h) on this box's cores, on a
              bounded sample of the same workload
`--impl reference` times only that CPU engine (rank 0) and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: grid, precision, stencil, walls
    "c2": dict(N=(512, 512, 256), precision=1, fcc=False, nmat=1, mb=11, rigid=False,
               desc="BASELINE configs[1]: shoebox 512x512x256, 7-pt Cartesian, fp32, lossy walls (1 material, 11 branches)"),
    "c2_rigid": dict(N=(512, 512, 256), precision=1, fcc=False, nmat=0, mb=0, rigid=True,
                     desc="shoebox 512x512x256, 7-pt Cartesian, fp32, rigid walls"),
    "c3s": dict(N=(1024, 512, 512), precision=1, fcc=True, nmat=1, mb=11, rigid=False,
                desc="BASELINE configs[2] stand-in: shoebox 1024x512x512 unfolded (1024x257x512 stored), 13-pt FCC folded, fp32, lossy walls"),
    "c4": dict(N=(1024, 1024, 1024), precision=2, fcc=False, nmat=0, mb=0, rigid=True,
               desc="BASELINE configs[3]: rigid shoebox 1024^3, 7-pt Cartesian, fp64"),
    "c5": dict(N=(2048, 2048, 1024), precision=1, fcc=False, nmat=1, mb=11, rigid=False, probes=True,
               desc="BASELINE configs[4]: shoebox 2048x2048x1024, 7-pt Cartesian, fp32, lossy walls (1 material, 11 branches); "
                    "N>1: x-slabs + NCCL halo exchange"),
    "small": dict(N=(128, 96, 64), precision=1, fcc=False, nmat=1, mb=11, rigid=False, desc="smoke-sized shoebox"),
    "small_probes": dict(N=(128, 96, 64), precision=1, fcc=False, nmat=1, mb=11, rigid=False, probes=True, desc="smoke-sized shoebox with the c5 probes"),
    # the real rooms at full size, voxelised by the reference tool chain (tools/make_large_models.py -> data_large/, not committed)
    "ctk_real": dict(folder="data_large/ctk_cart_gpu", precision=1, fcc=False,
                     desc="BASELINE configs[1] with the real model: CTK church, h = 0.041 m, 7-pt Cartesian, fp32, 8 materials x 11 branches"),
    "mv_real": dict(folder="data_large/mv_fcc_gpu", precision=1, fcc=True,
                    desc="BASELINE configs[2] with the real model: Musikverein, h = 0.06 m, 13-pt FCC folded, fp32, 5 materials x 11 branches"),
    "mv_big": dict(folder="data_large/mv_fcc_gpu_big", precision=1, fcc=True,
                   desc="BASELINE configs[2] with the real model at a finer grid: Musikverein, 13-pt FCC folded, fp32, 5 materials x 11 branches"),
    "mv_full": dict(folder="data_large/mv_fcc_gpu_full", precision=1, fcc=True,
                    desc="BASELINE configs[2] at the size of the reference's own script (python/test_script_MV_fcc_gpu.py: fmax 2500 Hz, PPW 7.7, "
                         "h = 0.0178 m): Musikverein, 13-pt FCC folded, fp32, 5 materials x 11 branches"),
}
BYTES_PER_NODE = {1: 12.125, 2: 24.125}  # u1 read + u0 read + u0 write + 1 mask bit
# CPU-arm sample grids for workloads whose full grid would take the CPU engine minutes per step
CPU_SAMPLE_GRID = {
    "c5": ((258, 2048, 1024), "one of eight x-slabs of the c5 grid (256 planes + 2 halo planes, 258x2048x1024 = 1/8 of the nodes; "
                              "the CPU engine's Gvox/s does not depend on the number of planes)"),
    "c4": ((130, 1024, 1024), "one eighth of the c4 grid (130x1024x1024)"),
}
# extra single-GPU lines of the N=1 run: workload -> timed steps
ALSO = (("c2", 200), ("c3s", 100), ("c4", 20), ("ctk_real", 200), ("mv_real", 150), ("mv_big", 40), ("mv_full", 20))
PARITY_STEPS = 24
PARITY_FILE = ROOT / "tests" / "golden" / "c5_parity.json"


def add_probes(files, Nx, Ny, Nz, fcc=False):
    """a source on every interface of the 8-way equal split and a line of 40 receivers across each (the set-up of
    tools/mgpu_equal.py): within a few steps every halo exchange of a 2/4/8-slab run carries non-zero data in both
    directions, so the receiver traces prove the exchange, whatever the split"""
    cm = files["comms_out"]
    sig = cm["in_sigs"][0] / np.abs(cm["in_sigs"][0]).max()
    faces = [Nx * k // 8 for k in range(1, 8)]
    node = lambda ix, iy, iz: (ix * Ny + iy) * Nz + iz
    half = min(20, Nx // 16)
    in_ixyz = np.array([node(f - (k & 1), Ny // 2 + (5 * k) % (Ny // 4), Nz // 2 - (3 * k) % (Nz // 4)) for k, f in enumerate(faces)], np.int64)
    out_ixyz = np.array(sorted(node(f + d, Ny // 2 + (5 * k) % (Ny // 4) + 2, Nz // 2 - (3 * k) % (Nz // 4) + 1)
                               for k, f in enumerate(faces) for d in range(-half, half)), np.int64)
    bn = files["vox_out"]["bn_ixyz"]
    assert not np.intersect1d(in_ixyz, bn).size
    cm.update(in_ixyz=in_ixyz, in_sigs=np.stack([sig * (1.0 + 0.1 * k) for k in range(len(faces))]), Ns=np.int64(in_ixyz.size),
              out_ixyz=out_ixyz, out_reorder=np.arange(out_ixyz.size, dtype=np.int64), Nr=np.int64(out_ixyz.size))
    return files


def build_problem(wl, Nt, x_range=None):
    from pffdtd_b200 import folder_prep, shoebox
    w = WORKLOADS[wl]
    if "folder" in w:  # a sim_setup folder: keep everything, stretch / cut the source signals to Nt steps (zeros after the end)
        d = ROOT / w["folder"]
        if not (d / "vox_out.h5").exists():
            raise SystemExit(f"{d} missing: python tools/make_large_models.py (needs the reference tool chain)")
        files = folder_prep.load_folder(d)
        cm = files["comms_out"]
        sig = np.asarray(cm["in_sigs"], np.float64)
        out = np.zeros((sig.shape[0], Nt))
        out[:, :min(Nt, sig.shape[1])] = sig[:, :Nt]
        cm["in_sigs"], cm["Nt"] = out, np.int64(Nt)
        return files
    Nx, Ny, Nz = w["N"]
    files = shoebox.make_shoebox(Nx, Ny, Nz, Nt, fcc=w["fcc"], nmat=w["nmat"], mb=max(w["mb"], 1), rigid=w["rigid"], diff=True,
                                 x_range=x_range)
    if w.get("probes"):
        files = add_probes(files, Nx, Ny, Nz)
    if w["fcc"]:
        files = folder_prep.gpu_folder(files)
    return files


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons of one GPU while the timed region runs (pynvml, 5 ms period).  The thread is started
    before the warm-up and `begin()` blocks until NVML answers, so that even a 100 ms timed region is sampled."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz, self.err = index, False, [], set(), None, None
        self.ready, self.recording = threading.Event(), False

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop_flag:
                mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                try:
                    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.ready.set()
                if self.recording:
                    self.samples.append(mhz)
                    for bit, name in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
                time.sleep(0.005)
        except Exception as ex:  # noqa: BLE001
            self.err = repr(ex)
            self.ready.set()

    def begin(self):
        self.ready.wait(timeout=10)
        self.recording = True

    def result(self):
        self.recording = False
        self.stop_flag = True
        self.join(timeout=2)
        d = {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
             "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.err:
            d["error"] = self.err
        return d


def nvml_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            pass
    return local_rank


def run_reference_cpu(wl, steps, warmup, budget_s=120.0, threads=None):
    """the unmodified reference CPU engine on `wl`: -> (Gvox/s, cores, steps timed, seconds, kind, what was sampled)"""
    import tempfile
    from oracle import Reference
    w = WORKLOADS[wl]
    cores = threads or os.cpu_count() or 1
    sample_note = f"the full {wl} grid"
    if wl in CPU_SAMPLE_GRID:  # the grid is too large for a CPU run of a few minutes: time one slab of it
        WORKLOADS[wl + "_cpu_sample"] = dict(w, N=CPU_SAMPLE_GRID[wl][0], probes=False)
        sample_note = CPU_SAMPLE_GRID[wl][1]
        wl = wl + "_cpu_sample"
        w = WORKLOADS[wl]
    if "folder" in w and "N" not in w:
        from pffdtd_b200 import h5lite
        v = h5lite.File(ROOT / w["folder"] / "vox_out.h5")
        w = WORKLOADS[wl] = dict(w, N=tuple(int(v[k][()]) for k in ("Nx", "Ny", "Nz")))
    Npts = int(np.prod(w["N"]))
    if w["fcc"] and "folder" not in w:
        Npts = w["N"][0] * (w["N"][1] // 2 + 1) * w["N"][2]
    # bound the sample: size it for ~0.8 Gvox/s (the pool's 16-core hosts measure 0.8-3.2), then report what was actually run
    per_step = Npts / 0.8e9
    k = int(max(2, min(steps, budget_s / per_step)))
    wu = int(max(1, min(warmup, max(1, k // 4))))
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    tmp = tempfile.mkdtemp(prefix="pffdtd_ref_")
    for fn in ("sim_consts.h5", "vox_out.h5", "comms_out.h5", "sim_mats.h5"):
        (Path(tmp) / fn).touch()
    try:
        sys.stdout.flush()
        os.dup2(devnull, 1)  # print_progress writes every step
        t_run = None
        for nt in (wu, k):
            files = build_problem(wl, nt)
            ref = Reference(w["precision"], files, tmp, threads=cores)
            ref.L.refdrv_scale_input()
            t_run = ref.run_sim_only()
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)
    return Npts * k / t_run / 1e9, cores, k, t_run, "reference", sample_note


def workload_grid(wl):
    """(Nx, Ny stored, Nz) of a workload; folder workloads read their vox_out.h5"""
    w = WORKLOADS[wl]
    if "folder" in w:  # stored grid of the folder (already folded for FCC)
        from pffdtd_b200 import h5lite
        v = h5lite.File(ROOT / w["folder"] / "vox_out.h5")
        Nx, Ny, Nz = (int(v[k][()]) for k in ("Nx", "Ny", "Nz"))
        WORKLOADS[wl] = dict(w, N=(Nx, Ny, Nz))
        return Nx, Ny, Nz
    Nx, Ny, Nz = w["N"]
    return Nx, (Ny // 2 + 1 if w["fcc"] else Ny), Nz


class Ctx:
    """rank / device / process group of this run"""

    def __init__(self, args):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.args = args

    def barrier(self, eng=None):
        import torch
        if eng is not None:
            eng.sync()
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def allmax(self, v):
        if self.dist is None:
            return float(v)
        import torch
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(self, v):
        if self.dist is None:
            return float(v)
        import torch
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather(self, obj):
        if self.dist is None:
            return [obj]
        parts = [None] * self.world
        self.dist.all_gather_object(parts, obj)
        return parts


def slab_plan(ctx, wl, Nx, Ny, Nz):
    """(starts, sizes) of the x-slabs: cost-weighted (air nodes + weighted boundary / lossy / shell nodes per plane) so that the
    ranks holding the x walls get fewer planes; --equal-slabs gives the reference's Nx/n split (gpu_engine.h:532-543)"""
    from pffdtd_b200 import shoebox
    from pffdtd_b200.sim_data import SimData
    if ctx.world == 1:
        return [0], [Nx]
    if ctx.args.equal_slabs or "folder" in WORKLOADS[wl] or WORKLOADS[wl]["fcc"]:
        return SimData.slab_planes(Nx, ctx.world)
    w = WORKLOADS[wl]
    return SimData.slab_planes(Nx, ctx.world, cost=shoebox.plane_costs(Nx, Ny, Nz, rigid=w["rigid"], mb=w["mb"]))


def make_engine(ctx, wl, Nt, config):
    """problem + engine of this rank for workload `wl` -> (sd_full, sd, eng, prep seconds)"""
    from pffdtd_b200 import shoebox
    from pffdtd_b200.engine import Engine, comm_unique_id
    args, w = ctx.args, WORKLOADS[wl]
    Nx, Ny_st, Nz = workload_grid(wl)
    t_prep = time.perf_counter()
    # a rank only generates the boundary / shell nodes of its own slab (+ halo planes): the node lists of the
    # 2048x2048x1024 grid have 5e7 entries
    xr = planes = None
    if ctx.world > 1:
        planes = slab_plan(ctx, wl, Nx, w["N"][1], Nz)
        config["slab_planes"] = list(planes[1])
        if not w["fcc"] and "folder" not in w:
            xr = (max(0, planes[0][ctx.rank] - 1), min(Nx, planes[0][ctx.rank] + planes[1][ctx.rank] + 1))
    files = build_problem(wl, Nt, x_range=xr)
    sd_full = shoebox.sim_data_from_files(files, w["precision"], abc_x_range=xr).scale_input()
    del files
    if ctx.world > 1:
        sd_full = sd_full.sorted()
    sd = sd_full.slab(ctx.rank, ctx.world, planes=planes) if ctx.world > 1 else sd_full
    eng = Engine(sd, ctx.local)
    if ctx.world > 1:
        box = [comm_unique_id() if ctx.rank == 0 else None]
        ctx.dist.broadcast_object_list(box, src=0)
        eng.comm_init(box[0], ctx.rank, ctx.world)
        eng.set_option("overlap", args.overlap)
        from pffdtd_b200 import parallel
        p2p = (not args.no_p2p) and parallel.connect_peers(eng, ctx.rank, ctx.world)
        how = "overlapped with the interior update on a second stream" if args.overlap else "after the step (not overlapped)"
        config["halo_exchange"] = (
            "peer memory: one device-to-device copy of a plane per neighbour per step into a CUDA-IPC mapping of its grid + a copied flag "
            "word, " + how + "; step pairs replay as CUDA graphs that include the exchange" if p2p else
            "ncclSend/ncclRecv of one plane per neighbour per step, " + how + "; the work before / after the exchange replays from CUDA graphs")
    eng.set_option("air_kernel", args.air_kernel)
    if args.xc:
        eng.set_option("air_xc", args.xc)
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
        config.setdefault("options", {})[k] = int(v)
    if args.air_cfg >= 0:
        eng.set_option("air_cfg", args.air_cfg)
        config["air_cfg"] = args.air_cfg
    t_prep = time.perf_counter() - t_prep
    if args.air_kernel == 1:
        try:  # informational only: never let a missing counter cost the bench line
            config["air_tile"] = {"cfg": int(eng.stat("air_cfg")), "lanes_z": int(eng.stat("air_lanes_z")), "fused_step": bool(eng.stat("fused"))}
        except Exception as ex:  # noqa: BLE001
            config["air_tile"] = {"error": repr(ex)}
    return sd_full, sd, eng, t_prep


def measure(ctx, wl, K, W, with_e2e=True, with_parity=False):
    """one workload on this run's GPUs -> the pieces of a bench line"""
    args, w = ctx.args, WORKLOADS[wl]
    Nx, Ny_st, Nz = workload_grid(wl)
    Npts = Nx * Ny_st * Nz
    config = {"workload": wl, "description": w["desc"], "grid": [Nx, Ny_st, Nz], "stencil": "13pt_fcc_folded" if w["fcc"] else "7pt_cartesian",
              "l2": "state (2 grids, %.0f MB) is larger than the 126 MB L2" % (2 * Npts * (4 if w["precision"] == 1 else 8) / 1e6)}
    W = W + (W & 1)  # an even warm-up leaves the grid roles as they started; both step graphs are captured during it anyway
    n_par = PARITY_STEPS if with_parity else 0
    Nt = max(W + K, n_par) + K + 2  # warm-up + timed steps (+ the rest of the parity window), then the roofline pass
    sd_full, sd, eng, t_prep = make_engine(ctx, wl, Nt, config)

    # ---- device-resident run
    sampler = ClockSampler(nvml_index(ctx.local))
    sampler.start()
    eng.run_steps(0, W)
    ctx.barrier(eng)
    sampler.begin()
    eng.reset_stats()
    eng.stat("timer_start")
    eng.run_steps(W, K)
    ms = eng.stat("timer_stop_ms")
    ctx.barrier(eng)
    clocks = sampler.result()
    launches = eng.stat("launches")
    n_done = W + K
    parity = None
    if with_parity:
        if n_done < n_par:
            eng.run_steps(n_done, n_par - n_done)
            n_done = n_par
        u = np.concatenate(ctx.gather(eng.read_outputs(0, n_par)), axis=0)  # rank order == sorted receiver order
        parity = {"steps": n_par, "receivers": int(u.shape[0]), "nonsilent_receivers": int((np.abs(u).max(axis=1) > 0).sum()),
                  "traces_sha256": hashlib.sha256(np.ascontiguousarray(u, np.float64).tobytes()).hexdigest()}
        try:
            exp = json.loads(PARITY_FILE.read_text()).get(wl)
        except Exception:  # noqa: BLE001
            exp = None
        if exp and exp.get("steps") == n_par:
            parity["expected_sha256"], parity["expected_from"] = exp["sha256"], exp.get("from")
            parity["equal_to_one_gpu_run"] = exp["sha256"] == parity["traces_sha256"]
        else:
            parity["equal_to_one_gpu_run"] = None
    # roofline pass: the next K steps with CUDA events around every air launch (events cannot sit inside the
    # replayed CUDA graph of the pass above, so this pass launches kernel by kernel; same kernels, same data)
    eng.set_option("profile_air", 1)
    eng.reset_stats()
    eng.stat("timer_start")
    eng.run_steps(n_done, K)
    ms_prof = eng.stat("timer_stop_ms")
    ctx.barrier(eng)
    eng.set_option("profile_air", 0)
    air_ms = eng.stat("air_ms")
    air_n = eng.stat("air_launches_timed")
    ms = ctx.allmax(ms)
    launches = ctx.allsum(launches)
    value = Npts * K / (ms * 1e-3) / 1e9

    # ---- end to end through host buffers: a fresh engine state is not needed, the work per step is identical
    e2e = None
    if with_e2e and not args.no_e2e:
        ins = np.ascontiguousarray(sd.in_sigs.T)  # [Nt][Ns] host samples
        for n in range(4):  # plain steps, then both host-step graphs are captured and replayed once
            eng.step_host(n, ins[n])
        ctx.barrier(eng)
        t0 = time.perf_counter()
        for n in range(W, W + K):
            eng.step_host(n, ins[n])
        eng.sync()
        dt = ctx.allmax(time.perf_counter() - t0)
        rs = 4 if w["precision"] == 1 else 8
        e2e = {"value": Npts * K / dt / 1e9, "unit": "Gvox/s", "h2d_bytes_per_step": int(sd_full.Ns * rs),
               "d2h_bytes_per_step": int(sd_full.Nr * rs),
               "note": "pffdtd_step_host: per step source samples H2D from pinned memory, receiver samples D2H, host sync"}

    # ---- roofline of the air kernel (this rank's launches)
    peak, peak_src = peaks()
    nodes_per_launch = (sd.Nx - 2) * sd.Ny * sd.Nz  # planes 1..Nx-2 of this rank's slab
    n_air_per_step = max(1.0, air_n / K)
    air_ms_per_step = air_ms / K
    achieved = BYTES_PER_NODE[w["precision"]] * nodes_per_launch / (air_ms_per_step * 1e-3) / 1e9 if air_ms > 0 else None
    traffic = None
    try:  # DRAM bytes per launch of this kernel from the committed ncu --set full capture of the same workload
        tr = json.loads((ROOT / "profiles" / "traffic.json").read_text()).get(wl)
        if tr and args.air_kernel == 1 and ctx.world == 1:
            traffic = tr["traffic_bytes"]
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "hbm", "kernel": ("k_air_tma_cart<FCC>" if w["fcc"] else "k_air_tma_cart") if args.air_kernel == 1 else "k_air_generic",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "algorithmic_bytes_per_launch": BYTES_PER_NODE[w["precision"]] * nodes_per_launch,
                "peak_source": peak_src, "bytes_per_node": BYTES_PER_NODE[w["precision"]], "nodes_per_launch": int(nodes_per_launch),
                "air_ms_per_step": air_ms_per_step, "air_launches_per_step": n_air_per_step, "air_share_of_step": air_ms_per_step / (ms_prof / K),
                "whole_step_frac": BYTES_PER_NODE[w["precision"]] * Npts / ctx.world / (ms / K * 1e-3) / 1e9 / peak,
                "note": "air launches timed with CUDA events in a second pass of K steps launched kernel by kernel (%.4f ms/step); "
                        "the `value` pass replays the step as a CUDA graph; whole_step_frac = the whole step's Gvox/s x bytes_per_node "
                        "against the same peak" % (ms_prof / K)}
    # the step's REQUIRED traffic also holds the lossy boundary nodes' branch state (k_fd: 16 B per branch + 30 B per node in fp32,
    # DESIGN.md 4), which bytes_per_node leaves out; with it the whole step reads as a fraction of the same peak
    rs4 = 1 if w["precision"] == 1 else 2
    fd_bytes = float(np.sum(16 * sd.Mb.astype(np.int64)[sd.mat_bnl.astype(np.int64)] + 30)) * rs4 if sd.Nbl else 0.0
    fd_bytes_all = sum(ctx.gather(fd_bytes) or [fd_bytes]) if ctx.world > 1 else fd_bytes
    roofline["boundary_state_bytes_per_step"] = fd_bytes_all
    roofline["whole_step_frac_with_boundary_state"] = (BYTES_PER_NODE[w["precision"]] * Npts + fd_bytes_all) / ctx.world / (ms / K * 1e-3) / 1e9 / peak
    if ctx.world > 1:
        # non-air time of every rank (the end ranks carry the x walls): what the cost-weighted split balances
        per_rank = ctx.gather({"rank": ctx.rank, "planes": int(sd.Nx), "air_ms": air_ms_per_step, "step_ms_kernel_by_kernel": ms_prof / K,
                               "non_air_ms": ms_prof / K - air_ms_per_step, "Nb": int(sd.Nb), "Nbl": int(sd.Nbl), "Nba": int(sd.Nba)})
        config["per_rank"] = per_rank
    eng.close()
    out = {"value": value, "ms_per_step": ms / K, "steps": K, "warmup": W, "dtype": "f32" if w["precision"] == 1 else "f64",
           "config": config, "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches), "host_prep_s": round(t_prep, 2)}
    if e2e:
        out["e2e"] = e2e
    if parity:
        out["parity"] = parity
    return out


def reduced_grid_parity(ctx):
    """CHECKER, outside every timed region: a reduced lossy shoebox with the probes, split N ways exactly like the c5 run
    (cost-weighted planes, NCCL exchange, captured step graphs), against the CPU engine on the whole reduced grid, bit for bit"""
    from pffdtd_b200 import shoebox
    from pffdtd_b200.engine import Engine, comm_unique_id
    from pffdtd_b200.sim_data import SimData
    Nx, Ny, Nz, Nt = 16 * max(ctx.world, 2) + 10, 40, 72, 48
    files = add_probes(shoebox.make_shoebox(Nx, Ny, Nz, Nt, nmat=2, mb=11, diff=True), Nx, Ny, Nz)
    full = shoebox.sim_data_from_files(files, 1).scale_input().sorted()
    planes = SimData.slab_planes(Nx, ctx.world, cost=None if ctx.args.equal_slabs else full.plane_costs()) if ctx.world > 1 else None
    sd = full.slab(ctx.rank, ctx.world, planes=planes) if ctx.world > 1 else full
    eng = Engine(sd, ctx.local)
    if ctx.world > 1:
        box = [comm_unique_id() if ctx.rank == 0 else None]
        ctx.dist.broadcast_object_list(box, src=0)
        eng.comm_init(box[0], ctx.rank, ctx.world)
        from pffdtd_b200 import parallel
        res_p2p = (not ctx.args.no_p2p) and parallel.connect_peers(eng, ctx.rank, ctx.world)
    eng.run_steps(0, Nt)
    graphed = Nt - 2
    u = np.concatenate(ctx.gather(eng.read_outputs(0, Nt)), axis=0)
    eng.close()
    res = {"grid": [Nx, Ny, Nz], "steps": Nt, "slab_planes": list(planes[1]) if planes else [Nx], "steps_replayed_from_graphs": graphed}
    if ctx.world > 1:
        res["halo_exchange"] = "peer memory" if res_p2p else "nccl"
    if ctx.rank == 0:
        try:
            from oracle import Oracle, Reference  # checker only
            if (ROOT / "oracle" / "_ref" / "libpffdtd_ref_f32.so").exists():
                import tempfile
                tmp = tempfile.mkdtemp(prefix="pffdtd_par_")
                shoebox.write_folder(files, tmp)
                devnull, saved = os.open(os.devnull, os.O_WRONLY), os.dup(1)
                try:
                    sys.stdout.flush()
                    os.dup2(devnull, 1)  # the reference prints its progress bar
                    want, _ = Reference(1, files, tmp).run()
                finally:
                    os.dup2(saved, 1)
                    os.close(devnull)
                    os.close(saved)
                got = full.reorder_output(full.rescale_output(u))  # what write_outputs puts in sim_outs.h5 (fdtd_data.h:912-980)
                res["checker"] = "oracle/_ref/libpffdtd_ref_f32.so: the UNMODIFIED reference CPU engine (load_sim_data .. write_outputs)"
            else:
                want, got = Oracle(full).run_all(), u
                res["checker"] = "oracle/liboracle.so (C restatement of cpu_engine.h, pinned to the unmodified engine by tests/test_oracle.py)"
            res["bit_exact"] = bool(np.array_equal(got, want))
            res["peak"] = float(np.abs(want).max())
            if not res["bit_exact"]:
                res["max_abs_diff"] = float(np.abs(got - want).max())
        except Exception as ex:  # noqa: BLE001
            res["bit_exact"] = None
            res["error"] = repr(ex)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--air-kernel", type=int, default=1, choices=(0, 1))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="N=1: skip the extra single-GPU workloads")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--also", default=None, help="comma-separated workloads for the `also` key (default: all that are available)")
    ap.add_argument("--no-p2p", action="store_true", help="N>1: keep the NCCL halo exchange instead of the peer-memory one")
    ap.add_argument("--equal-slabs", action="store_true", help="N>1: the reference's equal-plane split instead of the cost-weighted one")
    ap.add_argument("--xc", type=int, default=0)
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE", help="engine option (pffdtd_set_option), repeatable (tuning / A-B runs)")
    ap.add_argument("--air-cfg", type=int, default=-1, help="tile configuration of the TMA air kernel (tuning)")
    ap.add_argument("--overlap", type=int, default=1, choices=(0, 1), help="N>1: 0 = halo exchange after the whole step (diagnostic)")
    args = ap.parse_args()

    ctx = Ctx(args)
    N = args.gpus
    if ctx.world != N and ctx.world > 1:
        raise SystemExit(f"--gpus {N} but WORLD_SIZE={ctx.world}")
    wl = args.workload or "c5"
    w = WORKLOADS[wl]
    K, W = args.steps, max(args.warmup, 3)

    if args.impl == "reference":
        if ctx.rank != 0:
            return
        Nx, Ny_st, Nz = workload_grid(wl)
        Npts = Nx * Ny_st * Nz
        config = {"workload": wl, "description": w["desc"], "grid": [Nx, Ny_st, Nz], "stencil": "13pt_fcc_folded" if w["fcc"] else "7pt_cartesian",
                  "l2": "state (2 grids, %.0f MB) is larger than the 126 MB L2" % (2 * Npts * (4 if w["precision"] == 1 else 8) / 1e6)}
        v, cores, k, t, kind, note = run_reference_cpu(wl, K, W)
        line = {"impl": "reference", "metric": "Gvoxel-updates/s", "value": v, "unit": "Gvox/s", "n_gpus": N, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * t / k, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32" if w["precision"] == 1 else "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "Gvox/s", "cores": cores, "kind": kind,
                                 "sample": f"{k} time steps (of the {K} requested) on {note}, unmodified c_cuda/cpu_engine.h run_sim, OpenMP {cores} threads"},
                "e2e": {"value": v, "unit": "Gvox/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(ctx.local)
    if ctx.world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", ctx.local))
        ctx.dist = dist

    m = measure(ctx, wl, K, W, with_parity=bool(w.get("probes")) and not args.no_parity)
    line = {"metric": "Gvoxel-updates/s", "value": m["value"], "unit": "Gvox/s", "n_gpus": N, "steps": K, "warmup": m["warmup"],
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": m["dtype"],
            "data": "synthetic", "config": m["config"], "roofline": m["roofline"], "clocks": m["clocks"], "gpu_launches": m["gpu_launches"],
            "host_prep_s": m["host_prep_s"]}
    if "e2e" in m:
        line["e2e"] = m["e2e"]
    if "parity" in m:
        line["parity"] = m["parity"]
        try:
            line["parity"]["reduced_grid_vs_cpu_engine"] = reduced_grid_parity(ctx)
        except Exception as ex:  # noqa: BLE001
            line["parity"]["reduced_grid_vs_cpu_engine"] = {"bit_exact": None, "error": repr(ex)}

    # ---- the other single-GPU configurations (N=1): same measurement, shorter
    if N == 1 and not args.no_also and args.workload is None:
        names = [a for a in (args.also.split(",") if args.also else [n for n, _ in ALSO])]
        also = {}
        for name, k_also in ALSO:
            if name not in names:
                continue
            ww = WORKLOADS[name]
            if "folder" in ww and not (ROOT / ww["folder"] / "vox_out.h5").exists():
                continue
            try:
                r = measure(ctx, name, k_also, 6)
                also[name] = {"value": r["value"], "unit": "Gvox/s", "ms_per_step": r["ms_per_step"], "steps": r["steps"], "dtype": r["dtype"],
                              "config": r["config"], "roofline": r["roofline"], "e2e": r.get("e2e"), "gpu_launches": r["gpu_launches"]}
            except BaseException as ex:  # noqa: BLE001 -- an extra line must never cost the headline
                also[name] = {"error": repr(ex)}
        line["also"] = also

    # ---- the reference's CPU engine on this box's cores (rank 0, N=1 only)
    if ctx.rank == 0 and N == 1 and not args.no_cpu:
        try:
            v, cores, k, t, kind, note = run_reference_cpu(wl, 300, 3, budget_s=25.0)
            line["cpu_baseline"] = {"value": v, "unit": "Gvox/s", "cores": cores, "kind": kind,
                                    "sample": f"{k} time steps on {note}, unmodified c_cuda/cpu_engine.h run_sim, OpenMP {cores} threads, {t:.1f} s"}
        except Exception as ex:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "Gvox/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex!r}"}
    if ctx.rank == 0:
        print(json.dumps(line), flush=True)
    if ctx.dist is not None:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
