#!/usr/bin/env python
"""bench.py -- Gvoxel-updates/s of the simulation step on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c4|c5|...] [--impl reference]

A "step" is one FDTD time step over the whole grid (air update + boundary passes + source/receiver io).
N=1 runs BASELINE.json configs[1] (synthetic shoebox 512x512x256, 7-point Cartesian, fp32, one lossy wall
material with 11 RLC branches); N>1 (under torchrun, one rank per GPU) runs configs[4] (2048x2048x1024
fp32) split into x-slabs with the halo planes exchanged by NCCL -- fixed total work, "scaling": "strong".
Prints ONE JSON line on rank 0.

* value     : Npts*K / device time (CUDA events on the engine's stream, max over ranks), all state resident in HBM
* e2e       : the same K steps through pffdtd_step_host: per step the source samples go host->device and the
              receiver samples come back device->host, wall clock
* roofline  : the air kernel: 12.125 B (fp32) / 24.125 B (fp64) per node and launch / mean launch time from CUDA
              events around every air launch inside the timed region, against MEASURED_PEAKS.json hbm_gbs
* cpu_baseline : the UNMODIFIED reference CPU engine (oracle/_ref, c_cuda/cpu_engine.h) on this box's cores, on a
              bounded number of steps of the same workload
`--impl reference` times only that CPU engine (rank 0) and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
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
    "c5": dict(N=(2048, 2048, 1024), precision=1, fcc=False, nmat=1, mb=11, rigid=False,
               desc="BASELINE configs[4]: shoebox 2048x2048x1024, 7-pt Cartesian, fp32, lossy walls, x-slabs + NCCL halo exchange"),
    "small": dict(N=(128, 96, 64), precision=1, fcc=False, nmat=1, mb=11, rigid=False, desc="smoke-sized shoebox"),
    # the real rooms at full size, voxelised by the reference tool chain (tools/make_large_models.py -> data_large/, not committed)
    "ctk_real": dict(folder="data_large/ctk_cart_gpu", precision=1, fcc=False,
                     desc="BASELINE configs[1] with the real model: CTK church, h = 0.041 m, 7-pt Cartesian, fp32, 8 materials x 11 branches"),
    "mv_real": dict(folder="data_large/mv_fcc_gpu", precision=1, fcc=True,
                    desc="BASELINE configs[2] with the real model: Musikverein, h = 0.06 m, 13-pt FCC folded, fp32, 5 materials x 11 branches"),
}
BYTES_PER_NODE = {1: 12.125, 2: 24.125}  # u1 read + u0 read + u0 write + 1 mask bit
# CPU-arm sample grids for workloads whose full grid would take the CPU engine minutes per step
CPU_SAMPLE_GRID = {
    "c5": ((258, 2048, 1024), "one of the eight x-slabs of the c5 grid (256 planes + 2 halo planes, 258x2048x1024)"),
    "c4": ((130, 1024, 1024), "one eighth of the c4 grid (130x1024x1024)"),
}


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
    """the unmodified reference CPU engine on `wl`: -> (Gvox/s, cores, steps timed, seconds, kind)"""
    import tempfile
    from oracle import Reference
    w = WORKLOADS[wl]
    cores = threads or os.cpu_count() or 1
    sample_note = f"the full {wl} grid"
    if wl in CPU_SAMPLE_GRID:  # the grid is too large for a CPU run of a few minutes: time one slab of it
        WORKLOADS[wl + "_cpu_sample"] = dict(w, N=CPU_SAMPLE_GRID[wl][0])
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
    # bound the sample: size it for ~0.8 Gvox/s (the pool's 16-core hosts measure 1.2-3.2), then report what was actually run
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--air-kernel", type=int, default=1, choices=(0, 1))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--xc", type=int, default=0)
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE", help="engine option (pffdtd_set_option), repeatable (tuning / A-B runs)")
    ap.add_argument("--air-cfg", type=int, default=-1, help="tile configuration of the TMA air kernel (tuning)")
    ap.add_argument("--overlap", type=int, default=1, choices=(0, 1), help="N>1: 0 = halo exchange after the whole step (diagnostic)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    N = args.gpus
    if world != N and world > 1:
        raise SystemExit(f"--gpus {N} but WORLD_SIZE={world}")
    wl = args.workload or ("c2" if N == 1 else "c5")
    w = WORKLOADS[wl]
    K, W = args.steps, max(args.warmup, 3)
    if "folder" in w:  # stored grid of the folder (already folded for FCC)
        from pffdtd_b200 import h5lite
        v = h5lite.File(ROOT / w["folder"] / "vox_out.h5")
        Nx, Ny, Nz = (int(v[k][()]) for k in ("Nx", "Ny", "Nz"))
        Ny_st = Ny
        w = dict(w, N=(Nx, Ny, Nz))
        WORKLOADS[wl] = w
    else:
        Nx, Ny, Nz = w["N"]
        Ny_st = Ny // 2 + 1 if w["fcc"] else Ny
    Npts = Nx * Ny_st * Nz
    config = {"workload": wl, "description": w["desc"], "grid": [Nx, Ny_st, Nz], "stencil": "13pt_fcc_folded" if w["fcc"] else "7pt_cartesian",
              "l2": "state (2 grids, %.0f MB) is larger than the 126 MB L2" % (2 * Npts * (4 if w["precision"] == 1 else 8) / 1e6)}

    if args.impl == "reference":
        if rank != 0:
            return
        v, cores, k, t, kind, note = run_reference_cpu(wl, K, W)
        line = {"impl": "reference", "metric": "Gvoxel-updates/s", "value": v, "unit": "Gvox/s", "n_gpus": N, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * t / k, "higher_is_better": True, "scaling": "strong" if N > 1 else "weak", "vs_baseline": None,
                "dtype": "f32" if w["precision"] == 1 else "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "Gvox/s", "cores": cores, "kind": kind,
                                 "sample": f"{k} time steps (of the {K} requested) on {note}, unmodified c_cuda/cpu_engine.h run_sim, OpenMP {cores} threads"},
                "e2e": {"value": v, "unit": "Gvox/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from pffdtd_b200 import shoebox
    from pffdtd_b200.engine import Engine, comm_unique_id

    Nt = W + 2 * K  # warm-up, the timed K steps, the K steps of the roofline pass
    t_prep = time.perf_counter()
    # a rank only generates the boundary / shell nodes of its own slab (+ halo planes): the node lists of the
    # 2048x2048x1024 grid have 5e7 entries
    xr = None
    if world > 1 and not w["fcc"]:
        from pffdtd_b200.sim_data import SimData
        starts, sizes = SimData.slab_planes(Nx, world)
        xr = (max(0, starts[rank] - 1), min(Nx, starts[rank] + sizes[rank] + 1))
    files = build_problem(wl, Nt, x_range=xr)
    sd_full = shoebox.sim_data_from_files(files, w["precision"], abc_x_range=xr).scale_input()
    del files
    sd = sd_full.slab(rank, world) if world > 1 else sd_full
    eng = Engine(sd, local)
    if world > 1:
        box = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        eng.comm_init(box[0], rank, world)
    eng.set_option("air_kernel", args.air_kernel)
    if world > 1:
        eng.set_option("overlap", args.overlap)
        config["halo_exchange"] = "ncclSend/ncclRecv of one plane per neighbour per step, " + (
            "overlapped with the interior update on a second stream" if args.overlap else "after the step (not overlapped)")
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

    def barrier():
        eng.sync()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident run
    sampler = ClockSampler(nvml_index(local))
    sampler.start()
    eng.run_steps(0, W)
    barrier()
    sampler.begin()
    eng.reset_stats()
    eng.stat("timer_start")
    eng.run_steps(W, K)
    ms = eng.stat("timer_stop_ms")
    barrier()
    launches = eng.stat("launches")
    # roofline pass: the next K steps with CUDA events around every air launch (events cannot sit inside the
    # replayed CUDA graph of the pass above, so this pass launches kernel by kernel; same kernels, same data)
    eng.set_option("profile_air", 1)
    eng.reset_stats()
    eng.stat("timer_start")
    eng.run_steps(W + K, K)
    ms_prof = eng.stat("timer_stop_ms")
    barrier()
    eng.set_option("profile_air", 0)
    clocks = sampler.result()
    air_ms = eng.stat("air_ms")
    air_n = eng.stat("air_launches_timed")
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        t = torch.tensor([launches], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        launches = float(t.item())
    value = Npts * K / (ms * 1e-3) / 1e9

    # ---- end to end through host buffers: a fresh engine state is not needed, the work per step is identical
    e2e = None
    if not args.no_e2e:
        eng2 = eng
        ins = np.ascontiguousarray(sd.in_sigs.T)  # [Nt][Ns] host samples
        for n in range(min(W, 3)):
            eng2.step_host(n, ins[n])
        barrier()
        t0 = time.perf_counter()
        for n in range(W, W + K):
            eng2.step_host(n, ins[n])
        eng2.sync()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
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
        if tr and args.air_kernel == 1 and world == 1:
            traffic = tr["traffic_bytes"]
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "hbm", "kernel": ("k_air_tma_cart<FCC>" if w["fcc"] else "k_air_tma_cart") if args.air_kernel == 1 else "k_air_generic",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "algorithmic_bytes_per_launch": BYTES_PER_NODE[w["precision"]] * nodes_per_launch,
                "peak_source": peak_src, "bytes_per_node": BYTES_PER_NODE[w["precision"]], "nodes_per_launch": int(nodes_per_launch),
                "air_ms_per_step": air_ms_per_step, "air_launches_per_step": n_air_per_step, "air_share_of_step": air_ms_per_step / (ms_prof / K),
                "note": "air launches timed with CUDA events in a second pass of K steps launched kernel by kernel (%.4f ms/step); "
                        "the `value` pass replays the step as a CUDA graph when it can" % (ms_prof / K)}

    line = {"metric": "Gvoxel-updates/s", "value": value, "unit": "Gvox/s", "n_gpus": N, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "strong" if N > 1 else "weak", "vs_baseline": None,
            "dtype": "f32" if w["precision"] == 1 else "f64", "data": "synthetic", "config": config, "roofline": roofline,
            "clocks": clocks, "gpu_launches": int(launches), "host_prep_s": round(t_prep, 2)}
    if e2e:
        line["e2e"] = e2e
    if N > 1:
        # N = 1 of this script runs configs[1] (c2); the one-GPU number of THIS workload, for whoever computes a scaling efficiency,
        # is the committed measurement of `bench.py --workload c5` (same grid on one B200)
        try:
            one = json.loads((ROOT / "profiles" / f"r01_bench_{wl}_1gpu.json").read_text())
            line["same_workload_on_one_gpu"] = {"value": one["value"], "unit": one["unit"], "source": f"profiles/r01_bench_{wl}_1gpu.json",
                                                "speedup": value / one["value"]}
        except Exception:  # noqa: BLE001
            pass
    eng.close()

    # ---- the reference's CPU engine on this box's cores (rank 0, N=1 only)
    if rank == 0 and N == 1 and not args.no_cpu:
        try:
            v, cores, k, t, kind, note = run_reference_cpu(wl, 300, 3, budget_s=25.0)
            line["cpu_baseline"] = {"value": v, "unit": "Gvox/s", "cores": cores, "kind": kind,
                                    "sample": f"{k} time steps on {note}, unmodified c_cuda/cpu_engine.h run_sim, OpenMP {cores} threads, {t:.1f} s"}
        except Exception as ex:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "Gvox/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex!r}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
