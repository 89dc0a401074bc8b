"""TEST INFRASTRUCTURE -- performance comparator (not part of the product, not part of bench.py): the reference's OWN CUDA engine
(c_cuda/gpu_engine.h, unmodified, compiled for sm_100a by oracle/Makefile into oracle/_ref/libpffdtd_refgpu_*.so) on the
same B200 and the same inputs as this repo's engine.  Prints one JSON line per workload:

    python tests/diag/compare_reference_gpu_engine.py --workload c2 --steps 200

Both engines run `--steps` time steps of the bench workload from the same in-memory files; throughput is
Npts*steps / loop seconds for both (the reference's own clock, gpu_engine.h:1253; ours the same wall clock around
pffdtd_run_steps + sync).  The reference engine ends with cudaDeviceReset(), so it runs last.  Traces are compared:
the reference GPU engine is not bit-identical to its CPU engine (nvcc contracts a*b+c into FMAs), so
the difference is reported relative to the trace peak.
"""
import argparse
import json
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--steps", type=int, default=200)
    args = ap.parse_args()
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", "0")  # the reference engine takes every visible device
    from oracle import Reference
    from pffdtd_b200 import shoebox
    from pffdtd_b200.engine import Engine
    w = bench.WORKLOADS[args.workload]
    Nx, Ny, Nz = w["N"]
    Npts = Nx * (Ny // 2 + 1 if w["fcc"] else Ny) * Nz
    files = bench.build_problem(args.workload, args.steps)
    # ours
    sd = shoebox.sim_data_from_files(files, w["precision"]).scale_input()
    with Engine(sd, 0) as e:
        e.run_steps(0, min(10, args.steps))  # warm-up: first launches, graph capture
        e.sync()
    with Engine(sd, 0) as e:
        t0 = time.perf_counter()
        e.run_steps(0, args.steps)
        e.sync()
        t_ours = time.perf_counter() - t0
        u_ours = sd.reorder_output(sd.rescale_output(e.read_outputs()))
    # the reference's CUDA engine, through its own load_sim_data / scale_input / run_sim / rescale_output / write_outputs
    tmp = tempfile.mkdtemp(prefix="pffdtd_refgpu_")
    for fn in ("sim_consts.h5", "vox_out.h5", "comms_out.h5", "sim_mats.h5"):
        (Path(tmp) / fn).touch()
    ref = Reference(w["precision"], files, tmp, gpu=True)
    sys.stdout.flush()
    saved, devnull = os.dup(1), os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        u_ref, t_ref = ref.run()
    finally:
        os.dup2(saved, 1)
    peak = float(np.abs(u_ref).max())
    line = {"workload": args.workload, "description": w["desc"], "steps": args.steps, "dtype": "f32" if w["precision"] == 1 else "f64",
            "reference_gpu_engine": {"value": Npts * args.steps / t_ref / 1e9, "unit": "Gvox/s", "seconds": t_ref,
                                     "what": "unmodified c_cuda/gpu_engine.h run_sim, nvcc -O3 sm_100a, 1 GPU, its own timer"},
            "pffdtd_b200": {"value": Npts * args.steps / t_ours / 1e9, "unit": "Gvox/s", "seconds": t_ours,
                            "what": "pffdtd_run_steps + sync, wall clock, same inputs"},
            "speedup": t_ref / t_ours,
            "traces": {"max_abs_diff_over_peak": float(np.abs(u_ref - u_ours).max() / peak) if peak > 0 else None, "peak": peak,
                       "note": "the reference GPU engine itself differs from the reference CPU engine by FMA contraction"}}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
