"""Full-size slab check (BASELINE configs[4]): receiver traces of an N-GPU run of the 2048x2048x1024 fp32 grid must be
BITWISE equal to the 1-GPU run's -- the slab split does not change any node's arithmetic.

A source sits on every 256-plane slab interface of the 8-way split and a line of receivers crosses each interface,
so within `--steps` steps every halo exchange of the 8-GPU run carries non-zero data in both directions.

    python tools/mgpu_equal.py --steps 48 --out profiles/r01_c5_traces_1gpu.npy                      (1 GPU)
    torchrun --nproc-per-node 8 ... tools/mgpu_equal.py --steps 48 --compare profiles/r01_c5_traces_1gpu.npy
"""
import argparse
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c5")
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--out", default=None)
    ap.add_argument("--compare", default=None)
    ap.add_argument("--overlap", type=int, default=1)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    import torch
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from pffdtd_b200 import shoebox
    from pffdtd_b200.engine import Engine, comm_unique_id
    from pffdtd_b200.sim_data import SimData
    w = bench.WORKLOADS[args.workload]
    Nx, Ny, Nz = w["N"]
    Nt = args.steps
    xr = None
    if world > 1:
        starts, sizes = SimData.slab_planes(Nx, world)
        xr = (max(0, starts[rank] - 1), min(Nx, starts[rank] + sizes[rank] + 1))
    t0 = time.perf_counter()
    files = bench.build_problem(args.workload, Nt, x_range=xr)
    # sources on the interfaces of the 8-way split, receiver lines across them (same lists whatever `world` is)
    cm = files["comms_out"]
    sig = cm["in_sigs"][0] / np.abs(cm["in_sigs"][0]).max()
    faces = [Nx * k // 8 for k in range(1, 8)]
    node = lambda ix, iy, iz: (ix * Ny + iy) * Nz + iz
    in_ixyz = np.array([node(f - (k & 1), Ny // 2 + 5 * k, Nz // 2 - 3 * k) for k, f in enumerate(faces)], np.int64)
    out_ixyz = np.array(sorted(node(f + d, Ny // 2 + 5 * k + 2, Nz // 2 - 3 * k + 1) for k, f in enumerate(faces) for d in range(-20, 20)), np.int64)
    cm.update(in_ixyz=in_ixyz, in_sigs=np.stack([sig * (1.0 + 0.1 * k) for k in range(len(faces))]), Ns=np.int64(in_ixyz.size),
              out_ixyz=out_ixyz, out_reorder=np.arange(out_ixyz.size, dtype=np.int64), Nr=np.int64(out_ixyz.size))
    sd_full = shoebox.sim_data_from_files(files, w["precision"], abc_x_range=xr).scale_input()
    sd = sd_full.slab(rank, world) if world > 1 else sd_full
    eng = Engine(sd, local)
    if world > 1:
        box = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        eng.comm_init(box[0], rank, world)
        eng.set_option("overlap", args.overlap)
    t1 = time.perf_counter()
    eng.run_steps(0, Nt)
    eng.sync()
    u = eng.read_outputs(0, Nt)
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, u)
        u = np.concatenate(parts, axis=0)
    if rank == 0:
        nz = int((np.abs(u).max(axis=1) > 0).sum())
        print(f"world {world}: {u.shape[0]} receivers x {Nt} steps, {nz} non-silent, peak {np.abs(u).max():.6e}, prep {t1 - t0:.1f} s, "
              f"run {time.perf_counter() - t1:.1f} s", flush=True)
        if args.out:
            np.save(args.out, u)
        if args.compare:
            ref = np.load(args.compare)
            same = ref.shape == u.shape and np.array_equal(ref, u)
            print(f"BITWISE EQUAL to {args.compare}: {same}" + ("" if same else f" (max|d| {np.abs(ref - u).max():.3e})"), flush=True)
            if not same:
                sys.exit(1)
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
