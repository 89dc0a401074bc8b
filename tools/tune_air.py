"""Development aid (GPU): time the tiled air kernel for each tile configuration / chunk length on one
workload, after checking each configuration bit for bit against the generic kernel on a ragged grid.

    python tools/tune_air.py [--workload c2] [--cfgs 0,1,2] [--xcs 0,16,32] [--steps 60]
"""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402
from pffdtd_b200 import shoebox  # noqa: E402
from pffdtd_b200.engine import Engine  # noqa: E402


def check(cfg, precision):
    from cases import make_sim_data, noise_grids
    sd = make_sim_data("cart_wide", precision)
    g1, g0 = noise_grids(sd)
    outs = []
    for ak in (0, 1):
        with Engine(sd) as e:
            e.set_option("air_kernel", ak)
            if ak:
                e.set_option("air_cfg", cfg)
            e.write_grid(1, g1)
            e.write_grid(0, g0)
            e.run_steps(0, 12)
            outs.append((e.read_grid(1), e.read_grid(0)))
    return all(np.array_equal(a[1:-1, 1:-1, 1:-1], b[1:-1, 1:-1, 1:-1]) for a, b in zip(*outs))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--cfgs", default="0,1,2,3,4,5,6,7")
    ap.add_argument("--xcs", default="0")
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--grid", default=None, help="Nx,Ny,Nz override of the workload's grid")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--fuse", default="1")
    a = ap.parse_args()
    w = dict(bench.WORKLOADS[a.workload])
    if a.grid:
        w["N"] = tuple(int(v) for v in a.grid.split(","))
        bench.WORKLOADS[a.workload] = w
    Nt = a.steps + 10
    files = bench.build_problem(a.workload, Nt)
    sd = shoebox.sim_data_from_files(files, w["precision"]).scale_input()
    peak, _ = bench.peaks()
    nodes = (sd.Nx - 2) * sd.Ny * sd.Nz
    for cfg in [int(c) for c in a.cfgs.split(",")]:
        ok = a.no_check or check(cfg, w["precision"])
        for xc, fuse in [(int(x), int(f)) for x in a.xcs.split(",") for f in a.fuse.split(",")]:
            with Engine(sd) as e:
                e.set_option("air_cfg", cfg)
                e.set_option("fuse", fuse)
                e.set_option("air_xc", xc)
                e.run_steps(0, 10)
                e.sync()
                e.reset_stats()
                e.set_option("profile_air", 1)
                e.stat("timer_start")
                e.run_steps(10, a.steps)
                ms = e.stat("timer_stop_ms")
                air = e.stat("air_ms") / a.steps
            gbs = bench.BYTES_PER_NODE[w["precision"]] * nodes / (air * 1e-3) / 1e9
            print(f"cfg {cfg} xc {xc:3d} fuse {fuse} parity {'ok' if ok else 'FAIL'}  air {air*1e3:8.1f} us  {gbs:7.0f} GB/s  {gbs/peak*100:5.1f}% of measured peak"
                  f"   step {ms/a.steps*1e3:8.1f} us", flush=True)


if __name__ == "__main__":
    main()
