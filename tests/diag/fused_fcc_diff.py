"""Diagnostic (GPU): where does the fused 13-point step first differ from the oracle?  Prints, per case and step count, the differing
interior nodes with their coordinates / parity / shell class.  python tests/diag/fused_fcc_diff.py [case ...]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from cases import make_sim_data, noise_grids  # noqa: E402
from oracle import Oracle  # noqa: E402
from pffdtd_b200.engine import Engine  # noqa: E402

for name in (sys.argv[1:] or ["fcc1_lossy", "fcc2_lossy"]):
    for precision in (2, 1):
        sd = make_sim_data(name, precision)
        g1, g0 = noise_grids(sd)
        for nsteps in (1, 2, 3):
            o = Oracle(sd)
            o.write_grid(1, g1)
            o.write_grid(0, g0)
            o.run_steps(0, nsteps)
            with Engine(sd) as e:
                e.write_grid(1, g1)
                e.write_grid(0, g0)
                e.run_steps(0, nsteps)
                a, b = e.read_grid(1), o.read_grid(1)
                fused, svc = e.stat("fused"), e.stat("svc")
            d = np.argwhere(a[1:-1, 1:-1, 1:-1] != b[1:-1, 1:-1, 1:-1]) + 1
            print(f"{name} p{precision} steps={nsteps} fused={fused} svc={svc} grid {sd.Nx}x{sd.Ny}x{sd.Nz}: {len(d)} differing interior nodes")
            bn = set(sd.bn_ixyz.tolist())
            for x, y, z in d[:12]:
                lin = (x * sd.Ny + y) * sd.Nz + z
                print(f"   ({x},{y},{z}) parity {(x + y + z) & 1} boundary {lin in bn} got {a[x, y, z]!r} want {b[x, y, z]!r}")
            if len(d):
                break
