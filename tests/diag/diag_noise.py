"""where does the fused step differ from the oracle?  python tests/diag/diag_noise.py case precision cfg steps"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from cases import make_sim_data, noise_grids
from oracle import Oracle
from pffdtd_b200.engine import Engine
name, prec, cfg, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
sd = make_sim_data(name, prec)
g1, g0 = noise_grids(sd)
o = Oracle(sd); o.write_grid(1, g1); o.write_grid(0, g0); o.run_steps(0, steps)
for fuse in (0, 1):
    with Engine(sd) as e:
        e.set_option("air_cfg", cfg); e.set_option("fuse", fuse)
        e.write_grid(1, g1); e.write_grid(0, g0); e.run_steps(0, steps)
        for which in (1, 0):
            a, b = e.read_grid(which)[1:-1, 1:-1, 1:-1], o.read_grid(which)[1:-1, 1:-1, 1:-1]
            bad = np.argwhere(a != b) + 1
            print(f"{name} p{prec} cfg{cfg} fuse{fuse} steps{steps} grid{which}: {len(bad)} bad of {a.size}; dims {sd.Nx},{sd.Ny},{sd.Nz}")
            if len(bad):
                for ax, nm in enumerate("xyz"):
                    vals, cnt = np.unique(bad[:, ax], return_counts=True)
                    print("   ", nm, dict(zip(vals.tolist(), cnt.tolist())))
