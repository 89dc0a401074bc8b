"""does the CUDA path equal the oracle on one case?  python tests/diag/repro.py <case> <precision>   (diagnostic, uses the test oracle)"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT))
import numpy as np
from cases import make_sim_data
from oracle import Oracle
from pffdtd_b200.engine import Engine
name, prec = sys.argv[1], int(sys.argv[2])
sd = make_sim_data(name, prec)
ref = Oracle(sd).run_all()
for ak, fuse in ((1, 0), (1, 1)):
    try:
        with Engine(sd) as e:
            e.set_option("air_kernel", ak); e.set_option("fuse", fuse)
            e.run_steps(0, sd.Nt); got = e.read_outputs()
        print(name, prec, ak, fuse, np.array_equal(got, ref), flush=True)
    except Exception as ex:
        print(name, prec, ak, fuse, "EXC", ex, flush=True); break
