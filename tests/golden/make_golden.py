"""Generates tests/golden/traces_*.npz with the UNMODIFIED reference CPU engine (oracle/_ref, built from
/root/reference/c_cuda by oracle/Makefile): load_sim_data -> scale_input -> run_sim -> rescale_output ->
write_outputs on the deterministic cases of tests/cases.py.  Run in the build container (where
/root/reference exists); the fixtures let the oracle be pinned where the reference is absent.

    python tests/golden/make_golden.py
"""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))
from cases import CASES, OBSTACLE_CASES, make_files  # noqa: E402
from oracle import Reference  # noqa: E402
from pffdtd_b200 import shoebox  # noqa: E402


def main():
    out = {}
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    for name in sorted(CASES) + sorted(OBSTACLE_CASES):
        files = make_files(name)
        for prec in (1, 2):
            d = tempfile.mkdtemp(prefix="golden_")
            shoebox.write_folder(files, d)
            sys.stdout.flush()
            os.dup2(devnull, 1)
            try:
                u, _ = Reference(prec, files, d).run()
            finally:
                os.dup2(saved, 1)
            out[f"{name}_p{prec}"] = u
            print(name, prec, u.shape, float(np.abs(u).max()))
    np.savez_compressed(HERE / "traces_ref_cpu_engine.npz", **out)
    print("wrote", HERE / "traces_ref_cpu_engine.npz", os.path.getsize(HERE / "traces_ref_cpu_engine.npz"), "bytes")


if __name__ == "__main__":
    main()
