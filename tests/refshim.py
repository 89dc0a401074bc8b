"""Harness glue for running the UNMODIFIED reference Python code in this container: the
image has no h5py, so `h5py` is served by the repo's h5lite (which is exactly what needs validating against
the reference's access pattern); numpy aliases removed in numpy 2 are restored.  Test-only."""
import sys
import types
from pathlib import Path

import numpy as np

REF_PY = Path("/root/reference/python")


def available():
    return (REF_PY / "fdtd" / "rotate_sim_data.py").exists()


def install():
    from pffdtd_b200 import h5lite
    if "h5py" not in sys.modules:
        m = types.ModuleType("h5py")
        m.File = h5lite.File
        sys.modules["h5py"] = m
    sys.modules.setdefault("memory_profiler", types.SimpleNamespace(profile=lambda f: f))
    for name, val in (("float", float), ("bool8", np.bool_), ("float_", np.float64)):
        if not hasattr(np, name):
            setattr(np, name, val)
    if str(REF_PY) not in sys.path:
        sys.path.insert(0, str(REF_PY))
