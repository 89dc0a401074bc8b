#!/bin/bash
# r2 call Y (1 GPU): the voxeliser's ray casting on the GPU against the reference's golden vectors; timing
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 300 python -m pytest tests/test_vox.py -x -q -m gpu ) > $O/r2y_pytest.log 2>&1
tail -n 5 $O/r2y_pytest.log
timeout 120 python - > $O/r2y_time.log 2>&1 <<'PY'
import time, numpy as np, sys
sys.path.insert(0, ".")
from pffdtd_b200 import vox_accel as va
for name in ("ctk_h030", "ctk_h045_fcc", "mv_h060_fcc"):
    z = np.load(f"tests/golden/vox_{name}.npz")
    inp = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    va.ray_stage(inp, 0)
    t0 = time.perf_counter()
    bn, adj, tidx, nd = va.ray_stage(inp, 0)
    t1 = time.perf_counter()
    print(name, "Nb", bn.size, "voxels", inp["vox_start"].shape[0], "tris", inp["unor"].shape[0], "pffdtd_vox_run incl. upload / compaction: %.1f ms" % (1e3 * (t1 - t0)))
PY
cat $O/r2y_time.log
