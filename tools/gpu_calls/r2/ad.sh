#!/bin/bash
# r2 call AD (1 GPU, the round's last seconds): where pffdtd_vox_run's wall time goes at production size
cd "$GRAFT_REPO_ROOT" || exit 1
PFFDTD_VOX_TIMING=1 timeout 40 python - 2>&1 <<'PY' | tee gpurun_out/r2ad_vox_timing.log | tail -n 12
import time, numpy as np, sys
sys.path.insert(0, ".")
from pffdtd_b200 import vox_accel as va
z = np.load("data_large/vox_mv_h003.npz")
inp = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
for rep in range(2):
    t0 = time.perf_counter(); r = va.ray_stage(inp, 0); print("ray_stage wall %.3f s, Nb %d" % (time.perf_counter() - t0, r[0].size), flush=True)
PY
