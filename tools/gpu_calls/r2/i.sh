#!/bin/bash
# r2 call I (1 GPU): fused FCC step (mirror-on-write with halo edges + seam, shell by k_abc_faces + service warp), phase-split step: whole gpu suite, FCC lines
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 1800 python -m pytest tests -x -q -m gpu ) > $O/r2i_pytest.log 2>&1
tail -n 6 $O/r2i_pytest.log
b() { name=$1; shift; timeout 300 python bench.py --no-cpu --no-also --no-parity "$@" > $O/r2i_$name.json 2> $O/r2i_$name.err; python - <<PY
import json
try:
    d=json.load(open("$O/r2i_$name.json")); r=d["roofline"]
    print("$name", "value %.1f e2e %.1f ms %.4f air_ms %.4f air_frac %.3f whole %.3f launches %d" % (d["value"], d.get("e2e",{}).get("value",0), d["ms_per_step"], r["air_ms_per_step"], r["frac"], r["whole_step_frac"], d["gpu_launches"]), d["config"].get("air_tile"))
except Exception as ex:
    print("$name failed", ex); print(open("$O/r2i_$name.err").read()[-800:])
PY
}
b c3s --workload c3s --steps 100
b c3s_fuse0 --workload c3s --steps 100 --opt fuse=0
b mvbig --workload mv_big --steps 40
b mvreal --workload mv_real --steps 150
b c2 --workload c2 --steps 200
