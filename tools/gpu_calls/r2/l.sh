#!/bin/bash
# r2 call L (1 GPU): grouped branch-state layout (k_fd streams one contiguous block per warp): whole gpu suite, c2 / c5 / ctk lines
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 1800 python -m pytest tests -x -q -m gpu ) > $O/r2l_pytest.log 2>&1
tail -n 6 $O/r2l_pytest.log
b() { name=$1; shift; timeout 300 python bench.py --no-cpu --no-also --no-parity "$@" > $O/r2l_$name.json 2> $O/r2l_$name.err; python - <<PY
import json
try:
    d=json.load(open("$O/r2l_$name.json")); r=d["roofline"]
    print("$name", "value %.1f e2e %.1f ms %.4f air_ms %.4f air_frac %.3f whole %.3f launches %d" % (d["value"], d.get("e2e",{}).get("value",0), d["ms_per_step"], r["air_ms_per_step"], r["frac"], r["whole_step_frac"], d["gpu_launches"]), d["config"].get("air_tile"))
except Exception as ex:
    print("$name failed", ex); print(open("$O/r2l_$name.err").read()[-800:])
PY
}
b c2 --workload c2 --steps 200
b c5 --workload c5 --steps 20 --no-e2e
b ctk --workload ctk_real --steps 200
b c3s --workload c3s --steps 100
b mvbig --workload mv_big --steps 40
