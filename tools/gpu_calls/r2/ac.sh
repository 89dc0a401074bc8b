#!/bin/bash
# r2 call AC (1 GPU): whole GPU suite with the z halos written by k_abc in the unfused step (no k_flip_z), then the 13-point lines
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > $O/r2ac_pytest.log 2>&1
tail -n 6 $O/r2ac_pytest.log
b() { name=$1; shift; timeout 400 python bench.py --no-cpu --no-also --no-parity --no-e2e "$@" > $O/r2ac_$name.json 2> $O/r2ac_$name.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/r2ac_$name.json") if l.startswith("{")][-1]); r=d["roofline"]
    print("$name", "value %.1f ms %.4f air_ms %.4f air_frac %.3f whole %.3f with_state %.3f launches %d" % (d["value"], d["ms_per_step"], r["air_ms_per_step"], r["frac"], r["whole_step_frac"], r["whole_step_frac_with_boundary_state"], d["gpu_launches"]), d["config"].get("air_tile"))
except Exception as ex:
    print("$name failed", ex); print(open("$O/r2ac_$name.err").read()[-800:])
PY
}
b mv_full --workload mv_full --steps 20
b c3s --workload c3s --steps 100
b mv_real --workload mv_real --steps 150
