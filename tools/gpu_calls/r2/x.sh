#!/bin/bash
# r2 call X (1 GPU): z_edge handling as its own kernel variant (ZE): parity, c2 / ctk / c5 lines back to call V's numbers
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_obstacles.py tests/test_gpu_slabs.py -x -q -m gpu ) > $O/r2x_pytest.log 2>&1
tail -n 4 $O/r2x_pytest.log
b() { name=$1; shift; timeout 300 python bench.py --no-cpu --no-also --no-parity "$@" > $O/r2x_$name.json 2> $O/r2x_$name.err; python - <<PY
import json
try:
    d=json.load(open("$O/r2x_$name.json")); r=d["roofline"]
    print("$name", "value %.1f e2e %.1f ms %.4f air_ms %.4f air_frac %.3f whole %.3f launches %d" % (d["value"], d.get("e2e",{}).get("value",0), d["ms_per_step"], r["air_ms_per_step"], r["frac"], r["whole_step_frac"], d["gpu_launches"]), d["config"].get("air_tile"))
except Exception as ex:
    print("$name failed", ex); print(open("$O/r2x_$name.err").read()[-800:])
PY
}
b c2 --workload c2 --steps 200
b ctk --workload ctk_real --steps 200
b c5 --workload c5 --steps 20 --no-e2e
