#!/bin/bash
# r2 call E (1 GPU): service-warp entries staged by bulk copy (no global loads in that warp); single-process multi-slab API; A/B lines
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_zz_obstacles.py tests/test_gpu_slabs.py -x -q -m gpu ) > $O/r2e_pytest.log 2>&1
tail -n 6 $O/r2e_pytest.log
b() { name=$1; shift; timeout 300 python bench.py --no-cpu --no-also --no-parity "$@" > $O/r2e_$name.json 2> $O/r2e_$name.err; python - <<PY
import json
try:
    d=json.load(open("$O/r2e_$name.json")); r=d["roofline"]
    print("$name", "value %.1f e2e %.1f ms %.4f air_ms %.4f air_frac %.3f whole %.3f launches %d" % (d["value"], d.get("e2e",{}).get("value",0), d["ms_per_step"], r["air_ms_per_step"], r["frac"], r["whole_step_frac"], d["gpu_launches"]))
except Exception as ex:
    print("$name failed", ex); print(open("$O/r2e_$name.err").read()[-800:])
PY
}
b c2_svc1 --workload c2 --steps 200
b c2_svc0 --workload c2 --steps 200 --opt svc=0
b c2_cap0 --workload c2 --steps 200 --opt svc_cap=0
b c2_cap128 --workload c2 --steps 200 --opt svc_cap=128
b ctk_svc1 --workload ctk_real --steps 200
b ctk_svc0 --workload ctk_real --steps 200 --opt svc=0
b ctk_cap128 --workload ctk_real --steps 200 --opt svc_cap=128
b c5_svc1 --workload c5 --steps 20 --no-e2e
b c4_svc1 --workload c4 --steps 20
