#!/bin/bash
# r2 call C (1 GPU): packed fp32 arithmetic (FFMA2/FADD2) in the 7-point consumer + the shorter k_fd: parity, then A/B lines
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_zz_obstacles.py tests/test_gpu_slabs.py -x -q -m gpu ) > $O/r2c_pytest.log 2>&1
tail -n 6 $O/r2c_pytest.log
b() { name=$1; shift; timeout 300 python bench.py --no-cpu --no-also --no-parity "$@" > $O/r2c_$name.json 2> $O/r2c_$name.err; python - <<PY
import json
try:
    d=json.load(open("$O/r2c_$name.json")); r=d["roofline"]
    print("$name", "value %.1f e2e %.1f ms %.4f air_ms %.4f air_frac %.3f whole %.3f launches %d" % (d["value"], d.get("e2e",{}).get("value",0), d["ms_per_step"], r["air_ms_per_step"], r["frac"], r["whole_step_frac"], d["gpu_launches"]))
except Exception as ex:
    print("$name failed", ex); print(open("$O/r2c_$name.err").read()[-800:])
PY
}
b c2_svc1 --workload c2 --steps 200
b c2_svc0 --workload c2 --steps 200 --opt svc=0
b c2_svc0_fd0 --workload c2 --steps 200 --opt svc=0 --opt fd_fixed=0
b c2_cap0 --workload c2 --steps 200 --opt svc_cap=0
b ctk_svc1 --workload ctk_real --steps 200
b ctk_svc0 --workload ctk_real --steps 200 --opt svc=0
b c5_svc1 --workload c5 --steps 20 --no-e2e
b c5_svc0 --workload c5 --steps 20 --no-e2e --opt svc=0
b c4_svc1 --workload c4 --steps 20
b mvbig --workload mv_big --steps 40
b c3s --workload c3s --steps 100
