#!/bin/bash
# r2 call H (1 GPU): FCC service configuration with 11 consumer warps at 72 registers; k_fd with state loads decoupled from the material word
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 1800 python -m pytest tests -x -q -m gpu ) > $O/r2h_pytest.log 2>&1
tail -n 6 $O/r2h_pytest.log
b() { name=$1; shift; timeout 300 python bench.py --no-cpu --no-also --no-parity "$@" > $O/r2h_$name.json 2> $O/r2h_$name.err; python - <<PY
import json
try:
    d=json.load(open("$O/r2h_$name.json")); r=d["roofline"]
    print("$name", "value %.1f e2e %.1f ms %.4f air_ms %.4f air_frac %.3f whole %.3f launches %d" % (d["value"], d.get("e2e",{}).get("value",0), d["ms_per_step"], r["air_ms_per_step"], r["frac"], r["whole_step_frac"], d["gpu_launches"]))
except Exception as ex:
    print("$name failed", ex); print(open("$O/r2h_$name.err").read()[-800:])
PY
}
b c3s_svc1 --workload c3s --steps 100
b mvbig_svc1 --workload mv_big --steps 40
b mvreal_svc1 --workload mv_real --steps 150
b c2_svc1 --workload c2 --steps 200
b ctk_svc1 --workload ctk_real --steps 200
