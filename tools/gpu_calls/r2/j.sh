#!/bin/bash
# r2 call J (1 GPU): 7-point kernel back to its inline epilogue (c2 line), fused 13-point step as its own kernel: where does it differ?
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 300 python tests/diag/fused_fcc_diff.py > $O/r2j_diff.log 2>&1; tail -n 60 $O/r2j_diff.log
( time timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py tests/test_zz_obstacles.py -q -m gpu ) > $O/r2j_pytest.log 2>&1
tail -n 12 $O/r2j_pytest.log
b() { name=$1; shift; timeout 300 python bench.py --no-cpu --no-also --no-parity "$@" > $O/r2j_$name.json 2> $O/r2j_$name.err; python - <<PY
import json
try:
    d=json.load(open("$O/r2j_$name.json")); r=d["roofline"]
    print("$name", "value %.1f e2e %.1f ms %.4f air_ms %.4f air_frac %.3f whole %.3f launches %d" % (d["value"], d.get("e2e",{}).get("value",0), d["ms_per_step"], r["air_ms_per_step"], r["frac"], r["whole_step_frac"], d["gpu_launches"]), d["config"].get("air_tile"))
except Exception as ex:
    print("$name failed", ex); print(open("$O/r2j_$name.err").read()[-800:])
PY
}
b c2 --workload c2 --steps 200
b c3s --workload c3s --steps 100
b c3s_fuse0 --workload c3s --steps 100 --opt fuse=0
b mvbig --workload mv_big --steps 40
