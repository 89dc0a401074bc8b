#!/bin/bash
# r2 call M (1 GPU): k_fd issues its loads before the table staging: parity subset, c2 / c5 lines, ncu of k_fd
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "traces or noise" ) > $O/r2m_pytest.log 2>&1
tail -n 4 $O/r2m_pytest.log
b() { name=$1; shift; timeout 300 python bench.py --no-cpu --no-also --no-parity "$@" > $O/r2m_$name.json 2> $O/r2m_$name.err; python - <<PY
import json
try:
    d=json.load(open("$O/r2m_$name.json")); r=d["roofline"]
    print("$name", "value %.1f e2e %.1f ms %.4f air_ms %.4f air_frac %.3f whole %.3f launches %d" % (d["value"], d.get("e2e",{}).get("value",0), d["ms_per_step"], r["air_ms_per_step"], r["frac"], r["whole_step_frac"], d["gpu_launches"]), d["config"].get("air_tile"))
except Exception as ex:
    print("$name failed", ex); print(open("$O/r2m_$name.err").read()[-800:])
PY
}
b c2 --workload c2 --steps 200
b c5 --workload c5 --steps 20 --no-e2e
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_fd -s 6 -c 1 -o $O/r2m_fd_c2 -f python bench.py --workload c2 --steps 10 --warmup 4 --no-cpu --no-also --no-parity --no-e2e > $O/r2m_ncu.log 2>&1
tail -n 2 $O/r2m_ncu.log
