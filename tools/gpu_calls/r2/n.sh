#!/bin/bash
# r2 call N (1 GPU): k_fd_bulk (branch state through the TMA unit): parity of everything with lossy walls, A/B lines, ncu
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 1800 python -m pytest tests -x -q -m gpu ) > $O/r2n_pytest.log 2>&1
tail -n 5 $O/r2n_pytest.log
b() { name=$1; shift; timeout 300 python bench.py --no-cpu --no-also --no-parity "$@" > $O/r2n_$name.json 2> $O/r2n_$name.err; python - <<PY
import json
try:
    d=json.load(open("$O/r2n_$name.json")); r=d["roofline"]
    print("$name", "value %.1f e2e %.1f ms %.4f air_ms %.4f air_frac %.3f whole %.3f launches %d" % (d["value"], d.get("e2e",{}).get("value",0), d["ms_per_step"], r["air_ms_per_step"], r["frac"], r["whole_step_frac"], d["gpu_launches"]), d["config"].get("air_tile"))
except Exception as ex:
    print("$name failed", ex); print(open("$O/r2n_$name.err").read()[-800:])
PY
}
b c2 --workload c2 --steps 200
b c2_fd0 --workload c2 --steps 200 --opt fd_bulk=0
b c5 --workload c5 --steps 20 --no-e2e
b ctk --workload ctk_real --steps 200
b mvbig --workload mv_big --steps 40
b c3s --workload c3s --steps 100
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_fd -s 6 -c 1 -o $O/r2n_fd_c2 -f python bench.py --workload c2 --steps 10 --warmup 4 --no-cpu --no-also --no-parity --no-e2e > $O/r2n_ncu.log 2>&1
tail -n 2 $O/r2n_ncu.log
