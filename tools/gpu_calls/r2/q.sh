#!/bin/bash
# r2 call Q (1 GPU): edge planes of a two-neighbour slab on two streams (multi-slab tests on one device), k_abc_faces 2-D, k_io tick: gpu suite + lines + launch lists
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 1800 python -m pytest tests -x -q -m gpu ) > $O/r2q_pytest.log 2>&1
tail -n 4 $O/r2q_pytest.log
b() { name=$1; shift; timeout 300 python bench.py --no-cpu --no-also --no-parity "$@" > $O/r2q_$name.json 2> $O/r2q_$name.err; python - <<PY
import json
try:
    d=json.load(open("$O/r2q_$name.json")); r=d["roofline"]
    print("$name", "value %.1f e2e %.1f ms %.4f air_ms %.4f air_frac %.3f whole %.3f launches %d" % (d["value"], d.get("e2e",{}).get("value",0), d["ms_per_step"], r["air_ms_per_step"], r["frac"], r["whole_step_frac"], d["gpu_launches"]), d["config"].get("air_tile"))
except Exception as ex:
    print("$name failed", ex); print(open("$O/r2q_$name.err").read()[-800:])
PY
}
b c2 --workload c2 --steps 200
b ctk --workload ctk_real --steps 200
b c5 --workload c5 --steps 20 --no-e2e
for wl in c2 ctk_real c3s; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file $O/r2q_launches_$wl.csv python bench.py --workload $wl --steps 10 --warmup 4 --no-cpu --no-also --no-parity --no-e2e --opt use_graph=0 > /dev/null 2>&1
done
