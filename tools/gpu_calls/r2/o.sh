#!/bin/bash
# r2 call O (1 GPU): round-end rehearsal + profile artefacts: gpu suite, smoke, default bench (c5 + also + cpu), reference arm,
# ncu launch lists (c2, ctk) and --set full captures of the step's kernels on c2
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 1800 python -m pytest tests -x -q -m gpu ) > $O/r2o_pytest.log 2>&1
tail -n 4 $O/r2o_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/r2o_smoke.log 2>&1; tail -n 3 $O/r2o_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/r2o_bench.json 2> $O/r2o_bench.err; tail -n 4 $O/r2o_bench.err
( time timeout 300 python bench.py --impl reference --steps 20 --warmup 5 ) > $O/r2o_bench_ref.json 2> $O/r2o_bench_ref.err
python - <<PY
import json
d=json.loads([l for l in open("$O/r2o_bench.json") if l.startswith("{")][-1])
print("c5 N=1 value %.1f e2e %.1f ms %.4f air_frac %.3f whole %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["whole_step_frac"]), d.get("parity",{}).get("equal_to_one_gpu_run"), d.get("parity",{}).get("reduced_grid_vs_cpu_engine",{}).get("bit_exact"), d["cpu_baseline"]["value"])
for k,v in d.get("also",{}).items():
    print(k, "value %.1f e2e %.1f ms %.4f air_frac %.3f whole %.3f" % (v["value"], (v.get("e2e") or {}).get("value",0), v["ms_per_step"], v["roofline"]["frac"], v["roofline"]["whole_step_frac"]) if "value" in v else v)
PY
for wl in c2 ctk_real; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file $O/r2o_launches_$wl.csv python bench.py --workload $wl --steps 10 --warmup 4 --no-cpu --no-also --no-parity --no-e2e --opt use_graph=0 > /dev/null 2>&1
done
timeout 400 ncu --set full --import-source on --clock-control none -k "regex:k_air_tma|k_rigid|k_abc_faces|k_fd|k_io" -s 18 -c 6 -o $O/r2o_step_c2 -f python bench.py --workload c2 --steps 10 --warmup 4 --no-cpu --no-also --no-parity --no-e2e --opt use_graph=0 > $O/r2o_ncu.log 2>&1; tail -n 2 $O/r2o_ncu.log
b() { name=$1; shift; timeout 300 python bench.py --no-cpu --no-also --no-parity "$@" > $O/r2o_$name.json 2> $O/r2o_$name.err; python - <<PY
import json
try:
    d=json.load(open("$O/r2o_$name.json")); r=d["roofline"]
    print("$name", "value %.1f e2e %.1f ms %.4f air_ms %.4f air_frac %.3f whole %.3f launches %d" % (d["value"], d.get("e2e",{}).get("value",0), d["ms_per_step"], r["air_ms_per_step"], r["frac"], r["whole_step_frac"], d["gpu_launches"]), d["config"].get("air_tile"))
except Exception as ex:
    print("$name failed", ex); print(open("$O/r2o_$name.err").read()[-800:])
PY
}
b ctk_xc8 --workload ctk_real --steps 200 --xc 8
b ctk_xc12 --workload ctk_real --steps 200 --xc 12
b c2_xc8 --workload c2 --steps 200 --xc 8
