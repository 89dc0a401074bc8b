#!/bin/bash
# r2 call V (1 GPU): round-end rehearsal on the final tree: gpu suite, smoke, default bench, reference arm; DRAM bytes of the c5 air launch
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > $O/r2v_pytest.log 2>&1
tail -n 4 $O/r2v_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/r2v_smoke.log 2>&1; tail -n 3 $O/r2v_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/r2v_bench.json 2> $O/r2v_bench.err; tail -n 4 $O/r2v_bench.err
python - <<PY
import json
d=json.loads([l for l in open("$O/r2v_bench.json") if l.startswith("{")][-1])
print("c5 N=1 value %.1f e2e %.1f ms %.4f air_frac %.3f whole %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["whole_step_frac"]), d.get("parity",{}).get("equal_to_one_gpu_run"), d.get("parity",{}).get("reduced_grid_vs_cpu_engine",{}).get("bit_exact"), d["cpu_baseline"]["value"], d["gpu_launches"])
for k,v in d.get("also",{}).items():
    print(k, "value %.1f e2e %.1f ms %.4f air_frac %.3f whole %.3f" % (v["value"], (v.get("e2e") or {}).get("value",0), v["ms_per_step"], v["roofline"]["frac"], v["roofline"]["whole_step_frac"]) if "value" in v else v)
PY
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_air_tma -s 6 -c 1 --csv --log-file $O/r2v_c5_air_dram.csv python bench.py --workload c5 --steps 4 --warmup 4 --no-cpu --no-also --no-parity --no-e2e > /dev/null 2>&1
tail -n 4 $O/r2v_c5_air_dram.csv | cut -c1-400
