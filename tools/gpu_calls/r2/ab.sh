#!/bin/bash
# r2 call AB (1 GPU): where the real 13-point rooms spend their step: launch list of mv_full; the service warp (rigid lists) on the real FCC rooms;
# kernel times of the voxeliser at production size
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
b() { name=$1; shift; timeout 400 python bench.py --no-cpu --no-also --no-parity --no-e2e "$@" > $O/r2ab_$name.json 2> $O/r2ab_$name.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/r2ab_$name.json") if l.startswith("{")][-1]); r=d["roofline"]
    print("$name", "value %.1f ms %.4f air_ms %.4f air_frac %.3f whole %.3f with_state %.3f launches %d" % (d["value"], d["ms_per_step"], r["air_ms_per_step"], r["frac"], r["whole_step_frac"], r["whole_step_frac_with_boundary_state"], d["gpu_launches"]), d["config"].get("air_tile"))
except Exception as ex:
    print("$name failed", ex); print(open("$O/r2ab_$name.err").read()[-800:])
PY
}
b mv_full_svc --workload mv_full --steps 20 --opt svc=1
b mv_big_svc --workload mv_big --steps 40 --opt svc=1
b mv_real_svc --workload mv_real --steps 150 --opt svc=1
b mv_real --workload mv_real --steps 150
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 24 --csv --log-file $O/r2ab_launches_mv_full.csv python bench.py --workload mv_full --steps 4 --warmup 3 --no-cpu --no-also --no-parity --no-e2e --opt use_graph=0 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/r2ab_launches_mv_full.csv") if l.startswith('"'))]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
for r in rows[1:25]: print(r[ik][:60], r[iv])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_vox --csv --log-file $O/r2ab_vox_kernels.csv python -m pytest tests/test_vox.py -x -q -m gpu -k production > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/r2ab_vox_kernels.csv") if l.startswith('"'))]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value"); iu=h.index("Metric Unit")
for r in rows[1:]: print(r[ik][:60], r[iv], r[iu])
PY
