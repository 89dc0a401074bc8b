#!/bin/bash
# r2 call AA (1 GPU): voxeliser with the compaction on the device; the Musikverein at the reference script's own size (2852x552x850 folded,
# 1.34 G nodes) probed at the surface against the unmodified reference CPU engine, and its bench line; c2 line with the new roofline fields
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 300 python -m pytest tests/test_vox.py -x -q -m gpu -s ) > $O/r2aa_vox.log 2>&1
grep -a "vox production\|passed\|failed\|rror" $O/r2aa_vox.log | tail -n 6
( time timeout 600 python -m pytest tests/test_large_models.py -x -q -m gpu -s -k musikverein ) > $O/r2aa_mv.log 2>&1
grep -a "^\[mv_\|passed\|failed\|rror\|max|d|" $O/r2aa_mv.log | tail -n 8
b() { name=$1; shift; timeout 400 python bench.py --no-cpu --no-also --no-parity "$@" > $O/r2aa_$name.json 2> $O/r2aa_$name.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/r2aa_$name.json") if l.startswith("{")][-1]); r=d["roofline"]
    print("$name", "value %.1f e2e %.1f ms %.4f air_frac %.3f whole %.3f with_state %.3f prep %.1f s launches %d" % (d["value"], (d.get("e2e") or {}).get("value",0), d["ms_per_step"], r["frac"], r["whole_step_frac"], r["whole_step_frac_with_boundary_state"], d.get("host_prep_s", 0), d["gpu_launches"]), d["config"].get("grid"), d["config"].get("air_tile"))
except Exception as ex:
    print("$name failed", ex); print(open("$O/r2aa_$name.err").read()[-800:])
PY
}
b mv_full --workload mv_full --steps 20
b c2 --workload c2 --steps 200
