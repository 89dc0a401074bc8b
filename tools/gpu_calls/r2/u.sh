#!/bin/bash
# r2 call U (8 GPUs): c5 on 8 and 4 GPUs with the peer-memory halo exchange (weighted and equal slabs)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi -L > $O/r2u_gpus.txt
runN() { n=$1; name=$2; shift; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > $O/r2u_$name.json 2> $O/r2u_$name.err; python - <<PY
import json
try:
    line=[l for l in open("$O/r2u_$name.json") if l.startswith("{")][-1]
    d=json.loads(line); r=d["roofline"]
    print("$name", "value %.1f e2e %.1f ms %.4f air_frac %.3f whole %.3f launches %d" % (d["value"], d.get("e2e",{}).get("value",0), d["ms_per_step"], r["frac"], r["whole_step_frac"], d["gpu_launches"]), d["config"].get("slab_planes"))
    p=d.get("parity") or {}
    print("   parity equal", p.get("equal_to_one_gpu_run"), "reduced grid bit exact", (p.get("reduced_grid_vs_cpu_engine") or {}).get("bit_exact"), (p.get("reduced_grid_vs_cpu_engine") or {}).get("slab_planes"))
    for q in d["config"].get("per_rank", []): print("    rank %d planes %d air %.3f non-air %.3f ms" % (q["rank"], q["planes"], q["air_ms"], q["non_air_ms"]))
except Exception as ex:
    print("$name failed", ex); print(open("$O/r2u_$name.err").read()[-1500:])
PY
}
runN 8 c5_n8 --steps 100 --warmup 10
runN 8 c5_n8_equal --steps 100 --warmup 10 --equal-slabs --no-parity
runN 4 c5_n4 --steps 60 --warmup 6
