#!/bin/bash
# r2 call T (2 GPUs): halo exchange over peer memory (CUDA IPC + device flags): slab tests, c5 on 2 GPUs with parity, NCCL A/B
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
export PFFDTD_TEST_TIMEOUT=60
( time timeout 400 python -m pytest tests/test_gpu_slabs.py -x -q -m gpu -k "nccl or distinct" ) > $O/r2t_pytest.log 2>&1
tail -n 4 $O/r2t_pytest.log
runN() { n=$1; name=$2; shift; shift; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > $O/r2t_$name.json 2> $O/r2t_$name.err; python - <<PY
import json
try:
    line=[l for l in open("$O/r2t_$name.json") if l.startswith("{")][-1]
    d=json.loads(line); r=d["roofline"]
    print("$name", "value %.1f e2e %.1f ms %.4f air_frac %.3f whole %.3f launches %d" % (d["value"], d.get("e2e",{}).get("value",0), d["ms_per_step"], r["frac"], r["whole_step_frac"], d["gpu_launches"]), d["config"].get("slab_planes"), d["config"].get("halo_exchange","")[:40])
    p=d.get("parity") or {}
    print("   parity equal", p.get("equal_to_one_gpu_run"), "reduced grid", (p.get("reduced_grid_vs_cpu_engine") or {}))
except Exception as ex:
    print("$name failed", ex); print(open("$O/r2t_$name.err").read()[-1500:])
PY
}
runN 2 c5_n2 --steps 40 --warmup 6
runN 2 c5_n2_nccl --steps 40 --warmup 6 --no-p2p --no-parity
runN 2 c5_n2_noov --steps 40 --warmup 6 --overlap 0 --no-parity
