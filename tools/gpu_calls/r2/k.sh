#!/bin/bash
# r2 call K (2 GPUs): NCCL slabs with the step graphs split around an eager exchange; c5 on 2 GPUs with parity
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
export PFFDTD_TEST_TIMEOUT=60
( time timeout 400 python -m pytest tests/test_gpu_slabs.py -x -q -m gpu -k "nccl or distinct" ) > $O/r2k_pytest.log 2>&1
tail -n 6 $O/r2k_pytest.log
run2() { name=$1; shift; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 "$@" > $O/r2k_$name.json 2> $O/r2k_$name.err; python - <<PY
import json
try:
    line=[l for l in open("$O/r2k_$name.json") if l.startswith("{")][-1]
    d=json.loads(line); r=d["roofline"]
    print("$name", "value %.1f e2e %.1f ms %.4f air_frac %.3f whole %.3f launches %d" % (d["value"], d.get("e2e",{}).get("value",0), d["ms_per_step"], r["frac"], r["whole_step_frac"], d["gpu_launches"]), d["config"].get("slab_planes"))
    print("   parity", json.dumps(d.get("parity")))
    for p in d["config"].get("per_rank", []): print("   ", p)
except Exception as ex:
    print("$name failed", ex); print(open("$O/r2k_$name.err").read()[-1500:])
PY
}
run2 c5_n2 --steps 40 --warmup 6
run2 c5_n2_equal --steps 40 --warmup 6 --equal-slabs --no-parity
