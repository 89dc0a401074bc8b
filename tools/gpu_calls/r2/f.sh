#!/bin/bash
# r2 call F (2 GPUs): NCCL slab tests (captured step graphs with the exchange inside), single-process multi-GPU API on two devices,
# c5 on 2 GPUs: cost-weighted vs equal slabs, parity against the committed one-GPU hash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi -L > $O/r2f_gpus.txt
( time timeout 900 python -m pytest tests/test_gpu_slabs.py -x -q -m gpu ) > $O/r2f_pytest.log 2>&1
tail -n 6 $O/r2f_pytest.log
run2() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 "$@" > $O/r2f_$name.json 2> $O/r2f_$name.err; python - <<PY
import json
try:
    d=json.load(open("$O/r2f_$name.json")); r=d["roofline"]
    print("$name", "value %.1f e2e %.1f ms %.4f air_frac %.3f whole %.3f launches %d" % (d["value"], d.get("e2e",{}).get("value",0), d["ms_per_step"], r["frac"], r["whole_step_frac"], d["gpu_launches"]), d["config"].get("slab_planes"))
    print("   parity", json.dumps(d.get("parity")))
    for p in d["config"].get("per_rank", []): print("   ", p)
except Exception as ex:
    print("$name failed", ex); print(open("$O/r2f_$name.err").read()[-1500:])
PY
}
run2 c5_n2 --steps 40 --warmup 6
run2 c5_n2_equal --steps 40 --warmup 6 --equal-slabs --no-parity
run2 c5_n2_nograph --steps 40 --warmup 6 --opt use_graph=0 --no-parity
