#!/bin/bash
# GPU call E (8 GPUs): full-size slab == single-GPU bitwise check, c5 bench at N=8 and N=4
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522"
timeout 240 $TR8 tools/mgpu_equal.py --workload c5 --steps 48 --compare profiles/r01_c5_traces_1gpu.npy > $O/e_equal_n8.log 2>&1
timeout 240 $TR8 bench.py --gpus 8 --steps 200 --warmup 10 > $O/e_bench_c5_n8.json 2> $O/e_bench_c5_n8.err
timeout 240 $TR4 bench.py --gpus 4 --steps 100 --warmup 10 --no-e2e > $O/e_bench_c5_n4.json 2> $O/e_bench_c5_n4.err
grep -i "world\|BITWISE\|error" $O/e_equal_n8.log | tail -5; cat $O/e_bench_c5_n8.json $O/e_bench_c5_n4.json; tail -n 3 $O/e_bench_c5_n8.err
