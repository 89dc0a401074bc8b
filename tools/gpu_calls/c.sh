#!/bin/bash
# GPU call C (2 GPUs): NCCL slab tests, slab == single-GPU bitwise check, c5 bench at N=2
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi -L > $O/c_gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( time timeout 600 python -m pytest tests/test_gpu_slabs.py -m gpu -x -q ) > $O/c_pytest_slabs.log 2>&1
timeout 300 python tools/mgpu_equal.py --workload c2 --steps 48 --out $O/c_c2_traces_n1.npy > $O/c_equal_n1.log 2>&1
timeout 300 $TR tools/mgpu_equal.py --workload c2 --steps 48 --compare $O/c_c2_traces_n1.npy > $O/c_equal_n2.log 2>&1
timeout 300 $TR tools/mgpu_equal.py --workload c2 --steps 48 --overlap 0 --compare $O/c_c2_traces_n1.npy > $O/c_equal_n2_noov.log 2>&1
timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 10 > $O/c_bench_c5_n2.json 2> $O/c_bench_c5_n2.err
timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 10 --overlap 0 --no-e2e > $O/c_bench_c5_n2_noov.json 2> $O/c_bench_c5_n2_noov.err
tail -4 $O/c_pytest_slabs.log; tail -3 $O/c_equal_n1.log $O/c_equal_n2.log $O/c_equal_n2_noov.log; cat $O/c_bench_c5_n2.json $O/c_bench_c5_n2_noov.json; tail -5 $O/c_bench_c5_n2.err
