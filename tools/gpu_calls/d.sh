#!/bin/bash
# GPU call D (1 GPU): single-GPU traces of the full c5 grid (the comparison target of the 8-GPU run), reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python tools/mgpu_equal.py --workload c5 --steps 48 --out $O/d_c5_traces_n1.npy > $O/d_equal_n1.log 2>&1
( time timeout 600 python bench.py --impl reference ) > $O/d_bench_reference_c2.json 2> $O/d_bench_reference_c2.err
grep world $O/d_equal_n1.log; cat $O/d_bench_reference_c2.json; tail -4 $O/d_bench_reference_c2.err
