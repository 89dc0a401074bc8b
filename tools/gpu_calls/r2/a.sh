#!/bin/bash
# r2 call A (1 GPU): gpu suite after the graph / chunk-plan refactor, then the new default bench (c5 + also + cpu baseline + parity hash)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/r2a_gpu.txt; free -g >> $O/r2a_gpu.txt; nproc >> $O/r2a_gpu.txt
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > $O/r2a_pytest.log 2>&1
( time timeout 900 python bench.py ) > $O/r2a_bench.json 2> $O/r2a_bench.err
( time timeout 300 python bench.py --impl reference --steps 20 --warmup 5 ) > $O/r2a_bench_ref.json 2> $O/r2a_bench_ref.err
tail -n 8 $O/r2a_pytest.log; cat $O/r2a_bench.json; tail -n 5 $O/r2a_bench.err; cat $O/r2a_bench_ref.json; tail -n 3 $O/r2a_bench_ref.err; cat $O/r2a_gpu.txt
