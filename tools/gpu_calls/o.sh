#!/bin/bash
# GPU call O (1 GPU): real rooms with the narrow-tile defaults: bench lines, then the full-size bit-exactness tests
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 100 python bench.py --workload ctk_real --steps 200 --no-cpu > $O/o_bench_ctk_real.json 2> $O/o_bench_ctk_real.err
timeout 100 python bench.py --workload mv_real --steps 150 --no-cpu > $O/o_bench_mv_real.json 2> $O/o_bench_mv_real.err
timeout 100 python bench.py --no-cpu > $O/o_bench_c2.json 2> $O/o_bench_c2.err
( time timeout 200 python -m pytest tests/test_large_models.py -m gpu -q ) > $O/o_pytest_large.log 2>&1
cat $O/o_bench_ctk_real.json $O/o_bench_mv_real.json $O/o_bench_c2.json; grep -E "passed|failed" $O/o_pytest_large.log; tail -n 2 $O/o_bench_*.err
