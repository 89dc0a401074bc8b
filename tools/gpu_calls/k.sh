#!/bin/bash
# GPU call K (1 GPU): full-size real rooms vs the unmodified reference CPU engine; default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 600 python -m pytest tests/test_large_models.py -m gpu -q -rs ) > $O/k_pytest_large.log 2>&1
( time timeout 600 python bench.py ) > $O/k_bench_c2.json 2> $O/k_bench_c2.err
tail -6 $O/k_pytest_large.log; cat $O/k_bench_c2.json; tail -4 $O/k_bench_c2.err
