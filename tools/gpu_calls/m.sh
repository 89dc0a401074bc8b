#!/bin/bash
# GPU call M (1 GPU): bench lines and launch lists of the REAL rooms at full size (data_large/)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 200 python bench.py --workload ctk_real --steps 200 > $O/m_bench_ctk_real.json 2> $O/m_bench_ctk_real.err
timeout 200 python bench.py --workload mv_real --steps 150 > $O/m_bench_mv_real.json 2> $O/m_bench_mv_real.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 60 --csv --log-file $O/m_launches_ctk_real.csv \
    python bench.py --workload ctk_real --steps 20 --warmup 10 --no-cpu --no-e2e > $O/m_ncu1.log 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 66 --csv --log-file $O/m_launches_mv_real.csv \
    python bench.py --workload mv_real --steps 20 --warmup 10 --no-cpu --no-e2e > $O/m_ncu2.log 2>&1
cat $O/m_bench_ctk_real.json $O/m_bench_mv_real.json; tail -n 3 $O/m_bench_*.err
