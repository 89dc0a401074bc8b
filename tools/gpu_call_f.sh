#!/bin/bash
# GPU call F (1 GPU): parity after the boundary split + host-step graph, bench A/B
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/f_pytest.log 2>&1
timeout 300 python bench.py --no-cpu > $O/f_bench_c2_split.json 2> $O/f_bench_c2_split.err
timeout 300 python bench.py --no-cpu --split-boundary 0 > $O/f_bench_c2_nosplit.json 2> $O/f_bench_c2_nosplit.err
timeout 300 python bench.py --no-cpu --workload c3s --steps 200 > $O/f_bench_c3s_split.json 2> $O/f_bench_c3s_split.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 60 --csv --log-file $O/f_launches_c2.csv \
    python bench.py --steps 20 --warmup 10 --no-cpu --no-e2e > $O/f_ncu.log 2>&1
tail -5 $O/f_pytest.log; cat $O/f_bench_c2_split.json $O/f_bench_c2_nosplit.json $O/f_bench_c3s_split.json; cat $O/*.err | tail -5
