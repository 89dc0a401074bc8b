#!/bin/bash
# GPU call H (1 GPU): parity, then c2 A/B of fd_smem / abc_overlap
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/h_pytest.log 2>&1
for v in "0 0" "1 0" "0 1" "1 1"; do set -- $v
  timeout 300 python bench.py --no-cpu --no-e2e --opt fd_smem=$1 --opt abc_overlap=$2 > $O/h_bench_c2_fd$1_abc$2.json 2> $O/h_bench_c2_fd$1_abc$2.err
done
timeout 300 python bench.py --no-cpu --workload c5 --steps 40 --warmup 5 --no-e2e > $O/h_bench_c5.json 2> $O/h_bench_c5.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 60 --csv --log-file $O/h_launches_c2.csv \
    python bench.py --steps 20 --warmup 10 --no-cpu --no-e2e > $O/h_ncu.log 2>&1
tail -5 $O/h_pytest.log; cat $O/h_bench_c2_*.json $O/h_bench_c5.json; cat $O/h_*.err | tail -5
