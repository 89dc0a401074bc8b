#!/bin/bash
# GPU call J (1 GPU): full-size real rooms vs the unmodified reference CPU engine; full suite; FCC capture after the rotation
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 600 python -m pytest tests/test_large_models.py -m gpu -x -q -rs ) > $O/j_pytest_large.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/j_pytest.log 2>&1
timeout 300 python bench.py --workload c3s --steps 200 > $O/j_bench_c3s.json 2> $O/j_bench_c3s.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_air_tma -s 12 -c 1 -f -o $O/j_air_fcc \
    python bench.py --workload c3s --steps 20 --warmup 10 --no-cpu --no-e2e > $O/j_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 60 --csv --log-file $O/j_launches_c3s.csv \
    python bench.py --workload c3s --steps 20 --warmup 10 --no-cpu --no-e2e > $O/j_ncu2.log 2>&1
tail -6 $O/j_pytest_large.log; tail -4 $O/j_pytest.log; cat $O/j_bench_c3s.json
