#!/bin/bash
# round-1 GPU call A (1 GPU): parity tests, bench lines of every single-GPU workload, ncu launch lists + captures
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi -L > $O/a_gpu.txt; nproc >> $O/a_gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/a_pytest.log 2>&1
timeout 400 python bench.py > $O/a_bench_c2.json 2> $O/a_bench_c2.err
timeout 400 python bench.py --workload c3s --steps 200 > $O/a_bench_c3s.json 2> $O/a_bench_c3s.err
timeout 500 python bench.py --workload c4 --steps 100 --warmup 5 > $O/a_bench_c4.json 2> $O/a_bench_c4.err
timeout 600 python bench.py --workload c5 --steps 40 --warmup 5 --no-cpu > $O/a_bench_c5_1gpu.json 2> $O/a_bench_c5_1gpu.err
# launch lists (cold-cache serialised times: compare shares)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 60 --csv --log-file $O/a_launches_c3s.csv \
    python bench.py --workload c3s --steps 20 --warmup 10 --no-cpu --no-e2e > $O/a_ncu1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 60 --csv --log-file $O/a_launches_c2.csv \
    python bench.py --steps 20 --warmup 10 --no-cpu --no-e2e > $O/a_ncu2.log 2>&1
# full captures: FCC tile kernel (c3s), boundary kernels (c2)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_air_tma -s 12 -c 1 -f -o $O/a_air_fcc \
    python bench.py --workload c3s --steps 20 --warmup 10 --no-cpu --no-e2e > $O/a_ncu3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:k_fd|k_rigid|k_abc_faces" -s 30 -c 3 -f -o $O/a_boundary_c2 \
    python bench.py --steps 20 --warmup 10 --no-cpu --no-e2e > $O/a_ncu4.log 2>&1
tail -3 $O/a_pytest.log; cat $O/a_bench_c2.json $O/a_bench_c3s.json $O/a_bench_c4.json $O/a_bench_c5_1gpu.json
