#!/bin/bash
# GPU call G (1 GPU): FCC kernel with register rotation: parity, then c3s with three tile configurations
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/g_pytest.log 2>&1
for c in 0 5 7 3; do
  timeout 300 python bench.py --no-cpu --no-e2e --workload c3s --steps 100 --air-cfg $c > $O/g_bench_c3s_cfg$c.json 2> $O/g_bench_c3s_cfg$c.err
done
timeout 300 python bench.py --no-cpu > $O/g_bench_c2.json 2> $O/g_bench_c2.err
tail -5 $O/g_pytest.log; cat $O/g_bench_c3s_cfg*.json $O/g_bench_c2.json; cat $O/g_*.err | tail -5
