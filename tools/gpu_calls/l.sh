#!/bin/bash
# GPU call L (1 GPU): what the driver runs at round end -- gpu suite, smoke, default bench, reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > $O/l_pytest.log 2>&1
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/l_smoke.log 2>&1
( time timeout 600 python bench.py ) > $O/l_bench_c2.json 2> $O/l_bench_c2.err
tail -4 $O/l_pytest.log; tail -4 $O/l_smoke.log; cat $O/l_bench_c2.json; tail -4 $O/l_bench_c2.err
