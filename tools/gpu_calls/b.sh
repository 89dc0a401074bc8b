#!/bin/bash
# GPU call B (1 GPU): energy tests + full gpu suite
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 900 python -m pytest tests/test_energy.py -m gpu -x -q ) > $O/b_pytest_energy.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/b_pytest.log 2>&1
tail -15 $O/b_pytest_energy.log; tail -5 $O/b_pytest.log
