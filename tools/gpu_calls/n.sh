#!/bin/bash
# GPU call N (1 GPU): parity suite without the two long full-size tests
cd "$GRAFT_REPO_ROOT" || exit 1
( time timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py tests/test_models.py -m gpu -q -x ) > gpurun_out/n_pytest.log 2>&1
grep -E "passed|failed|Error|assert|real" gpurun_out/n_pytest.log | head -20
