#!/bin/bash
# r2 call Z (1 GPU): voxel-grid fill (pffdtd_voxfill_run) and the ray casting at production size (Musikverein FCC h = 0.03 m) against the reference's output
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 420 python -m pytest tests/test_vox.py -x -q -m gpu -s ) > $O/r2z_pytest.log 2>&1
grep -a "vox production\|passed\|failed\|rror" $O/r2z_pytest.log | tail -n 8
