#!/bin/bash
# GPU call I (1 GPU): the reference's own CUDA engine (unmodified, sm_100a) beside ours on the same inputs
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 300 python tests/diag/compare_reference_gpu_engine.py --workload c2 --steps 200 > $O/i_refgpu_c2.json 2> $O/i_refgpu_c2.err
timeout 300 python tests/diag/compare_reference_gpu_engine.py --workload c3s --steps 100 > $O/i_refgpu_c3s.json 2> $O/i_refgpu_c3s.err
timeout 400 python tests/diag/compare_reference_gpu_engine.py --workload c4 --steps 40 > $O/i_refgpu_c4.json 2> $O/i_refgpu_c4.err
cat $O/i_refgpu_*.json; tail -n 5 $O/i_refgpu_*.err
