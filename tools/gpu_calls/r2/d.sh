#!/bin/bash
# r2 call D (1 GPU): FCC packed arithmetic parity + lines; where does the service-warp kernel lose time on c2 (idle-service run, ncu source-level capture)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_zz_obstacles.py tests/test_models.py -x -q -m gpu ) > $O/r2d_pytest.log 2>&1
tail -n 6 $O/r2d_pytest.log
b() { name=$1; shift; timeout 300 python bench.py --no-cpu --no-also --no-parity "$@" > $O/r2d_$name.json 2> $O/r2d_$name.err; python - <<PY
import json
try:
    d=json.load(open("$O/r2d_$name.json")); r=d["roofline"]
    print("$name", "value %.1f e2e %.1f ms %.4f air_ms %.4f air_frac %.3f whole %.3f launches %d" % (d["value"], d.get("e2e",{}).get("value",0), d["ms_per_step"], r["air_ms_per_step"], r["frac"], r["whole_step_frac"], d["gpu_launches"]))
except Exception as ex:
    print("$name failed", ex); print(open("$O/r2d_$name.err").read()[-800:])
PY
}
b c3s --workload c3s --steps 100
b mvbig --workload mv_big --steps 40
b mvreal --workload mv_real --steps 150
b c2_svc0 --workload c2 --steps 200 --opt svc=0
b c2_svc0_fd0 --workload c2 --steps 200 --opt svc=0 --opt fd_fixed=0
b c2_cfg12_idle --workload c2 --steps 200 --opt svc=0 --air-cfg 12
b c2_svc1 --workload c2 --steps 200
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_air_tma -s 12 -c 1 -o $O/r2d_air_svc_c2 -f python bench.py --workload c2 --steps 10 --warmup 4 --no-cpu --no-also --no-parity --no-e2e > $O/r2d_ncu.log 2>&1
tail -n 3 $O/r2d_ncu.log; ls -la $O/r2d_air_svc_c2.ncu-rep
python -c "import h5py; print('h5py', h5py.__version__)" 2>&1 | tail -1
