"""2-rank debug driver: python tools/mgpu_debug.py  (spawns ranks itself)"""
import os, sys, subprocess, socket, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
if os.environ.get("PF_WORKER") != "1":
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    from cases import make_files
    from pffdtd_b200 import shoebox
    import tempfile
    d = tempfile.mkdtemp()
    shoebox.write_folder(make_files(sys.argv[1] if len(sys.argv) > 1 else "cart_lossy_mb11"), d)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   PF_WORKER="1", PF_DATA=d, NCCL_DEBUG=os.environ.get("NCCL_DEBUG", "WARN"))
        procs.append(subprocess.Popen([sys.executable, __file__], env=env))
    t0 = time.time()
    while any(p.poll() is None for p in procs) and time.time() - t0 < float(os.environ.get("PF_TIMEOUT", "90")):
        time.sleep(0.5)
    for p in procs:
        if p.poll() is None:
            print("KILLING hung rank", flush=True); p.kill()
    print("exit codes", [p.returncode for p in procs], "elapsed %.1f" % (time.time() - t0), flush=True)
    sys.exit(0)

import faulthandler
faulthandler.dump_traceback_later(45, exit=False)
sys.path.insert(0, str(ROOT))
rank = int(os.environ["RANK"])
def say(*a): print(f"[rank {rank}]", *a, flush=True)
say("start")
from pffdtd_b200.sim_fdtd import SimEngine
eng = SimEngine(os.environ["PF_DATA"], precision=1, quiet=True)
eng.load_h5_data(); say("loaded")
eng.allocate_mem(); say("engine + comm ready")
eng.eng.set_option("overlap", int(os.environ.get("PF_OVERLAP", "1")))
eng.eng.run_steps(0, 1); eng.eng.sync(); say("1 step ok")
eng.eng.run_steps(1, eng.Nt - 1); eng.eng.sync(); say("all steps ok")
eng._collect(); say("collected", eng.u_out.shape)
eng.close(); say("closed")
