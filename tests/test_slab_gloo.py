"""The N>1 path on CPU: world_size-2 and -3 `gloo` process groups, one slab per rank (SimData.slab), the CPU
oracle as the per-rank engine and torch.distributed send/recv as the halo exchange.  The concatenated receiver
rows must equal the single-domain run bit for bit -- the slab split does not change any node's arithmetic."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent

WORKER = r'''
import sys, numpy as np
sys.path.insert(0, "{root}"); sys.path.insert(0, "{root}/tests")
from cases import make_sim_data
from oracle import Oracle
from pffdtd_b200 import parallel
name, precision, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rank, world, _ = parallel.dist_env()
parallel.init("gloo")
full = make_sim_data(name, precision).sorted()
sd = full.slab(rank, world)
o = Oracle(sd)
for n in range(sd.Nt):
    o.run_steps(n, 1)
    lo, hi = parallel.exchange_planes(o.read_plane(1), o.read_plane(sd.Nx - 2), rank, world)
    if lo is not None: o.write_plane(0, lo)
    if hi is not None: o.write_plane(sd.Nx - 1, hi)
u = parallel.gather_rows(o.u_out[:, :sd.Nt])
payload = parallel.broadcast_bytes(b"id-from-rank-0" if rank == 0 else None)
assert payload == b"id-from-rank-0"
tot = parallel.sum_arrays(np.arange(5.0) * (rank + 1))   # what adds the per-slab energy sums
assert np.array_equal(tot, np.arange(5.0) * (world * (world + 1) // 2))
if rank == 0:
    np.save(out, full.reorder_output(full.rescale_output(u)))
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("name,precision,world", (("cart_lossy_mb11", 2, 2), ("cart_lossy_mb11", 1, 3), ("fcc2_lossy", 2, 2), ("cart_tight", 1, 2)))
def test_slabs_over_gloo_equal_the_single_domain(tmp_path, name, precision, world):
    gold = np.load(ROOT / "tests" / "golden" / "traces_ref_cpu_engine.npz")[f"{name}_p{precision}"]
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    out = tmp_path / "u.npy"
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script), name, str(precision), str(out)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    logs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    assert np.array_equal(np.load(out), gold)


HOST_WORKER = r'''
import sys, numpy as np
sys.path.insert(0, "{root}"); sys.path.insert(0, "{root}/tests")
from oracle import Oracle
from pffdtd_b200 import parallel, sim_fdtd

class SlabOracleEngine:
    """what SimEngine asks of the CUDA engine, served by the CPU oracle; the halo planes travel over gloo"""
    def __init__(self, sd, device=0):
        self.sd, self.o = sd, Oracle(sd)
    def comm_init(self, uid, rank, world):
        assert uid == b"u" * 128
        self.rank, self.world = rank, world
    def set_option(self, k, v): pass
    def run_steps(self, n0, k):
        for n in range(n0, n0 + k):
            self.o.run_steps(n, 1)
            lo, hi = parallel.exchange_planes(self.o.read_plane(1), self.o.read_plane(self.sd.Nx - 2), self.rank, self.world)
            if lo is not None: self.o.write_plane(0, lo)
            if hi is not None: self.o.write_plane(self.sd.Nx - 1, hi)
    def sync(self): pass
    def read_outputs(self, n0=0, n1=None): return self.o.u_out[:, n0:n1].copy()
    def close(self): pass

sim_fdtd.Engine = SlabOracleEngine
sim_fdtd.comm_unique_id = lambda: b"u" * 128
parallel.init("gloo")
u = sim_fdtd.run_folder(sys.argv[1], precision=int(sys.argv[2]), nsteps=9)
if parallel.dist_env()[0] == 0:
    np.save(sys.argv[3], u)
'''


@pytest.mark.parametrize("name,precision,world", (("cart_lossy_mb11", 1, 2), ("fcc2_lossy", 2, 3)))
def test_host_loop_under_torchrun_env_with_slabs(tmp_path, name, precision, world):
    """SimEngine with WORLD_SIZE > 1: every rank loads the folder, sorts, takes its slab, shares the communicator id, runs, and rank 0
    gathers the receiver rows and writes sim_outs.h5 -- equal to the single-domain reference run"""
    import sys as _sys
    _sys.path.insert(0, str(ROOT / "tests"))
    from cases import make_files
    from pffdtd_b200 import h5lite, shoebox
    gold = np.load(ROOT / "tests" / "golden" / "traces_ref_cpu_engine.npz")[f"{name}_p{precision}"]
    data = tmp_path / "data"
    shoebox.write_folder(make_files(name), data)
    script = tmp_path / "worker.py"
    script.write_text(HOST_WORKER.format(root=ROOT))
    out = tmp_path / "u.npy"
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script), str(data), str(precision), str(out)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    logs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    assert np.array_equal(np.load(out), gold)
    assert np.array_equal(h5lite.File(data / "sim_outs.h5")["u_out"][...], gold)
