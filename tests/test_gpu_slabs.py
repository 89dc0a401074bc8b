"""Slab decomposition on the GPU.
 * two slab engines on ONE device, the test moving the halo planes by hand: covers the per-slab kernels (x edges
   that are neighbours' planes instead of mirrors, re-based node lists) on the single-GPU box;
 * real multi-GPU: one process per GPU, NCCL halo exchange inside the library, overlapped and not
   (skipped unless >= 2 devices are visible); compared with the reference CPU engine's golden traces.
"""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from cases import make_files, make_sim_data
from oracle import Oracle
from pffdtd_b200 import shoebox
from pffdtd_b200.engine import Engine

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _ngpu():
    import torch
    return torch.cuda.device_count()
GOLD = np.load(ROOT / "tests" / "golden" / "traces_ref_cpu_engine.npz")


@pytest.mark.parametrize("name,precision,fuse", (("cart_lossy_mb11", 2, 1), ("cart_lossy_mb11", 1, 0), ("cart_tight", 1, 1), ("fcc2_lossy", 2, 0)))
def test_two_slabs_on_one_device_manual_halo(name, precision, fuse):
    full = make_sim_data(name, precision).sorted()
    ref = Oracle(full).run_all()
    slabs = [full.slab(r, 2) for r in range(2)]
    engs = [Engine(s) for s in slabs]
    try:
        for e in engs:
            e.set_option("manual_halo", 1)
            e.set_option("fuse", fuse)
        for n in range(full.Nt):
            for e in engs:
                e.run_steps(n, 1)
            g = [e.read_grid(1) for e in engs]
            g[0][-1] = g[1][1]    # upper halo of slab 0 <- first owned plane of slab 1
            g[1][0] = g[0][-2]    # lower halo of slab 1 <- last owned plane of slab 0
            for e, a in zip(engs, g):
                e.write_grid(1, a)
        got = np.concatenate([e.read_outputs() for e in engs], axis=0)
    finally:
        for e in engs:
            e.close()
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("name,precision,devices,balance", (("cart_lossy_mb11", 1, (0, 0), True), ("cart_lossy_mb11", 2, (0, 0, 0), False),
                                                             ("cart_tight", 1, (0, 0, 0), True), ("fcc2_lossy", 2, (0, 0), True),
                                                             ("cart_long", 1, (0, 0, 0, 0, 0), True)))
def test_single_process_multi_slab_engine(name, precision, devices, balance):
    """pffdtd_multi_*: one host thread, one engine per slab, edge planes pushed into the neighbours' halos by (peer) copies on a second
    stream, the next step waiting on events -- the reference's own multi-GPU model (gpu_engine.h:994, 1086-1126).  Here the slabs share
    device 0 (the test box has one GPU; with more, `devices` names them); batches of steps, bit-exact against the oracle"""
    from pffdtd_b200.engine import MultiEngine
    sd = make_sim_data(name, precision).sorted()
    ref = Oracle(sd).run_all()
    with MultiEngine(sd, devices=devices, balance=balance) as m:
        assert m.nslabs == len(devices) and sum(m.planes) == sd.Nx and min(m.planes) >= 2
        for n in range(0, sd.Nt, 7):
            m.run_steps(n, min(7, sd.Nt - n))
        got = m.read_outputs()
    assert np.array_equal(got, ref)


def test_run_sim_multi_uses_every_visible_device():
    """nslabs = 0: one slab per visible device (1 on the test box: then it is pffdtd_run_sim)"""
    from pffdtd_b200.engine import run_sim_multi
    sd = make_sim_data("cart_lossy", 2).sorted()
    out, t = run_sim_multi(sd)
    assert np.array_equal(out, Oracle(sd).run_all()) and t > 0
    out3, _ = run_sim_multi(sd, devices=(0, 0, 0))
    assert np.array_equal(out3, out)


def test_single_process_multi_gpu_on_distinct_devices():
    """the same API with one slab per visible GPU (peer copies over NVLink); skipped on a one-GPU box"""
    from pffdtd_b200.engine import MultiEngine
    n = min(_ngpu(), 4)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    for name, precision in (("cart_lossy_mb11", 1), ("fcc2_lossy", 2), ("cart_long", 1)):
        sd = make_sim_data(name, precision).sorted()
        ref = Oracle(sd).run_all()
        with MultiEngine(sd, nslabs=0) as m:
            assert m.nslabs == _ngpu() or m.nslabs == n
            m.run_steps(0, sd.Nt)
            assert np.array_equal(m.read_outputs(), ref), name


def test_multi_refuses_unsorted_lists():
    from pffdtd_b200.engine import MultiEngine, PffdtdError
    sd = make_sim_data("cart_lossy", 2)
    sd.bn_ixyz = sd.bn_ixyz[::-1].copy()
    sd.adj_bn = sd.adj_bn[::-1].copy()
    with pytest.raises(PffdtdError):
        MultiEngine(sd, devices=(0, 0))


WORKER = r'''
import sys, numpy as np
sys.path.insert(0, "{root}")
from pffdtd_b200.sim_fdtd import SimEngine
data_dir, precision, overlap, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
eng = SimEngine(data_dir, precision=precision, quiet=True)
eng.load_h5_data(); eng.allocate_mem()
eng.eng.set_option("overlap", overlap)
eng.run_all(); eng.save_outputs()
if eng.rank == 0:
    np.save(out, eng.sd_full.reorder_output(eng.u_out))
eng.close()
'''


@pytest.mark.parametrize("name,precision,overlap", (("cart_lossy_mb11", 1, 1), ("cart_lossy_mb11", 2, 0), ("cart_tight", 2, 1), ("fcc2_lossy", 1, 1)))
def test_multi_gpu_nccl_halo_exchange(tmp_path, name, precision, overlap):
    world = min(_ngpu(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    shoebox.write_folder(make_files(name), tmp_path / "data")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   PFFDTD_DIST_TIMEOUT="60")
        procs.append(subprocess.Popen([sys.executable, str(script), str(tmp_path / "data"), str(precision), str(overlap), str(tmp_path / "u.npy")],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = []
    try:
        logs = [p.communicate(timeout=int(os.environ.get("PFFDTD_TEST_TIMEOUT", "120")))[0] for p in procs]
    finally:
        for p in procs:  # never leave a rank behind (a hung rank would hold the GPU box until the outer timeout)
            if p.poll() is None:
                p.kill()
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    assert np.array_equal(np.load(tmp_path / "u.npy"), GOLD[f"{name}_p{precision}"])
