"""Host-side plumbing for the slab decomposition: one process per GPU (torchrun), `torch.distributed` only for
the bootstrap (sharing the NCCL unique id) and for collecting the receiver rows at the end.  The per-step
halo exchange itself happens inside libpffdtd_b200.so (ncclSend/ncclRecv on the engine's streams), replacing the
reference's single-process cudaMemcpyPeerAsync waves (c_cuda/gpu_engine.h:1086-1126)."""
from __future__ import annotations

import os

import numpy as np


def dist_env():
    """(rank, world_size, local_rank) from the torchrun environment, (0, 1, 0) when run plainly"""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend="gloo"):
    """join the default process group if WORLD_SIZE > 1 and nobody did yet; returns torch.distributed or None"""
    rank, world, _ = dist_env()
    if world == 1:
        return None
    import datetime
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group(backend, rank=rank, world_size=world,
                                timeout=datetime.timedelta(seconds=int(os.environ.get("PFFDTD_DIST_TIMEOUT", "600"))))
    return dist


def broadcast_bytes(payload, src=0):
    """`payload` (bytes) from rank `src` to everyone"""
    dist = init()
    if dist is None:
        return payload
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def gather_rows(u: np.ndarray) -> np.ndarray:
    """concatenate each rank's receiver rows in rank order (== the sorted receiver order of the whole grid,
    because slabs are contiguous in x and the lists are sorted, gpu_engine.h:562-661)"""
    dist = init()
    if dist is None:
        return u
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, u)
    return np.concatenate(parts, axis=0)


def sum_arrays(a: np.ndarray) -> np.ndarray:
    """element-wise sum over ranks (per-rank partial sums of the energy balance), identical on every rank"""
    dist = init()
    if dist is None:
        return a
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, np.asarray(a, np.float64))
    out = np.zeros_like(parts[0])
    for p in parts:  # fixed rank order: the same bits on every rank
        out = out + p
    return out


def exchange_planes(send_lo, send_hi, rank, world):
    """CPU stand-in of the halo exchange, used by the gloo tests with the oracle as the per-rank engine:
    send plane 1 down / plane Nx-2 up, receive the neighbours' into plane Nx-1 / 0.  Returns (from_lo, from_hi)."""
    import torch
    dist = init()
    from_lo = from_hi = None
    reqs = []
    if rank > 0:
        t = torch.from_numpy(np.ascontiguousarray(send_lo))
        r = torch.empty_like(t)
        reqs += [dist.isend(t, rank - 1), dist.irecv(r, rank - 1)]
        from_lo = r
    if rank < world - 1:
        t2 = torch.from_numpy(np.ascontiguousarray(send_hi))
        r2 = torch.empty_like(t2)
        reqs += [dist.isend(t2, rank + 1), dist.irecv(r2, rank + 1)]
        from_hi = r2
    for q in reqs:
        q.wait()
    return (None if from_lo is None else from_lo.numpy()), (None if from_hi is None else from_hi.numpy())


def connect_peers(eng, rank, world):
    """halo exchange over peer memory (pffdtd_peer_export / _connect) between the ranks of one node: every rank publishes its
    blob, takes its neighbours'.  Returns True when connected; on any failure (GPUs without peer access, another node) every rank
    keeps the NCCL exchange -- the decision is collective, so that neighbours never disagree."""
    dist = init()
    if dist is None or world == 1 or os.environ.get("PFFDTD_P2P", "1") == "0":
        return False
    try:
        blob = eng.peer_export()
    except Exception:  # noqa: BLE001
        blob = None
    blobs = [None] * world
    dist.all_gather_object(blobs, blob)
    ok = all(b is not None for b in blobs)
    if ok:
        try:
            eng.peer_connect(blobs[rank - 1] if rank > 0 else None, blobs[rank + 1] if rank < world - 1 else None)
        except Exception:  # noqa: BLE001
            ok = False
    oks = [None] * world
    dist.all_gather_object(oks, ok)
    if not all(oks):
        if ok:
            eng.set_option("p2p", 0)
        return False
    return True
