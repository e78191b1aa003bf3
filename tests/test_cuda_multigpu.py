"""GPU (needs >= 2 devices, skipped otherwise): the multi-GPU product path on hardware.  One process per GPU under NCCL,
streams sharded by ShardedEventNet, no collective on the data path; detections reach rank 0 (a) through
gather_detections (one NCCL gather, staged through the device) and (b) through the shared page-locked host array
of the pipelined form (every rank's GPU copies its head straight into its rows) - both must equal the
single-process run bit for bit."""
import os
import socket

import numpy as np
import pytest

import async_ev_cnn_b200 as P
from async_ev_cnn_b200.engine import EventNetCuda, pack_events
from async_ev_cnn_b200.sharding import ShardedEventNet, shard_bounds

pytestmark = pytest.mark.gpu
LAYERS = "conv1=3,3,1,16 pool1=2,2 conv2=3,3,16,32 pool2=2,2 conv3=1,1,32,6"
H, W, S, STEPS, B = 32, 48, 7, 8, 30


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _events():
    return P.synthetic_events("uniform", S, STEPS, B, H, W, seed=33, dt_int=(1, 30))


def _single():
    wts = P.xavier_weights(LAYERS, seed=6)
    net = EventNetCuda(H, W, LAYERS, wts, 0.002, 0.1, "SAME", n_streams=S)
    evs = _events()
    out = [net.step([evs[s, t] if (s + t) % 4 else None for s in range(S)]).copy() for t in range(STEPS)]
    net.close()
    return np.stack(out)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["LOCAL_RANK"] = str(rank)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        wts = P.xavier_weights(LAYERS, seed=6)
        net = ShardedEventNet(H, W, LAYERS, wts, 0.002, 0.1, "SAME", n_streams=S, device=rank)
        evs = _events()
        lo, hi = shard_bounds(S, world, rank)
        gathered = []
        for t in range(STEPS // 2):                       # (a) blocking steps + NCCL gather of the detections
            got = net.step([evs[s, t] if (s + t) % 4 else None for s in range(S)])
            if rank == 0:
                gathered.append(np.array(got))
            else:
                assert got is None
        net.open_host_gather(slots=2)                     # (b) pipelined steps into the shared page-locked host array
        assembled = []
        for t in range(STEPS // 2, STEPS):
            ev, off = pack_events([evs[s, t] if (s + t) % 4 else None for s in range(lo, hi)])
            slot = net.step_packed_async(ev, off)
            net.sync()
            if rank == 0:
                assembled.append(np.array(net.assembled(slot)))
        if rank == 0:
            q.put(np.stack(gathered + assembled))
        dist.barrier()
        net.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least 2 GPUs")
def test_two_gpus_under_nccl_equal_one_process():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert np.array_equal(got, _single())
