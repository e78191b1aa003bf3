"""GPU: ShardedEventNet with the real CUDA engine.  World size 1 in-process, and world size 2 as two processes under
gloo that share cuda:0 - the sharded run must equal the single-process run bit for bit (streams are independent,
no collective on the data path; detections are gathered on the host)."""
import os
import socket

import numpy as np
import pytest

import async_ev_cnn_b200 as P
from async_ev_cnn_b200.engine import EventNetCuda
from async_ev_cnn_b200.sharding import ShardedEventNet

pytestmark = pytest.mark.gpu
LAYERS = "conv1=3,3,1,4 pool1=2,2 conv2=3,3,4,8 pool2=2,2 conv3=1,1,8,6"
H, W, S, STEPS, B = 32, 48, 5, 10, 25


def _events():
    return P.synthetic_events("uniform", S, STEPS, B, H, W, seed=21, dt_int=(1, 30))


def _single():
    wts = P.xavier_weights(LAYERS, seed=4)
    net = EventNetCuda(H, W, LAYERS, wts, 0.002, 0.1, "SAME", n_streams=S)
    evs = _events()
    out = [net.step([evs[s, t] if (s + t) % 4 else None for s in range(S)]).copy() for t in range(STEPS)]
    net.close()
    return np.stack(out)


def test_world_size_one_is_the_plain_engine():
    wts = P.xavier_weights(LAYERS, seed=4)
    net = ShardedEventNet(H, W, LAYERS, wts, 0.002, 0.1, "SAME", n_streams=S, device=0)
    evs = _events()
    out = np.stack([net.step([evs[s, t] if (s + t) % 4 else None for s in range(S)]).copy() for t in range(STEPS)])
    net.close()
    assert np.array_equal(out, _single())


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wts = P.xavier_weights(LAYERS, seed=4)
        net = ShardedEventNet(H, W, LAYERS, wts, 0.002, 0.1, "SAME", n_streams=S, device=0)
        evs = _events()
        outs = []
        for t in range(STEPS):
            got = net.step([evs[s, t] if (s + t) % 4 else None for s in range(S)])
            if rank == 0:
                outs.append(np.array(got))
            else:
                assert got is None
        if rank == 0:
            q.put(np.stack(outs))
        net.close()
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_equal_one_process():
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
        p.join(180)
        assert p.exitcode == 0
    assert np.array_equal(got, _single())
