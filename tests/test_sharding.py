"""CPU: the N>1 host logic (stream partition + host-side gather of detections) under gloo with
world_size 2 and 3.  The CUDA engine is replaced by the CPU oracle as the per-rank worker, so the
test checks exactly what multi-GPU adds: every stream owned once, no cross-rank data dependence,
detections reassembled in global stream order."""
import os
import socket

import numpy as np
import pytest

import async_ev_cnn_b200 as P
from async_ev_cnn_b200.sharding import SharedHostGather, gather_detections, owner_of, shard_bounds, shard_events, shard_reset_mask

LAYERS = "conv1=3,3,1,4 pool1=2,2 conv2=1,1,4,5"
H, W, STEPS, BATCH = 16, 24, 6, 10


@pytest.mark.parametrize("n,world", [(0, 1), (1, 1), (7, 2), (8, 2), (5, 8), (4096, 8), (1023, 4)])
def test_shard_bounds_partition(n, world):
    seen = []
    for r in range(world):
        lo, hi = shard_bounds(n, world, r)
        assert 0 <= lo <= hi <= n
        seen += list(range(lo, hi))
        for s in range(lo, hi):
            assert owner_of(s, n, world) == r
    assert seen == list(range(n))
    sizes = [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(n, world, world)


def test_shard_events_and_mask():
    per = [np.full((i, 3), i, np.int32) for i in range(7)]
    assert [len(e) for e in shard_events(per, 2, 0)] == [0, 1, 2, 3]
    assert [len(e) for e in shard_events(per, 2, 1)] == [4, 5, 6]
    assert shard_reset_mask(True, 7, 2, 1) is True
    assert shard_reset_mask([1, 0, 0, 0, 0, 1, 0], 7, 2, 1).tolist() == [0, 1, 0]


def test_gather_single_process_is_identity():
    a = np.arange(24, dtype=np.float32).reshape(4, 2, 3)
    assert np.array_equal(gather_detections(a, 4), a)
    with pytest.raises(ValueError):
        gather_detections(a, 5)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _heads_for(streams, seed=11):
    """Per-stream detections for global stream ids `streams`, computed with the CPU oracle."""
    from oracle.event_oracle import OracleEventNet
    wts = P.xavier_weights(LAYERS, seed=3)
    out = []
    for s in streams:
        ev = P.synthetic_events("uniform", 1, STEPS, BATCH, H, W, seed=seed + s)[0]
        net = OracleEventNet(H, W, LAYERS, wts, 0.001, 0.1, "SAME")
        for t in range(STEPS):
            head = net.step(ev[t])
        out.append(head)
    return np.stack(out).astype(np.float32) if out else np.zeros((0, H // 2, W // 2, 5), np.float32)


def _worker(rank, world, port, n_streams, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(n_streams, world, rank)
        local = _heads_for(range(lo, hi))
        got = gather_detections(local, n_streams)
        if rank == 0:
            q.put(got)
        else:
            assert got is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_streams", [(2, 5), (3, 4)])
def test_gloo_sharded_run_equals_single_process(world, n_streams):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    want = _heads_for(range(n_streams))
    assert got.shape == want.shape
    assert np.array_equal(got, want)          # same oracle, same streams: sharding must not change a bit


def _shm_worker(rank, world, port, n_streams, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = SharedHostGather(n_streams, (H // 2, W // 2, 5), slots=2, pin=False)
        lo, hi = shard_bounds(n_streams, world, rank)
        for slot, seed in ((0, 11), (1, 40)):                # two steps in flight, one slot each
            g.mine(slot)[...] = _heads_for(range(lo, hi), seed=seed)
        g.complete()
        if rank == 0:
            q.put((np.array(g.assembled(0)), np.array(g.assembled(1))))
        else:
            assert g.assembled(0) is None
        dist.barrier()
        g.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_streams", [(2, 5), (3, 7)])
def test_shared_host_gather_assembles_in_global_stream_order(world, n_streams):
    """The one-box gather of bench.py / ShardedEventNet.step_packed_async: every rank writes its rows of a shared
    host array (on the GPU box: the D2H target of its head), rank 0 reads the assembled detections after a barrier."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_shm_worker, args=(r, world, port, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    a, b = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert np.array_equal(a, _heads_for(range(n_streams), seed=11))
    assert np.array_equal(b, _heads_for(range(n_streams), seed=40))


def test_shared_host_gather_single_process():
    g = SharedHostGather(3, (2, 2), slots=1, pin=False)
    g.mine(0)[...] = 7.0
    g.complete()
    assert g.assembled(0).shape == (3, 2, 2) and float(g.assembled(0).sum()) == 84.0
    g.close()
