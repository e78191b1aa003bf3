"""Stream sharding across the GPUs of one box.

Streams are independent units: each has private recurrent state and no cross-stream term
(integration.py:23-26, conv2d.py:58-63, maxpool.py:33-36), and the weights are read-only and
replicated.  So one process per GPU owns a contiguous block of streams, no collective runs on the
data path, and the only exchange is a host-side gather of the per-stream detections
[S_rank, h_cells, w_cells, C+5B] to rank 0 (north_star: "detections are gathered on the host").

Nothing here touches CUDA: the same code runs under `gloo` on CPU (tests/test_sharding.py,
world_size 2) and under `nccl` in bench.py / ShardedEventNet on the GPU box.
"""
import numpy as np


def shard_bounds(n_streams, world_size, rank):
    """Contiguous block [start, stop) of global stream ids owned by `rank`; block sizes differ by at
    most one and every stream is owned exactly once."""
    if n_streams < 0 or world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad shard request: n_streams=%d world_size=%d rank=%d" % (n_streams, world_size, rank))
    base, extra = divmod(n_streams, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def owner_of(stream, n_streams, world_size):
    """Rank owning global stream id `stream` under shard_bounds."""
    if not 0 <= stream < n_streams:
        raise ValueError("stream %d out of range" % stream)
    base, extra = divmod(n_streams, world_size)
    edge = extra * (base + 1)
    return stream // (base + 1) if stream < edge else extra + (stream - edge) // max(base, 1)


def shard_events(per_stream_events, world_size, rank):
    """Slice of a global per-stream event list (index = global stream id) owned by `rank`."""
    lo, hi = shard_bounds(len(per_stream_events), world_size, rank)
    return per_stream_events[lo:hi]


def shard_reset_mask(reset, n_streams, world_size, rank):
    """graph(events, reset)'s reset argument for this rank: bool stays bool, a mask is sliced."""
    if reset is True or reset is False or reset is None:
        return bool(reset)
    lo, hi = shard_bounds(n_streams, world_size, rank)
    return np.asarray(reset, np.uint8)[lo:hi]


def gather_detections(local_heads, n_streams, group=None, dst=0):
    """Gathers every rank's [S_rank, ...] float32 detections to rank `dst` in global stream order.

    Returns the [n_streams, ...] array on `dst` and None elsewhere.  Single-process (no initialised
    process group) returns the input.  The exchange runs on host tensors when the group's backend
    is gloo; under nccl the local block is staged through the rank's current CUDA device.
    """
    import torch
    import torch.distributed as dist

    local_heads = np.ascontiguousarray(local_heads, dtype=np.float32)
    if not (dist.is_available() and dist.is_initialized()):
        if local_heads.shape[0] != n_streams:
            raise ValueError("single process must hold all %d streams, has %d" % (n_streams, local_heads.shape[0]))
        return local_heads
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(n_streams, world, rank)
    if local_heads.shape[0] != hi - lo:
        raise ValueError("rank %d owns %d streams but passed %d" % (rank, hi - lo, local_heads.shape[0]))
    tail = local_heads.shape[1:]
    on_gpu = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    # blocks may differ by one stream: pad to the largest block so one gather moves everything
    width = -(-n_streams // world)
    send = torch.zeros((width,) + tail, dtype=torch.float32, device=dev)
    send[:hi - lo] = torch.from_numpy(local_heads).to(dev)
    if rank == dst:
        recv = [torch.empty_like(send) for _ in range(world)]
        dist.gather(send, recv, dst=dst, group=group)
        out = np.empty((n_streams,) + tail, np.float32)
        for r in range(world):
            a, b = shard_bounds(n_streams, world, r)
            out[a:b] = recv[r][:b - a].cpu().numpy()
        return out
    dist.gather(send, None, dst=dst, group=group)
    return None


class SharedHostGather:
    """Detections of all ranks of ONE box assembled in one host array without a staging copy.

    Rank `dst` creates a POSIX shared-memory segment holding `slots` arrays [n_streams, *tail] float32; every rank
    maps it and owns the rows of its stream block (`mine(slot)`).  On a GPU box each rank page-locks its mapping
    (cudaHostRegister), so its device-to-host copy of the head lands directly in the assembled array over the GPU's
    own PCIe link: "detections are gathered on the host" (north_star) costs no extra pass and no collective.  After
    `complete()` (a barrier of the group) rank `dst` may read `assembled(slot)`.  Under gloo on CPU the same object
    works unpinned (tests/test_sharding.py).  Multi-node groups must use gather_detections instead."""

    def __init__(self, n_streams, tail, slots=2, group=None, dst=0, pin=None):
        import torch.distributed as dist
        from multiprocessing import shared_memory
        self.group, self.dst, self.slots = group, dst, int(slots)
        self.n_streams, self.tail = int(n_streams), tuple(int(t) for t in tail)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lo, self.hi = shard_bounds(self.n_streams, self.world, self.rank)
        nbytes = max(1, self.slots * self.n_streams * int(np.prod(self.tail)) * 4)
        name = [None]
        if self.rank == dst:
            self._shm = shared_memory.SharedMemory(create=True, size=nbytes)
            name[0] = self._shm.name
        if self.world > 1:
            dist.broadcast_object_list(name, src=dst, group=group)
        if self.rank != dst:
            self._shm = shared_memory.SharedMemory(name=name[0])
            try:                                       # the creator unlinks the segment; attached ranks must not
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self._shm._name, "shared_memory")
            except Exception:
                pass
        self.buf = np.ndarray((self.slots, self.n_streams) + self.tail, np.float32, buffer=self._shm.buf)
        self._pinned = False
        if pin is None:
            try:
                import torch
                pin = torch.cuda.is_available()
            except Exception:
                pin = False
        if pin:
            import torch
            rc = torch.cuda.cudart().cudaHostRegister(self.buf.ctypes.data, self.buf.nbytes, 0)
            if int(rc) != 0:
                raise RuntimeError("cudaHostRegister of the shared detection buffer failed: %s" % rc)
            self._pinned = True
        if self.world > 1:
            dist.barrier(group=group)

    def mine(self, slot):
        """This rank's rows of slot `slot` (a writable view: the D2H target of the rank's head)."""
        return self.buf[slot % self.slots, self.lo:self.hi]

    def complete(self):
        """Every rank has finished writing its rows (call after the rank's own copies are done)."""
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier(group=self.group)

    def assembled(self, slot):
        """[n_streams, *tail] view on rank `dst` (None elsewhere); valid after complete()."""
        return self.buf[slot % self.slots] if self.rank == self.dst else None

    def close(self):
        if self._shm is None:
            return
        if self._pinned:
            import torch
            torch.cuda.cudart().cudaHostUnregister(self.buf.ctypes.data)
            self._pinned = False
        self.buf = None
        shm, self._shm = self._shm, None
        shm.close()
        if self.rank == self.dst:
            shm.unlink()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedEventNet:
    """`n_streams` GLOBAL streams spread over the ranks of the current process group, one
    EventNetCuda per rank on GPU `device` (default: LOCAL_RANK).  step() takes the GLOBAL per-stream
    event list on every rank (each rank reads only its block) and returns the gathered
    [n_streams, H, W, C] detections on rank 0, None elsewhere."""

    def __init__(self, height, width, layers, weights, leak, alpha=0.1, padding="SAME", n_streams=1, device=None,
                 max_events_per_step=0, group=None):
        import os

        import torch.distributed as dist
        from .engine import EventNetCuda
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_streams = int(n_streams)
        self.lo, self.hi = shard_bounds(self.n_streams, self.world, self.rank)
        if self.hi == self.lo:
            raise ValueError("rank %d would own no stream: use n_streams >= world_size" % self.rank)
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.net = EventNetCuda(height, width, layers, weights, leak, alpha, padding, n_streams=self.hi - self.lo,
                                device=device, max_events_per_step=max_events_per_step)
        self.gather = None

    def reset(self, reset=True):
        m = shard_reset_mask(reset, self.n_streams, self.world, self.rank)
        if isinstance(m, bool):
            if m:
                self.net.reset()
        elif m.any():
            self.net.reset(m)

    def step(self, per_stream_events):
        if len(per_stream_events) != self.n_streams:
            raise ValueError("expected %d per-stream event arrays, got %d" % (self.n_streams, len(per_stream_events)))
        heads = self.net.step(per_stream_events[self.lo:self.hi])
        return gather_detections(heads, self.n_streams, self.group)

    # -- pipelined form: packed local events in, detections assembled in host shared memory (one box) ------------
    def open_host_gather(self, slots=2):
        """Creates the shared, page-locked detection buffer (collective: every rank must call it)."""
        self.gather = SharedHostGather(self.n_streams, self.net.head_shape, slots=slots, group=self.group)
        self._slot = 0
        return self.gather

    def step_packed_async(self, events, offsets, cuda_stream=None):
        """This rank's packed events (int32 [total,3], int32 [S_rank+1], pinned for real overlap) -> the step is
        enqueued and its head is copied straight into this rank's rows of the next shared slot.  Returns the slot."""
        slot = self._slot
        self.net.step_packed_async(events, offsets, self.gather.mine(slot), cuda_stream=cuda_stream)
        self._slot = (slot + 1) % self.gather.slots
        return slot

    def sync(self, cuda_stream=None):
        """Waits for this rank's enqueued steps and for every other rank's: afterwards rank 0 may read
        `assembled(slot)` = [n_streams, H, W, C] of the steps enqueued so far."""
        self.net.host_sync(cuda_stream)
        self.gather.complete()

    def assembled(self, slot):
        return self.gather.assembled(slot)

    def close(self):
        if getattr(self, "gather", None) is not None:
            self.gather.close()
            self.gather = None
        self.net.close()
