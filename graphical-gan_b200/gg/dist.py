"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch) as transport.

The path shards by minibatch: parameters and both Adam states are replicated, each rank runs the
step on its B/P slice of the batch and of the injected noise, and there is exactly ONE exchange per optimiser step —
a summing all-reduce of the flat gradient bucket (≈12.5 MB for G+E, ≈16.3 MB for D on gmgan-CIFAR) — plus 2·C-float
all-reduces inside batch norm so that batch statistics equal the un-sharded reference's (SyncBN).  The reference has
no multi-GPU code at all; this module is new functionality, and it is the only place a collective is issued.
"""
import os

_state = {"initialized": False}


def _td():
    import torch.distributed as td
    return td


def world_size():
    td = _td()
    if td.is_available() and td.is_initialized():
        return td.get_world_size()
    return 1


def rank():
    td = _td()
    if td.is_available() and td.is_initialized():
        return td.get_rank()
    return 0


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
    import torch
    td = _td()
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws <= 1 or td.is_initialized():
        return rank(), world_size()
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    td.init_process_group(backend=backend)
    if backend == "nccl":
        # warm the communicator up outside any CUDA-graph capture
        t = torch.zeros(1, device="cuda")
        td.all_reduce(t)
        torch.cuda.synchronize()
    import atexit
    atexit.register(shutdown)
    return td.get_rank(), td.get_world_size()


def all_reduce_sum(t):
    """In-place sum over ranks on the current stream (capturable in a CUDA graph with NCCL)."""
    if world_size() > 1:
        _td().all_reduce(t)
    return t


class SmallAllReduce(object):
    """One-kernel all-reduce of <= max_floats fp32 values through NVLink peer memory (csrc/gg_comm.cu).  Used for the
    SyncBN statistic exchange: it is CUDA-graph capturable, so a data-parallel step is no longer cut into a graph segment
    per batch-norm layer.  Single node only (CUDA IPC)."""
    MAX_FLOATS = 16384

    def __init__(self):
        import ctypes as C
        import torch
        from . import cabi
        self.cabi, self.C = cabi, C
        self.rank, self.world = rank(), world_size()
        buf = C.c_void_p()
        handle = C.create_string_buffer(64)
        cabi.call("gg_comm_alloc", self.MAX_FLOATS, C.byref(buf), handle)
        handles = [None] * self.world
        _td().all_gather_object(handles, bytes(handle.raw))
        ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                ptrs.append(buf.value)
            else:
                p = C.c_void_p()
                cabi.call("gg_comm_open", C.create_string_buffer(h, 64), C.byref(p))
                ptrs.append(p.value)
        self.peers = (C.c_void_p * self.world)(*ptrs)
        self.epoch = torch.zeros(1, dtype=torch.int32, device="cuda")
        _td().barrier()

    def __call__(self, src, dst, n, stream):
        self.cabi.call("gg_allreduce_small", src.data_ptr(), dst.data_ptr(), int(n), self.peers, self.rank, self.world,
                       self.MAX_FLOATS, self.epoch.data_ptr(), stream)


class PeerArena(object):
    """One CUDA-IPC exchange arena per rank, mapped by every peer over NVLink (csrc/gg_comm.cu::gg_comm_alloc_bytes).  The
    one-launch SyncBN kernels (gg_bn_*_fused_dp) exchange their per-channel sums through it; every batch-norm call site
    bump-allocates its own region here — in graph-construction order, i.e. at the same offset on every rank."""
    BYTES = int(os.environ.get("GG_ARENA_MB", "64")) << 20

    def __init__(self):
        import ctypes as C
        from . import cabi
        self.rank, self.world = rank(), world_size()
        buf = C.c_void_p()
        handle = C.create_string_buffer(64)
        cabi.call("gg_comm_alloc_bytes", self.BYTES, C.byref(buf), handle)
        handles = [None] * self.world
        _td().all_gather_object(handles, bytes(handle.raw))
        ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                ptrs.append(buf.value)
            else:
                q = C.c_void_p()
                cabi.call("gg_comm_open", C.create_string_buffer(h, 64), C.byref(q))
                ptrs.append(q.value)
        self.peers = (C.c_void_p * self.world)(*ptrs)
        self.off = 0
        _td().barrier()

    def alloc(self, nbytes):
        off = self.off
        self.off += (int(nbytes) + 127) & ~127
        if self.off > self.BYTES:
            raise MemoryError("peer exchange arena exhausted (%d bytes): raise GG_ARENA_MB" % self.BYTES)
        return off


_arena = {}


def peer_arena():
    """process-wide PeerArena, or None (single rank, or GG_BN_DP=0 -> the multi-kernel SyncBN form)"""
    if world_size() <= 1 or os.environ.get("GG_BN_DP", "1") == "0":
        return None
    if "obj" not in _arena:
        _arena["obj"] = PeerArena()
    return _arena["obj"]


def shutdown():
    """Leave the process group BEFORE the interpreter tears down CUDA graphs that hold captured NCCL kernels (an exit with the
    communicator alive inside captured graphs hung round 1's dp_check): drop the plans, synchronise, destroy the group."""
    import gc
    td = _td()
    if not (td.is_available() and td.is_initialized()):
        return
    try:
        import torch
        from . import executor
        for plan in executor.ALL_PLANS:      # module-level references to a Plan must not keep its graphs alive
            plan.graph = None
            plan.keep = []
        del executor.ALL_PLANS[:]
        executor.RT.plans.clear()
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        td.destroy_process_group()
    except Exception:
        pass


_small = {}


def small_all_reduce():
    """process-wide SmallAllReduce, or None (single rank, or GG_SMALL_ALLREDUCE=0 -> NCCL for everything)"""
    if world_size() <= 1 or os.environ.get("GG_SMALL_ALLREDUCE", "1") == "0":
        return None
    if "obj" not in _small:
        _small["obj"] = SmallAllReduce()
    return _small["obj"]


def shard_bounds(n, r=None, w=None):
    """[lo, hi) of the contiguous dim-0 slice of an n-row batch owned by rank r of w (n must divide evenly:
    the mean-reduced losses only average correctly over equal shards)."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    if n % w != 0:
        raise ValueError("global batch %d is not divisible by world size %d" % (n, w))
    per = n // w
    return r * per, (r + 1) * per


def shard(array, r=None, w=None):
    lo, hi = shard_bounds(array.shape[0], r, w)
    return array[lo:hi]


def bucket_offsets(sizes):
    """element offsets of each gradient inside the flat all-reduce bucket, and the bucket length"""
    offs, acc = [], 0
    for s in sizes:
        offs.append(acc)
        acc += int(s)
    return offs, acc
