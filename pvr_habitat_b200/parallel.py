"""Data-parallel plumbing (one process per GPU, torch.distributed / NCCL). The reference has no distributed code
(SURVEY.md D7); the parity target is the single-process reference at the GLOBAL batch:

* embedding pass: observations are cut into contiguous blocks per rank, no collective;
* BC training: every rank draws the SAME seeded `sample_with_minimum_distance` and keeps the sequences
  `starting_i[rank*B/G:(rank+1)*B/G]`; the loss is scaled by 1/(T*B_global) and gradients are SUM all-reduced, so the
  update equals the reference's mean over the global batch; BatchNorm1d statistics are computed from all-reduced
  per-feature sums (synchronised batch statistics, global count).
"""
import ctypes
import os
import sys

import torch


def env_world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process if unset)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of n items owned by `rank` (blocks differ by at most one item)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_starts(starting_i, rank, world):
    """Sequences of the global BC batch trained by `rank` (B must be divisible by the world size so that every rank
    runs the same shapes)."""
    b = len(starting_i)
    if b % world:
        raise ValueError(f"batch_size {b} is not divisible by the world size {world}")
    per = b // world
    return list(starting_i[rank * per:(rank + 1) * per])


def resolve_group(process_group=None):
    """The process group the data-parallel collectives run on, or None for a single process. `None` under an
    initialised torch.distributed with more than one rank — the normal torchrun idiom — means the default (WORLD)
    group, never "no collectives"."""
    if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        return None
    group = process_group if process_group is not None else torch.distributed.group.WORLD
    return group if torch.distributed.get_world_size(group) > 1 else None


class Comm:
    """The collectives of the data-parallel BC path over one process group.

    NCCL groups: a communicator of our own created through the C ABI (`pvr_comm_*`, csrc/comm.cu; the unique id travels
    over the torch.distributed group once). Every collective is a plain launch on ONE communication stream — forked
    from / joined to the caller's stream with events, so it can overlap the kernels issued after it and is captured
    into the whole-step CUDA graph like any kernel. Other backends (gloo: the CPU-side tests, two ranks sharing one
    GPU) go through torch.distributed on the caller's stream.
    """
    _DT = {torch.float32: 0, torch.float64: 1, torch.bfloat16: 2, torch.int64: 3}

    def __init__(self, group):
        from . import _lib
        self.group = group
        self.rank = torch.distributed.get_rank(group)
        self.world = torch.distributed.get_world_size(group)
        self.native = None
        self._stream = None
        self._pending = False
        if torch.distributed.get_backend(group) == "nccl" and torch.cuda.is_available():
            lib = _lib.lib()
            path = next((os.path.join(p, "nvidia", "nccl", "lib", "libnccl.so.2") for p in sys.path
                         if os.path.exists(os.path.join(p, "nvidia", "nccl", "lib", "libnccl.so.2"))), "")
            _lib.check(lib.pvr_comm_load(path.encode()), "pvr_comm_load")
            uid = (ctypes.c_char * 128)()
            if self.rank == 0:
                _lib.check(lib.pvr_comm_unique_id(uid), "pvr_comm_unique_id")
            box = [bytes(uid)]
            torch.distributed.broadcast_object_list(box, src=torch.distributed.get_global_rank(group, 0), group=group)
            handle = ctypes.c_void_p()
            _lib.check(lib.pvr_comm_init(self.rank, self.world, box[0], ctypes.byref(handle)), "pvr_comm_init")
            self.native, self._lib = handle, lib
            self._stream = torch.cuda.Stream()

    @property
    def capturable(self):
        return self.native is not None

    def all_reduce(self, t, wait=True):
        """In-place SUM over the ranks. wait=False: the caller's stream does not wait for the result — call `join()`
        before using it (gradient buckets all-reduced while the backward continues)."""
        if self.native is None:
            torch.distributed.all_reduce(t, group=self.group)
            return t
        from . import _lib
        assert t.is_contiguous() and t.is_cuda
        cur = torch.cuda.current_stream(t.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        self._stream.wait_event(ev)
        _lib.check(self._lib.pvr_comm_allreduce(self.native, t.data_ptr(), t.numel(), self._DT[t.dtype],
                                                ctypes.c_void_p(self._stream.cuda_stream)), "pvr_comm_allreduce")
        self._pending = True
        if wait:
            self.join()
        return t

    def join(self):
        """Make the caller's current stream wait for every collective issued so far."""
        if self.native is not None and self._pending:
            ev = torch.cuda.Event()
            ev.record(self._stream)
            torch.cuda.current_stream().wait_event(ev)
            self._pending = False

    def broadcast(self, t, root=0):
        if self.native is None:
            torch.distributed.broadcast(t, src=torch.distributed.get_global_rank(self.group, root), group=self.group)
            return t
        from . import _lib
        cur = torch.cuda.current_stream(t.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        self._stream.wait_event(ev)
        _lib.check(self._lib.pvr_comm_broadcast(self.native, t.data_ptr(), t.numel(), self._DT[t.dtype], root,
                                                ctypes.c_void_p(self._stream.cuda_stream)), "pvr_comm_broadcast")
        self._pending = True
        self.join()
        return t


_COMMS = {}


def comm_for(group):
    """One Comm per process group (NCCL communicators are expensive and must be created collectively)."""
    key = id(group)
    if key not in _COMMS:
        _COMMS[key] = Comm(group)
    return _COMMS[key]


def attach(policy, process_group, global_rows, broadcast=True):
    """Make `policy` (pvr_habitat_b200.models.PolicyNet) synchronise BatchNorm sums and gradients over the group.
    Parameters and buffers are broadcast from the group's first rank, so the replicas start identical whatever the
    seeds of the ranks were."""
    group = resolve_group(process_group)
    policy.process_group = group
    policy.comm = comm_for(group) if group is not None else None
    policy.global_rows = global_rows if group is not None else None
    if group is not None and broadcast:
        with torch.no_grad():
            for t in list(policy.parameters()) + list(policy.buffers()):
                if t.is_floating_point() or t.dtype == torch.int64:
                    policy.comm.broadcast(t.data, 0)
    return policy


def allreduce_sum_(t, group=None):
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(t, group=group)
    return t
