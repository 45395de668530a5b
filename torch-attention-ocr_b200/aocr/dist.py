"""Data-parallel plumbing for libaocr: one process per GPU.  Two flavours of the same exchange:
  * attach_native (default on GPUs): the library issues the NCCL all-reduces itself (include/aocr.h: aocr_dp_init);
    torch.distributed only bootstraps the NCCL unique id;
  * attach / GradSync: torch.distributed does the collectives, the engine says WHEN through its exchange hook
    (aocr_set_allreduce) — the flavour the CPU (gloo) tests exercise, and the seam for a custom exchange.

The reference is single-device; training shards the batch over ranks.  To reproduce the
single-device reference at the GLOBAL batch the engine (i) scales the loss by 1/global_batch, (ii) sums the
batch-norm statistics of the 3 BN layers over ranks (kind 0, ordered on the engine stream) and (iii) sums the
flat gradient buffer in 3 buckets [proj|decoder], [enc_fw|enc_bw], [cnn] in the order backward completes them
(kind 1: issued on a side stream so the all-reduce of a finished bucket overlaps the rest of backward), then
joins (kind 2) before the identical clip+SGD on every rank.  Greedy decode shards batches with no communication.
"""
import ctypes as C

ALLREDUCE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int64, C.c_int)


class ExchangeLog:
    """what the engine asked for during one step (tests / profiling)"""

    def __init__(self):
        self.calls = []


class GradSync:
    """Implements the exchange hook with torch.distributed.  `wrap(ptr, n)` turns a device pointer into a
    tensor (CUDA: zero-copy through __cuda_array_interface__; the CPU/gloo test passes its own wrapper)."""

    def __init__(self, wrap, engine_stream=None, comm_stream=None, group=None, overlap=True):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.wrap, self.group = wrap, group
        self.engine_stream, self.comm_stream = engine_stream, comm_stream
        self.overlap = overlap and comm_stream is not None
        self.pending = False
        self.log = ExchangeLog()
        self.cb = ALLREDUCE_FN(self._hook)      # keep a reference: ctypes callbacks must outlive their use

    def _hook(self, user, ptr, n, kind):
        torch, dist = self.torch, self.dist
        self.log.calls.append((int(kind), int(n)))
        if kind == 2:                            # join
            if self.pending and self.engine_stream is not None:
                self.engine_stream.wait_stream(self.comm_stream)
            self.pending = False
            return
        t = self.wrap(ptr, int(n))
        if self.engine_stream is None:           # CPU (gloo) flavour used by the host-logic tests
            dist.all_reduce(t, group=self.group)
            return
        if kind == 1 and self.overlap:
            self.comm_stream.wait_stream(self.engine_stream)   # bucket is complete on the engine stream
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(t, group=self.group)
            self.pending = True
        else:
            with torch.cuda.stream(self.engine_stream):
                dist.all_reduce(t, group=self.group)


def cuda_wrap(torch):
    def wrap(ptr, n):
        class _W:
            pass
        w = _W()
        w.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}
        return torch.as_tensor(w, device="cuda")
    return wrap


def attach(handle, device_index, group=None, overlap=True):
    """Install the NCCL exchange hook on an aocr Handle created with dp_world > 1."""
    import torch
    es = torch.cuda.ExternalStream(handle.stream(), device=device_index)
    cs = torch.cuda.Stream(device=device_index)
    gs = GradSync(cuda_wrap(torch), engine_stream=es, comm_stream=cs, group=group, overlap=overlap)
    handle.set_allreduce(gs.cb)
    handle._grad_sync = gs                        # lifetime: as long as the handle
    return gs


def attach_native(handle, device_index, group=None):
    """Native exchange: the library itself issues the NCCL all-reduces (no Python in the step, data-parallel steps are
    graph-captured).  torch.distributed only ships rank 0's NCCL unique id to the other ranks of `group`."""
    import torch
    import torch.distributed as dist
    from .capi import Lib
    rank = dist.get_rank(group)
    uid = Lib.get().dp_unique_id() if rank == 0 else bytes(128)
    dev = torch.device("cuda", device_index) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor(list(uid), dtype=torch.uint8, device=dev)
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast(t, src=src, group=group)
    handle.dp_init(bytes(t.cpu().tolist()))
    return handle


def shard(batch, rank, world):
    """rank r takes rows [r*b/world, (r+1)*b/world) of every batch tensor (all rows share one width)."""
    images, targets, targets_eval = batch[0], batch[1], batch[2]
    b = images.shape[0]
    assert b % world == 0, "global batch must divide evenly over the ranks"
    lo, hi = rank * b // world, (rank + 1) * b // world
    te = targets_eval[lo:hi]
    return [images[lo:hi], targets[lo:hi], te, int((te != 1).sum()), batch[4][lo:hi] if batch[4] else None]
