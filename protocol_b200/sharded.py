"""Worker-sharded matching across the GPUs of one node.

One process per GPU.  Every rank holds the (small) full worker and ask tables
but evaluates only its contiguous canonical-order range of workers
(pm_cfg.shard_first/shard_count), so the T x W evaluation — the hot path — is
split N ways with no communication.  The only data-path exchange, once per pass:

  * all-gather of the per-worker "first feasible ask" shards        (uint32[W])
  * all-reduce MIN of the per-ask packed (cost<<32 | worker) argmin  (int64[T])
  * all-reduce SUM of the per-ask feasible counts                    (uint32[T])

over NCCL (NVLink 5 / NVSwitch; a few MB, latency-bound), after which every rank
runs the cheap resolution sweep on identical global arrays and therefore holds
the identical assignment.  `exchange` is backend-agnostic so the host logic is
covered on CPU with gloo (tests/test_sharded_gloo.py).
"""
from __future__ import annotations

from . import abi


class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch can alias an engine buffer."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def wrap_device_buffers(eng, n_workers: int, n_asks: int):
    """torch views (no copy) of the engine's exchange buffers: (first_ask i32[W], ask_best i64[T], ask_count i32[T])."""
    import torch

    p, nb = eng.device_buffer(abi.PM_BUF_WORKER_FIRST_ASK)
    assert nb == n_workers * 4
    fa = torch.as_tensor(_DevArray(p, n_workers, "<i4"), device="cuda")
    p, nb = eng.device_buffer(abi.PM_BUF_ASK_BEST)
    assert nb == n_asks * 8
    best = torch.as_tensor(_DevArray(p, n_asks, "<i8"), device="cuda")
    p, nb = eng.device_buffer(abi.PM_BUF_ASK_COUNT)
    assert nb == n_asks * 4
    cnt = torch.as_tensor(_DevArray(p, n_asks, "<i4"), device="cuda")
    return fa, best, cnt


def shard_range(n_workers: int, world: int, rank: int):
    """Equal contiguous ranges of the canonical worker order (the last rank takes the remainder)."""
    per = (n_workers + world - 1) // world
    lo = min(rank * per, n_workers)
    hi = min(lo + per, n_workers)
    return lo, hi


def exchange(first_ask, ask_best, ask_count, lo: int, hi: int, group=None):
    """In-place exchange on torch tensors (CUDA+NCCL or CPU+gloo).

    first_ask: int32[W] with only [lo,hi) valid on entry; ask_best: int64[T]; ask_count: int32[T]."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    W = first_ask.numel()
    per = (W + world - 1) // world
    if per * world == W:
        dist.all_gather_into_tensor(first_ask, first_ask[lo:hi].clone() if first_ask.device.type == "cpu"
                                    else first_ask[lo:hi], group=group)
    else:  # ragged last shard: pad
        pad = torch.full((per,), -1, dtype=first_ask.dtype, device=first_ask.device)
        pad[: hi - lo] = first_ask[lo:hi]
        out = torch.empty(per * world, dtype=first_ask.dtype, device=first_ask.device)
        dist.all_gather_into_tensor(out, pad, group=group)
        first_ask.copy_(out[:W])
    dist.all_reduce(ask_best, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(ask_count, op=dist.ReduceOp.SUM, group=group)


def packed_stride(n_workers: int, n_asks: int, world: int):
    """(per, stride): the library's exchange layout (pm_xchg_pack in csrc/pm_kernels.cuh) — one buffer per rank,
    [ask_best int64[T] | ask_count u32[T] | first_ask of the rank's own range u32[per]], rounded up to 16 bytes."""
    per = (n_workers + world - 1) // world
    return per, ((n_asks * 12 + per * 4) + 15) & ~15


def packed_exchange(first_ask, ask_best, ask_count, lo: int, hi: int, group=None):
    """The exchange as the library does it inside pm_match (pm_comm): pack, ONE all-gather of bytes, fold.  Same
    arguments and result as `exchange`; used on CPU/gloo to pin the layout and the fold rules the CUDA kernels follow."""
    import numpy as np
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    W, T = first_ask.numel(), ask_best.numel()
    per, stride = packed_stride(W, T, world)
    send = np.zeros(stride, dtype=np.uint8)
    send[:T * 8] = ask_best.numpy().view(np.uint8)
    send[T * 8:T * 12] = ask_count.numpy().view(np.uint8)
    mine = np.full(per, 0xFFFFFFFF, dtype=np.uint32)
    mine[:hi - lo] = first_ask.numpy()[lo:hi].view(np.uint32)
    send[T * 12:T * 12 + per * 4] = mine.view(np.uint8)
    recv = torch.empty(stride * world, dtype=torch.uint8)
    dist.all_gather_into_tensor(recv, torch.from_numpy(send), group=group)
    r = recv.numpy().reshape(world, stride)
    best = r[:, :T * 8].copy().view(np.int64).reshape(world, T).min(axis=0)
    cnt = r[:, T * 8:T * 12].copy().view(np.uint32).reshape(world, T).sum(axis=0, dtype=np.uint32)
    first = r[:, T * 12:T * 12 + per * 4].copy().view(np.uint32).reshape(world * per)[:W]
    ask_best.copy_(torch.from_numpy(best))
    ask_count.copy_(torch.from_numpy(cnt.view(np.int32)))
    first_ask.copy_(torch.from_numpy(first.view(np.int32).copy()))


class ShardedMatcher:
    """One rank's view of a sharded pass over an Engine created with this rank's shard range."""

    def __init__(self, engine, n_workers: int, n_asks: int, rank: int, world: int):
        self.eng, self.W, self.T, self.rank, self.world = engine, n_workers, n_asks, rank, world
        self.lo, self.hi = shard_range(n_workers, world, rank)
        self._bufs = None

    def match(self, mode: int = abi.PM_MODE_FIRST_FIT):
        self.eng.match_local(mode)
        if self._bufs is None:
            self._bufs = wrap_device_buffers(self.eng, self.W, self.T)
        if self.world > 1:
            exchange(*self._bufs, self.lo, self.hi)
        self.eng.match_finish(mode)
        return self.eng.fetch()
