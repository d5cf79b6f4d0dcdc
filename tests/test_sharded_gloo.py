"""Host-side logic of the worker-sharded pass (protocol_b200/sharded.py) on CPU: two gloo ranks,
per-shard partial results stand in for pm_match_local (computed here by the oracle), the exchange
must reproduce the single-shard arrays exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pm_oracle as orc
from protocol_b200 import synth
from protocol_b200.sharded import exchange, packed_exchange, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, n_workers, n_asks, out_dir, packed):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = synth.make_workers(n_workers)
    a = synth.make_asks(n_asks, "mixed")
    bits, npat, nmod, words = synth.intern_tables(w, a)
    lo, hi = shard_range(n_workers, world, rank)
    ev = orc.soa_eval_matrix(w.a, w.b, a.asks, a.opts, bits, words, 0, n_asks, lo, hi)
    first = torch.full((n_workers,), -1, dtype=torch.int32)
    first[lo:hi] = torch.from_numpy(ev["col_first"].view(np.int32))
    best = torch.from_numpy(ev["row_best"].copy())
    cnt = torch.from_numpy(ev["row_count"].view(np.int32).copy())
    (packed_exchange if packed else exchange)(first, best, cnt, lo, hi)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), first=first.numpy(), best=best.numpy(), cnt=cnt.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("packed", [False, True], ids=["three_collectives", "one_packed_all_gather"])
@pytest.mark.parametrize("n_workers", [4000, 4001])   # even split and ragged last shard
def test_two_rank_exchange_matches_single_shard(tmp_path, n_workers, packed):
    """packed=True is the layout and fold the library uses inside pm_match (pm_comm, one all-gather per pass)."""
    n_asks, world = 300, 2
    port = _free_port()
    mp.spawn(_rank_main, args=(world, port, n_workers, n_asks, str(tmp_path), packed), nprocs=world, join=True)
    w = synth.make_workers(n_workers)
    a = synth.make_asks(n_asks, "mixed")
    bits, npat, nmod, words = synth.intern_tables(w, a)
    full = orc.soa_eval_matrix(w.a, w.b, a.asks, a.opts, bits, words, 0, n_asks, 0, n_workers)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(z["first"].view(np.uint32), full["col_first"])
        assert np.array_equal(z["best"], full["row_best"])
        assert np.array_equal(z["cnt"].view(np.uint32), full["row_count"])


def test_shard_ranges_cover_the_table():
    for n, world in [(10, 3), (8, 8), (5, 8), (1_000_000, 4), (0, 2)]:
        spans = [shard_range(n, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 == b0 and a0 <= a1
