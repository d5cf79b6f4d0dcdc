"""EXTENSION self-oracle sanity (no reference counterpart — SURVEY 0): the sequential auction used as
the checker for PM_MODE_AUCTION is itself pinned against brute force on tiny instances: with
cost_scale = T+1 and eps = 1 the assignment is optimal (max cardinality is not the objective: each ask
either takes a worker at price <= cap or stays out, and the total of (cap+1 - price) is maximised)."""
import itertools

import numpy as np
import pytest

from oracle import pm_oracle as orc
from protocol_b200 import abi

NONE = 0xFFFFFFFF


def tiny(seed, T, W):
    rng = np.random.default_rng(seed)
    a = np.zeros(W, dtype=abi.WORKER_A)
    b = np.zeros(W, dtype=abi.WORKER_B)
    a["flags"] = abi.PM_W_HEALTHY | abi.PM_W_P2P
    b["ext_ask_price"] = rng.integers(1, 30, W)
    asks = np.zeros(T, dtype=abi.ASK)            # no requirements: everybody compatible
    asks["min_group_size"] = asks["max_group_size"] = 1
    opts = np.zeros(0, dtype=abi.GPU_OPT)
    cap = rng.integers(1, 30, T).astype(np.uint32)
    return a, b, asks, opts, cap


def brute_force_best_surplus(price, cap):
    """max over partial injective assignments of sum(cap[t] + 1 - price[w]) with price[w] <= cap[t]."""
    T, W = len(cap), len(price)
    best = 0
    for k in range(0, min(T, W) + 1):
        for ts in itertools.combinations(range(T), k):
            for ws in itertools.permutations(range(W), k):
                if all(price[w] <= cap[t] for t, w in zip(ts, ws)):
                    best = max(best, sum(int(cap[t]) + 1 - int(price[w]) for t, w in zip(ts, ws)))
    return best


@pytest.mark.parametrize("seed", range(6))
def test_exact_scale_is_optimal_on_tiny_instances(seed):
    T, W = 4, 5
    a, b, asks, opts, cap = tiny(seed, T, W)
    bits = np.array([0xFFFFFFFF], dtype=np.uint32)
    out, price, rounds = orc.soa_auction(a, b, asks, opts, bits, 1, cap, cost_scale=T + 1)
    used = out[out != NONE]
    assert len(set(used.tolist())) == len(used)
    surplus = sum(int(cap[t]) + 1 - int(b["ext_ask_price"][out[t]]) for t in range(T) if out[t] != NONE)
    assert all(b["ext_ask_price"][out[t]] <= cap[t] for t in range(T) if out[t] != NONE)
    assert surplus == brute_force_best_surplus(b["ext_ask_price"], cap)


def test_unit_scale_is_within_T_eps_of_optimal():
    T, W = 4, 6
    for seed in range(6):
        a, b, asks, opts, cap = tiny(100 + seed, T, W)
        bits = np.array([0xFFFFFFFF], dtype=np.uint32)
        out, _, _ = orc.soa_auction(a, b, asks, opts, bits, 1, cap, cost_scale=1)
        surplus = sum(int(cap[t]) + 1 - int(b["ext_ask_price"][out[t]]) for t in range(T) if out[t] != NONE)
        assert surplus >= brute_force_best_surplus(b["ext_ask_price"], cap) - T


def test_infeasible_asks_withdraw():
    a, b, asks, opts, cap = tiny(7, 3, 4)
    cap[:] = 0                                   # nobody is cheap enough
    bits = np.array([0xFFFFFFFF], dtype=np.uint32)
    out, _, rounds = orc.soa_auction(a, b, asks, opts, bits, 1, cap)
    assert (out == NONE).all() and rounds == 1
