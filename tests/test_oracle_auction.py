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


def brute_force_best_surplus_masked(price, cap, ok):
    T, W = len(cap), len(price)
    best = 0
    for k in range(0, min(T, W) + 1):
        for ts in itertools.combinations(range(T), k):
            for ws in itertools.permutations(range(W), k):
                if all(ok[t, w] and price[w] <= cap[t] for t, w in zip(ts, ws)):
                    best = max(best, sum(int(cap[t]) + 1 - int(price[w]) for t, w in zip(ts, ws)))
    return best


@pytest.mark.parametrize("seed", range(6))
def test_reputation_floor_is_a_feasibility_clause(seed):
    """north_star worker column `reputation`: reputation[w] >= min_reputation[t] joins the feasibility predicate;
    with the exact cost scale the result is the optimum of the masked problem (brute force)."""
    T, W = 4, 5
    a, b, asks, opts, cap = tiny(200 + seed, T, W)
    rng = np.random.default_rng(seed)
    rep = rng.integers(0, 4, W).astype(np.uint32)
    floor = rng.integers(0, 5, T).astype(np.uint32)
    bits = np.array([0xFFFFFFFF], dtype=np.uint32)
    out, _, _ = orc.soa_auction(a, b, asks, opts, bits, 1, cap, cost_scale=T + 1, reputation=rep, min_reputation=floor)
    ok = rep[None, :] >= floor[:, None]
    for t in range(T):
        if out[t] != NONE:
            assert ok[t, out[t]] and b["ext_ask_price"][out[t]] <= cap[t]
    used = out[out != NONE]
    assert len(set(used.tolist())) == len(used)
    surplus = sum(int(cap[t]) + 1 - int(b["ext_ask_price"][out[t]]) for t in range(T) if out[t] != NONE)
    assert surplus == brute_force_best_surplus_masked(b["ext_ask_price"], cap, ok)
    # missing columns are columns of zeros
    base, _, r0 = orc.soa_auction(a, b, asks, opts, bits, 1, cap)
    z, _, r1 = orc.soa_auction(a, b, asks, opts, bits, 1, cap, reputation=rep, min_reputation=np.zeros(T, dtype=np.uint32))
    assert np.array_equal(base, z) and r0 == r1
    none, _, _ = orc.soa_auction(a, b, asks, opts, bits, 1, cap, min_reputation=np.ones(T, dtype=np.uint32))
    assert (none == NONE).all()


def lsa_optimum(price, cap, feasible, scale=1):
    """Minimum of sum over asks of (price * scale if assigned to a feasible worker within the cap, else the outside
    option (cap + 1) * scale) — scipy's Hungarian solver on the masked matrix: an anchor that is NOT this repo's code."""
    from scipy.optimize import linear_sum_assignment

    T, W = feasible.shape
    BIG = 1 << 40
    C = np.full((T, W + T), BIG, dtype=np.int64)
    ok = feasible & (price[None, :] <= cap[:, None])
    C[:, :W][ok] = np.broadcast_to(price[None, :].astype(np.int64) * scale, (T, W))[ok]
    C[np.arange(T), W + np.arange(T)] = (cap.astype(np.int64) + 1) * scale        # ask t stays out
    r, c = linear_sum_assignment(C)
    return int(C[r, c].sum())


def assignment_cost(out, price, cap, scale=1):
    return sum((int(price[out[t]]) if out[t] != NONE else int(cap[t]) + 1) * scale for t in range(len(cap)))


@pytest.mark.parametrize("T,W,seed", [(60, 200, 1), (300, 2000, 2), (500, 300, 3)])
def test_auction_total_cost_against_scipy_linear_sum_assignment(T, W, seed):
    """External anchor: with cost_scale = 1, eps = 1 the auction's total cost is within T * eps of
    the optimum of the feasibility-masked, cap-filtered assignment problem; with cost_scale = T + 1 it IS the optimum."""
    from protocol_b200 import synth

    w = synth.make_workers(W, seed=synth.SEED_WORKERS + seed, price="loguniform")
    a = synth.make_asks(T, "mixed", seed=synth.SEED_ASKS + seed)
    bits, npat, nmod, words = synth.intern_tables(w, a)
    cap = np.exp(np.log(20) + synth._unit(synth.SEED_EXT + seed, T, 3) * np.log(60)).astype(np.uint32)
    ev = orc.soa_eval_matrix(w.a, w.b, a.asks, a.opts, bits, words, 0, T, 0, W, want_cost=True)
    feasible = ev["cost"] != abi.PM_COST_INF
    price = w.b["ext_ask_price"].astype(np.int64)
    opt = lsa_optimum(price, cap, feasible)
    out, _, rounds = orc.soa_auction(w.a, w.b, a.asks, a.opts, bits, words, cap, cost_scale=1)
    for t in np.flatnonzero(out != NONE):
        assert feasible[t, out[t]] and price[out[t]] <= cap[t]
    used = out[out != NONE]
    assert len(set(used.tolist())) == len(used)
    got = assignment_cost(out, price, cap)
    assert opt <= got <= opt + T, (opt, got)
    out2, _, _ = orc.soa_auction(w.a, w.b, a.asks, a.opts, bits, words, cap, cost_scale=T + 1)
    assert assignment_cost(out2, price, cap) == opt
    # eps-scaling (phases 64, 16, 4, 1; assignment cleared, prices kept) is NOT covered by that bound here: prices only
    # rise, a coarse phase overshoots them, and with price caps (an outside option per ask) the asks that were outbid at
    # inflated prices withdraw for good in the eps = 1 phase.  The result stays feasible but can be far from the optimum
    # (54 % above it on the first instance) — which is why eps_start = 1 stays the engine's default.
    out3, _, rounds3 = orc.soa_auction(w.a, w.b, a.asks, a.opts, bits, words, cap, cost_scale=1, eps_start=64, eps_div=4)
    for t in np.flatnonzero(out3 != NONE):
        assert feasible[t, out3[t]] and price[out3[t]] <= cap[t]
    assert assignment_cost(out3, price, cap) >= opt
