"""The class-cache auction of protocol_b200/csrc/pm_auction.cuh restated in Python and held against the sequential
checker (orc_soa_auction: every unassigned ask scans every worker every round): ask classes sharing one 32-entry cache
and bound, the price cap folded into the outside option, the pool of candidates kept from the last table walk and
re-ranked before any new walk, walks over a cost-sorted worker order (re-sorted every few rounds) that stop once enough
kept candidates beat the composite bound on the unseen part, walks that start at the class's cost level (skip keys:
compatibility is static and sort keys only rise), and the per-ask scan when a value ties exactly with the outside
option.  Same assignments and the same number of rounds, on tie-heavy inputs.  This guards the ALGORITHM (the
exactness argument of DESIGN.md 4); the GPU tests guard the kernels."""
import numpy as np
import pytest

from oracle import pm_oracle as orc
from protocol_b200 import synth

NEG = -(1 << 61)                  # "no value" (kAucNeg)
INVALID = (1 << 63) - 1           # no cache yet (kThetaInvalid)
COMPLETE = -(1 << 63)             # the cache / pool holds every compatible worker (kThetaComplete)
NONE = 0xFFFFFFFF


def better(v1, w1, v2, w2):
    """(value desc, worker asc) order."""
    return v1 > v2 or (v1 == v2 and w1 < w2)


def bmax(a, b):
    return a if better(a[0], a[1], b[0], b[1]) else b


def auction_class_algorithm(t, cap, S=1, eps=1, K=32, stripe=1024, lanes=256, resort_every=5):
    """Returns (ask -> worker, rounds, counters).  `stripe` workers per walk step, `lanes` threads per scanning CTA
    (4 kept candidates each), the worker order re-sorted every `resort_every` rounds."""
    asks, T, W = t["asks"], len(t["asks"]), len(t["wa"])
    # classes: asks with identical requirement rows
    keys, reps, class_of = {}, [], np.zeros(T, int)
    for i in range(T):
        r = asks[i]
        key = (int(r["flags"]), int(r["cpu_cores"]), int(r["ram_mb"]), int(r["storage_gb"]),
               t["opts"][r["opt_off"]:r["opt_off"] + r["n_opts"]].tobytes())
        if key not in keys:
            keys[key] = len(reps)
            reps.append(i)
        class_of[i] = keys[key]
    C = len(reps)
    ev = orc.soa_eval_matrix(t["wa"], t["wb"], np.ascontiguousarray(asks[reps]), t["opts"], t["bits"], t["words"], 0, C, 0, W,
                             threads=4, want_cost=True)
    compat = ev["cost"] != np.iinfo(np.int64).max
    ap = t["wb"]["ext_ask_price"].astype(np.int64)
    perm = np.lexsort((np.arange(W), ap))                     # walk order: (cost at the last sort, worker)
    order = {"perm": perm, "key": ap[perm] * S}
    price = np.zeros(W, np.int64)
    owner = np.full(W, NONE, np.int64)
    assigned = np.full(T, NONE, np.int64)
    withdrawn = np.zeros(T, bool)
    cand = [[] for _ in range(C)]                             # the 32-entry cache of each class ...
    theta = [(INVALID, NONE)] * C                             # ... and the bound on everything outside it
    pool = [None] * C                                         # what the lanes held after the class's last walk ...
    pool_bound = [(INVALID, NONE)] * C                        # ... and the bound on everything outside the pool
    skip_key = [0] * C                                        # every compatible worker of the class has a sort key >= this
    count = dict(full=0, refill=0, fb=0, skipped=0)

    def value(w):
        return int(-(ap[w] * S) - price[w])

    def select(lists, dropped):
        """top K of the kept candidates, best two values, and the best (value, worker) left out"""
        kept = sorted((e for lane in lists for e in lane), key=lambda e: (-e[0], e[1]))
        top = kept[:K]
        bound = kept[K] if len(kept) > K else (NEG, NONE)
        bound = bmax(bound, dropped)
        b1 = top[0] if top else (NEG, NONE)
        b2 = top[1][0] if len(top) > 1 else NEG
        return top, b1, b2, bound

    def store_cache(c, top, bound):
        cand[c] = [e[1] for e in top]
        theta[c] = (COMPLETE, NONE) if bound[0] == NEG else bound

    def class_refresh(c):
        # 1. re-rank the pool at the current prices; good enough while its best two beat its bound
        if pool_bound[c][0] != INVALID:
            lists = [[(value(w), int(w)) for w in lane] for lane in pool[c]]
            outside = pool_bound[c] if pool_bound[c][0] != COMPLETE else (NEG, NONE)
            top, b1, b2, bound = select(lists, outside)
            if bound[0] == NEG or (b2 >= bound[0] and better(b1[0], b1[1], bound[0], bound[1])):
                store_cache(c, top, bound)
                count["refill"] += 1
                return
        # 2. walk the cost-sorted table: each lane keeps its 4 best and remembers the best it dropped
        count["full"] += 1
        lists = [[] for _ in range(lanes)]
        dropped = [(NEG, NONE)] * lanes
        unseen, first_good = None, None
        n_stripes = (W + stripe - 1) // stripe
        # skip keys: the first stripe whose last key reaches the class's key (stripes before it hold nobody compatible)
        j0 = next((j for j in range(n_stripes) if order["key"][min((j + 1) * stripe, W) - 1] >= skip_key[c]), n_stripes)
        count["skipped"] += j0
        first_hit = None
        for k in range(j0, n_stripes):
            n = min(stripe, W - k * stripe)
            for i in range(n):
                w = order["perm"][k * stripe + i]
                if not compat[c, w]:
                    continue
                if first_hit is None:
                    first_hit = k
                lane = i % lanes
                lists[lane].append((value(w), int(w)))
                lists[lane].sort(key=lambda e: (-e[0], e[1]))
                if len(lists[lane]) > 4:
                    dropped[lane] = bmax(dropped[lane], lists[lane].pop())
            # every later worker ranks at or below (-key_last, worker_last + 1): prices only rise
            u, last = -int(order["key"][k * stripe + n - 1]), int(order["perm"][k * stripe + n - 1])
            beating = sum(1 for lane in lists for e in lane if e[0] > u or (e[0] == u and e[1] <= last))
            if beating > K and first_good is None:
                first_good = k
            if first_good is not None and (beating > 256 or k - first_good >= 8):
                if k + 1 < n_stripes:
                    unseen = (u, last + 1)
                break
        # nobody compatible before stripe first_hit (none at all if the walk ran to the end without meeting one); keys only rise
        skip_key[c] = max(skip_key[c], int(order["key"][first_hit * stripe]) if first_hit is not None else (1 << 62))
        pool[c] = [[e[1] for e in lane] for lane in lists]
        best_dropped = (NEG, NONE)
        for d in dropped:
            best_dropped = bmax(best_dropped, d)
        top, b1, b2, bound = select(lists, best_dropped)
        store_cache(c, top, bound)
        outside = best_dropped if unseen is None else bmax(best_dropped, unseen)
        pool_bound[c] = (COMPLETE, NONE) if outside[0] == NEG else outside

    def decide(tk):
        """0 withdraw, 1 bid (worker, best, second), 2 the class cache cannot tell"""
        c, out = class_of[tk], -((int(cap[tk]) + 1) * S)
        th, thw = theta[c]
        if th == INVALID:
            return 2, None
        entries = [(value(w), int(w)) for w in cand[c]]
        affordable = [e for e in entries if ap[e[1]] <= cap[tk]]
        if affordable:
            bf, wf = max(affordable, key=lambda e: (e[0], -e[1]))
            if better(bf, wf, th, thw):
                if bf < out:
                    return 0, None
                others = [e[0] for e in entries if e[1] != wf]
                bx = max(others) if others else NEG
                return (1, (wf, bf, max(bx, out))) if th <= max(bx, out) else (2, None)
            upper = max(bf, th)
        else:
            upper = th
        return (0, None) if upper < out else (2, None)

    rounds = 0
    while True:
        active = [i for i in range(T) if assigned[i] == NONE and not withdrawn[i]]
        if not active:
            break
        if rounds and rounds % resort_every == 0:             # bid-up workers move back in the walk order
            key = ap * S + price
            perm = np.lexsort((np.arange(W), key))
            order["perm"], order["key"] = perm, key[perm]
        rounds += 1
        bid_w, bid_p = {}, {}

        def place(tk, verdict, d):
            if verdict == 0:
                withdrawn[tk] = True
            else:
                bid_w[tk] = d[0]
                bid_p[tk] = int(price[d[0]]) + (d[1] - d[2]) + eps

        retry, requested = [], set()
        for tk in active:                                     # pass 0: from the class caches as they are
            verdict, d = decide(tk)
            if verdict == 2:
                retry.append(tk)
                requested.add(class_of[tk])
            else:
                place(tk, verdict, d)
        for c in requested:
            class_refresh(c)
        for tk in retry:                                      # pass 1: after the refreshes
            verdict, d = decide(tk)
            if verdict != 2:
                place(tk, verdict, d)
                continue
            count["fb"] += 1                                  # a tie exactly at the outside option: the ask's own scan
            c, out = class_of[tk], -((int(cap[tk]) + 1) * S)
            ws = np.flatnonzero(compat[c] & (ap <= cap[tk]))
            if len(ws) == 0:
                withdrawn[tk] = True
                continue
            v = -(ap[ws] * S) - price[ws]
            by_value = np.lexsort((ws, -v))
            ws, v = ws[by_value], v[by_value]
            if v[0] < out:
                withdrawn[tk] = True
                continue
            second = max(int(v[1]) if len(ws) > 1 else NEG, out)
            place(tk, 1, (int(ws[0]), int(v[0]), second))
        winner = {}                                           # highest bid, ties to the lowest ask
        for tk in sorted(bid_w):
            w = bid_w[tk]
            if w not in winner or bid_p[tk] > bid_p[winner[w]]:
                winner[w] = tk
        for w, tk in winner.items():
            if owner[w] != NONE:
                assigned[owner[w]] = NONE
            owner[w], assigned[tk], price[w] = tk, w, bid_p[tk]
    return assigned.astype(np.uint32), rounds, count


def _tables(n_base, copies, n_workers, plo, phi, seed):
    w = synth.make_workers(n_workers, seed=synth.SEED_WORKERS + seed)
    a = synth.make_asks(n_base, "mixed", seed=synth.SEED_ASKS + seed)
    bits, npat, nmod, words = synth.intern_tables(w, a)
    rng = np.random.default_rng(seed)
    idx = rng.permutation(np.repeat(np.arange(n_base), copies))
    wb = w.b.copy()
    wb["ext_ask_price"] = rng.integers(plo, phi + 1, n_workers).astype(np.uint32)
    t = dict(asks=np.ascontiguousarray(a.asks[idx]), opts=a.opts, wa=w.a, wb=wb, bits=bits, words=words)
    cap = rng.integers(max(plo - 2, 0), phi + 3, len(idx)).astype(np.uint32)
    return t, cap


@pytest.mark.parametrize("n_base,copies,n_workers,plo,phi,seed", [
    (8, 30, 1500, 10, 14, 5),       # few classes, many identical bidders, almost every value tied
    (20, 12, 2500, 10, 40, 6),
    (3, 80, 1200, 7, 9, 8),         # caps straddle the whole price range: ties exactly at the outside option
])
@pytest.mark.parametrize("stripe,lanes", [(1024, 256), (64, 16)], ids=["kernel_shape", "tiny_stripes"])
def test_class_cache_auction_equals_the_sequential_checker(n_base, copies, n_workers, plo, phi, seed, stripe, lanes):
    t, cap = _tables(n_base, copies, n_workers, plo, phi, seed)
    want, price, rounds = orc.soa_auction(t["wa"], t["wb"], t["asks"], t["opts"], t["bits"], t["words"], cap)
    got, r, st = auction_class_algorithm(t, cap, stripe=stripe, lanes=lanes, resort_every=7)
    assert np.array_equal(got, want) and r == rounds
    assert st["full"] >= 1 and st["refill"] >= 1          # both refresh paths ran


def test_skip_keys_leave_out_the_cheap_end_of_the_table():
    """Prices that grow with the GPU count: the classes that want many GPUs have all their compatible workers deep in the
    cost order, and every walk of theirs after the first starts there.  Same assignment and rounds as the checker."""
    t, cap = _tables(20, 12, 2500, 10, 40, 16)
    rng = np.random.default_rng(16)
    wb = t["wb"].copy()
    wb["ext_ask_price"] = (10 + 12 * np.log2(np.maximum(t["wa"]["gpu_count"], 1)) + rng.integers(0, 4, len(wb))).astype(np.uint32)
    t["wb"] = wb
    cap = rng.integers(8, 60, len(cap)).astype(np.uint32)
    want, price, rounds = orc.soa_auction(t["wa"], t["wb"], t["asks"], t["opts"], t["bits"], t["words"], cap)
    got, r, st = auction_class_algorithm(t, cap, K=4, stripe=32, lanes=2, resort_every=7)   # a tiny cache and pool: many walks
    assert np.array_equal(got, want) and r == rounds
    assert st["full"] > 2 * 20 and st["skipped"] > 5 * st["full"]   # classes walk again and again, and start deep in the table


# ---- the two device formulations the selection and the skip rest on, restated and checked exhaustively enough on the CPU
def _better(a, b):
    return a[0] > b[0] or (a[0] == b[0] and a[1] < b[1])


def best64_network(arr):
    """auction_select_cta: 1024 entries, sorted runs of 4 (odd threads reversed) -> blocks of 64 sorted by the bitonic
    network (even blocks descending, odd ascending) -> 8 + 4 + 2 + 1 merges that keep the entry-wise better half of a
    descending and an ascending block and re-sort it (descending / ascending by turns).  Returns the first 64 entries."""
    arr = list(arr)
    k = 8
    while k <= 64:
        j = k >> 1
        while j:
            new = list(arr)
            for idx in range(512):
                i = 2 * idx - (idx & (j - 1))
                a, b = arr[i], arr[i + j]
                if (_better(b, a) if (i & k) == 0 else _better(a, b)):
                    new[i], new[i + j] = b, a
            arr, j = new, j >> 1
        k <<= 1
    S, P = 64, 8
    while P:
        new = list(arr)
        for q in range(P):
            A, B = q * 2 * S, q * 2 * S + S
            for i in range(64):
                if _better(arr[B + i], arr[A + i]):
                    new[A + i] = arr[B + i]
        arr, j = new, 32
        while j:
            new = list(arr)
            for q in range(P):
                A = q * 2 * S
                for lane in range(32):
                    i = A + 2 * lane - (lane & (j - 1))
                    a, b = arr[i], arr[i + j]
                    if (_better(b, a) if q % 2 == 0 else _better(a, b)):
                        new[i], new[i + j] = b, a
            arr, j = new, j >> 1
        S, P = S * 2, P // 2
    return arr[:64]


@pytest.mark.parametrize("seed", range(12))
def test_selection_network_keeps_exactly_the_best_64(seed):
    import random
    rng = random.Random(seed)
    arr = [None] * 1024
    fill = [0, 1, 2, 3, 4, 4, 4] if seed % 3 else [0, 0, 0, 1]          # dense pools and nearly empty ones
    workers = rng.sample(range(1 << 20), 1024)
    for tid in range(256):
        items = sorted(((rng.randint(-6, 0) if seed % 2 else -rng.randint(0, 1 << 30), workers[tid * 4 + q])
                        for q in range(rng.choice(fill))), key=lambda e: (-e[0], e[1]))
        items += [(NEG, NONE)] * (4 - len(items))
        for q in range(4):
            arr[tid * 4 + (3 - q if tid & 1 else q)] = items[q]
    assert best64_network(arr) == sorted(arr, key=lambda e: (-e[0], e[1]))[:64]
    # the packed form: one 64-bit key per entry, ascending = (value desc, worker asc); an empty slot is all ones
    pack = lambda e: (1 << 64) - 1 if e[1] == NONE else ((-e[0]) << 24) | e[1]
    assert sorted(arr, key=pack)[:64] == sorted(arr, key=lambda e: (-e[0], e[1]))[:64]


def test_start_stripe_search_is_a_lower_bound_search():
    """The 32-ary search one warp runs over the stripe ends: first stripe whose last key is >= the class's skip key."""
    import bisect
    import random

    def search(ends, key):
        lo, hi = 0, len(ends)
        while lo < hi:
            step = (hi - lo + 31) // 32
            ge = [True if lo + lane * step >= hi else ends[lo + lane * step] >= key for lane in range(32)]
            f = ge.index(True) if True in ge else 32
            if f == 0:
                hi = lo
                break
            nlo = lo + (f - 1) * step + 1
            if f < 32 and lo + f * step < hi:
                hi = lo + f * step
            lo = nlo
        return min(lo, hi)

    rng = random.Random(3)
    for _ in range(4000):
        n = rng.choice([1, 2, 31, 32, 33, 64, 100, 977, 1024, 1025, 4000])
        ends = sorted(rng.randint(0, 60) for _ in range(n))
        key = rng.randint(-1, 62)
        assert search(ends, key) == bisect.bisect_left(ends, key)


def test_packed_claim_orders_by_bid_then_lowest_ask():
    """auction_commit: bid and bidder in one signed 64-bit word, (bid << 23) | (2^23 - 1 - ask): the maximum is the highest bid
    and, among equal bids, the lowest ask — what atomicMax on the bid followed by atomicMin on the bidder gives; stays
    positive (the 'no bid' value is negative) for bids below 2^40 and asks below 2^23."""
    import random
    rng = random.Random(23)
    mask = (1 << 23) - 1
    for _ in range(2000):
        bids = [(rng.choice([1, 2, 3, rng.randrange(1, 1 << 40), (1 << 40) - 1]), rng.randrange(0, 1 << 23)) for _ in range(rng.randint(1, 12))]
        keys = [(b << 23) | (mask - t) for b, t in bids]
        assert all(0 < k < (1 << 63) for k in keys)
        best = max(keys)
        want = min((t for b, t in bids if b == max(b for b, _ in bids)))
        assert best >> 23 == max(b for b, _ in bids) and mask - (best & mask) == want
