"""Diagnostic: smallest duplicated-ask / tied-price auction instance where the engine differs from the checker."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from protocol_b200 import abi
from protocol_b200.engine import Engine
from oracle import pm_oracle as orc
from test_gpu_parity import synth_tables
from helpers import load_engine

def run(n_base, copies, W, plo, phi, seed):
    w, a, t = synth_tables(n_base, W, "mixed", seed_shift=seed)
    rng = np.random.default_rng(seed)
    idx = rng.permutation(np.repeat(np.arange(n_base), copies))
    t["asks"] = np.ascontiguousarray(t["asks"][idx])
    wb = t["wb"].copy(); wb["ext_ask_price"] = rng.integers(plo, phi + 1, W).astype(np.uint32); t["wb"] = wb
    cap = rng.integers(max(plo - 2, 0), phi + 3, len(idx)).astype(np.uint32)
    eng = Engine(); load_engine(eng, t); eng.set_price_caps(cap); eng.match(abi.PM_MODE_AUCTION); res = eng.fetch()
    got = np.full(len(cap), abi.PM_NONE, dtype=np.uint32)
    for g, (ask, members) in enumerate(res.groups()): got[ask] = members[0]
    want, price, rounds = orc.soa_auction(t["wa"], t["wb"], t["asks"], t["opts"], t["bits"], t["words"], cap)
    eng.close()
    return t, cap, got, want, res.stats["n_rounds"], rounds, res.stats

if len(sys.argv) > 1:
    W, n_base, copies, seed = map(int, sys.argv[1:5])
    t, cap, got, want, r, rw, st = run(n_base, copies, W, 7, 9, seed)
    print("single", W, n_base, copies, seed, np.array_equal(got, want), r, rw, st["n_tiles"], st["n_fused_launches"])
    sys.exit(0)
found = None
for W in (1100, 2100, 4200, 6000):
    for n_base, copies in ((1, 4), (1, 20), (2, 10), (3, 30), (3, 120), (3, 200)):
        for seed in (1, 2, 3):
            t, cap, got, want, r, rw, st = run(n_base, copies, W, 7, 12, seed)
            ok = np.array_equal(got, want) and r == rw
            print(W, n_base, copies, seed, "ok" if ok else "DIFF", r, rw, st["n_tiles"], st["n_fused_launches"], flush=True)
            if not ok and found is None:
                found = (W, n_base, copies, seed)
                os.makedirs("gpurun_out", exist_ok=True)
                np.savez("gpurun_out/auction_diff.npz", got=got, want=want, cap=cap, W=W, n_base=n_base, copies=copies, seed=seed)
    if found: break
print("first failing:", found)
