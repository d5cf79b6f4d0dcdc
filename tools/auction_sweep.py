"""Extension mode: one synthetic market (BASELINE configs[2] shape), several engine settings (environment read at
pm_create): PM_TUNE_AUCTION_POOL=good,extra (pool fill of a class walk), PM_TUNE_AUCTION bits 8+ (batches of 32 rounds
between re-sorts of the cost-sorted worker copy).  usage: auction_sweep.py TxW "POOL;TUNE" ..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from protocol_b200 import abi, synth
from protocol_b200.engine import Engine

T, W = (int(x) for x in sys.argv[1].split("x"))
w = synth.make_workers(W)
a = synth.make_asks(T, "mixed")
bits, npat, nmod, words = synth.intern_tables(w, a)
w.b["ext_ask_price"] = np.exp(np.log(10) + synth._unit(synth.SEED_EXT, W, 1) * np.log(200)).astype(np.uint32)
cap = np.exp(np.log(20) + synth._unit(synth.SEED_EXT, T, 2) * np.log(150)).astype(np.uint32)
first = None
for setting in sys.argv[2:]:
    pool, tune = (setting.split(";") + [""])[:2]
    for k, v in (("PM_TUNE_AUCTION_POOL", pool), ("PM_TUNE_AUCTION", tune)):
        if v:
            os.environ[k] = v
        else:
            os.environ.pop(k, None)
    eng = Engine(timing=True)
    eng.set_asks(a.asks, a.opts); eng.set_model_table(bits, npat, nmod, words); eng.set_workers(w.a, w.b)
    eng.set_price_caps(cap)
    t0 = time.time()
    eng.match(abi.PM_MODE_AUCTION)
    dt = time.time() - t0
    res = eng.fetch()
    st = res.stats
    sold = (res.group_ask.copy(), res.group_members.copy())
    same = True if first is None else (np.array_equal(first[0], sold[0]) and np.array_equal(first[1], sold[1]))
    first = first or sold
    print({"pool": pool, "tune": tune, "wall_s": round(dt, 3), "rounds": st["n_rounds"], "evals": st["evals"], "class_scans": st["n_tiles"],
           "refills": st["n_build_launches"], "ask_scans": st["n_fused_launches"], "assigned": res.n_groups, "same_assignment": same}, flush=True)
    eng.close()
