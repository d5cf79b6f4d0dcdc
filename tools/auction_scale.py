"""Extension mode at scale (self-oracle mode; no reference counterpart): rounds and time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from protocol_b200 import abi, synth
from protocol_b200.engine import Engine

for T, W in [(int(x.split("x")[0]), int(x.split("x")[1])) for x in sys.argv[1:]]:
    w = synth.make_workers(W)
    a = synth.make_asks(T, "mixed")
    bits, npat, nmod, words = synth.intern_tables(w, a)
    u = synth._unit(synth.SEED_EXT, W, 1)
    w.b["ext_ask_price"] = np.exp(np.log(10) + u * np.log(200)).astype(np.uint32)        # log-uniform 10..2000
    cap = np.exp(np.log(20) + synth._unit(synth.SEED_EXT, T, 2) * np.log(150)).astype(np.uint32)
    eng = Engine(timing=True)
    eng.set_asks(a.asks, a.opts); eng.set_model_table(bits, npat, nmod, words); eng.set_workers(w.a, w.b)
    eng.set_price_caps(cap)
    t0 = time.time()
    eng.match(abi.PM_MODE_AUCTION)
    dt = time.time() - t0
    res = eng.fetch()
    st = res.stats
    print({"T": T, "W": W, "wall_s": round(dt, 3), "rounds": st["n_rounds"], "evals": st["evals"],
           "class_scans": st["n_tiles"], "refills": st["n_build_launches"], "ask_scans": st["n_fused_launches"], "assigned": res.n_groups,
           "ms_rounds": round(st["ms_fused"], 1)}, flush=True)
    eng.close()
