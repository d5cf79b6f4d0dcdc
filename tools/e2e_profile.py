"""Where does the end-to-end step (host tables -> H2D -> match -> D2H) spend its time?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from protocol_b200 import abi, synth
from protocol_b200.engine import Engine, pinned_empty

T, W = 100_000, 100_000
w = synth.make_workers(W); a = synth.make_asks(T, "mixed")
bits, npat, nmod, words = synth.intern_tables(w, a)
eng = Engine()
eng.set_asks(a.asks, a.opts); eng.set_model_table(bits, npat, nmod, words); eng.set_workers(w.a, w.b)
pa = pinned_empty(W, abi.WORKER_A); pb = pinned_empty(W, abi.WORKER_B); pa[:] = w.a; pb[:] = w.b
eng.match(); eng.fetch()
acc = {}
N = 10
for _ in range(N):
    t0 = time.perf_counter(); eng.set_asks(a.asks, a.opts)
    t1 = time.perf_counter(); eng.upsert_workers(pa, pb, sync=False)
    t2 = time.perf_counter(); eng.match()
    t3 = time.perf_counter(); r = eng.fetch(copy=False)
    t4 = time.perf_counter()
    for k, v in (("set_asks", t1 - t0), ("upsert", t2 - t1), ("match", t3 - t2), ("fetch", t4 - t3), ("total", t4 - t0)):
        acc[k] = acc.get(k, 0) + v / N * 1e3
print({k: round(v, 3) for k, v in acc.items()})
