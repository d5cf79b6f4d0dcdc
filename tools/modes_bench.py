"""Time the non-headline modes of the pass (tail sweep, proximity formation) next to the CPU oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pm_oracle as orc
from protocol_b200 import abi, synth
from protocol_b200.engine import Engine

SIZES = [(1, 1), (2, 2), (2, 4), (3, 3), (4, 8), (1, 3)]


def run(T, W, mode, group_sizes, label, check=True):
    w = synth.make_workers(W, with_addresses=False)
    a = synth.make_asks(T, "mixed", group_sizes=group_sizes)
    bits, npat, nmod, words = synth.intern_tables(w, a)
    eng = Engine(timing=True)
    eng.set_asks(a.asks, a.opts); eng.set_model_table(bits, npat, nmod, words); eng.set_workers(w.a, w.b)
    eng.set_locations(w.lat, w.lon)
    eng.match(mode)
    t0 = time.perf_counter(); eng.match(mode); eng.sync(); dt = time.perf_counter() - t0
    res = eng.fetch()
    out = {"case": label, "T": T, "W": W, "gpu_ms": round(dt * 1e3, 2), "ms_resolve": round(res.stats["ms_resolve"], 2),
           "groups": res.n_groups, "bumped": res.stats["n_bumped"]}
    if check:
        t0 = time.perf_counter()
        og = orc.soa_form_groups(w.a, w.b, a.asks, a.opts, bits, words, lat=w.lat, lon=w.lon,
                                 proximity=(mode & 0xFF) == abi.PM_MODE_PROXIMITY)
        out["oracle_cpu_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
        out["identical"] = bool(np.array_equal(res.group_ask, og.cfg) and np.array_equal(res.group_off, og.off)
                                and np.array_equal(res.group_members, og.members))
    print(out, flush=True)
    eng.close()


if __name__ == "__main__":
    run(2000, 100_000, abi.PM_MODE_FIRST_FIT, SIZES, "first-fit, mixed group sizes (tail sweep)")
    run(2000, 100_000, abi.PM_MODE_PROXIMITY, None, "proximity, solo groups")
    run(2000, 20_000, abi.PM_MODE_PROXIMITY, SIZES, "proximity, mixed group sizes")
    run(2000, 100_000, abi.PM_MODE_PROXIMITY, SIZES, "proximity, mixed group sizes")
