"""Generates tests/golden/*.npz: expected results of the matching pass on small seeded swarms,
computed by the CPU oracle (oracle/pm_oracle.cpp — itself pinned to the reference's known-answer
tests).  The reference is Rust and cannot run here, so these are oracle outputs, not reference
outputs; they make the -m gpu suite independent of the oracle build on the GPU box and freeze the
expected answers in the repository.  Re-run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pm_oracle as orc  # noqa: E402
from protocol_b200 import synth  # noqa: E402

CASES = {
    # name: (n_asks, n_workers, kind, group_sizes, proximity, seed_shift)
    "first_fit_solo": (300, 4000, "mixed", None, False, 0),
    "first_fit_sizes": (200, 5000, "mixed", [(1, 1), (2, 2), (2, 4), (3, 3), (4, 8)], False, 1),
    "proximity_sizes": (100, 2500, "mixed", [(2, 2), (2, 4), (3, 3)], True, 2),
    "skewed_infeasible": (400, 3000, "skewed", None, False, 3),
}


def build(name):
    T, W, kind, sizes, prox, shift = CASES[name]
    w = synth.make_workers(W, seed=synth.SEED_WORKERS + shift, with_addresses=True)
    a = synth.make_asks(T, kind, seed=synth.SEED_ASKS + shift, group_sizes=sizes)
    bits, npat, nmod, words = synth.intern_tables(w, a)
    return w, a, bits, npat, nmod, words, prox


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name in CASES:
        w, a, bits, npat, nmod, words, prox = build(name)
        g = orc.soa_form_groups(w.a, w.b, a.asks, a.opts, bits, words, addr_rank=w.addr_rank, lat=w.lat, lon=w.lon,
                                proximity=prox)
        ev = orc.soa_eval_matrix(w.a, w.b, a.asks, a.opts, bits, words, 0, len(a), 0, len(w), threads=8)
        np.savez_compressed(os.path.join(here, f"{name}.npz"), group_ask=g.cfg, group_off=g.off, group_members=g.members,
                            ask_best=ev["row_best"], ask_count=ev["row_count"], col_first=ev["col_first"],
                            wa_crc=np.array([int(np.frombuffer(w.a.tobytes(), dtype=np.uint32).sum(dtype=np.uint64))]),
                            asks_crc=np.array([int(np.frombuffer(a.asks.tobytes(), dtype=np.uint32).sum(dtype=np.uint64))]))
        print(name, len(g), "groups")
