"""CPU checks of the checker's own helpers added for the wide-model and T = 1M GPU tests:
the acceptance table computed from strings with the reference's model clause (node.rs:463-484) equals the product
interner's table, and the per-worker first-feasible search equals the per-configuration loop for solo configurations."""
import numpy as np

import kat_vectors as kv
from oracle import pm_oracle as orc
from protocol_b200 import abi, synth
from protocol_b200.engine import Interner


def product_table(models, patterns):
    it = Interner()
    for i, m in enumerate(models):
        assert it.model(m) == i
    for i, p in enumerate(patterns):
        assert it.pattern(p) == i
    out = it.table()
    it.close()
    return out


def test_oracle_table_equals_interner_table_on_the_catalogues():
    models = [m for m, _ in synth.wide_model_catalogue(700)]
    patterns = list(synth.PATTERN_CATALOGUE) + ["mi300x", "rev3", "sxm,nvl", "geforce rtx 4090 ti", " h100 , ,a100", "", "_", "nvidia_h100"]
    ob, npat, nmod, words = orc.model_table(models, patterns)
    pb, p_npat, p_nmod, p_words = product_table(models, patterns)
    assert (npat, nmod, words) == (p_npat, p_nmod, p_words) == (len(patterns), 700, 22)
    assert np.array_equal(ob, pb)
    # spot-check bits against the scalar clause
    rng = np.random.default_rng(1)
    for p, m in zip(rng.integers(0, npat, 300), rng.integers(0, nmod, 300)):
        bit = (ob[p * words + (m >> 5)] >> (m & 31)) & 1
        assert bool(bit) == orc.model_matches(models[m], patterns[p])


def test_model_clause_vectors_of_the_reference_tests():
    """The (spec model, requirement model, expected) triples of node.rs:659-1241 through the table builder."""
    triples = [(s.get("gpu_model"), r, e) for _, _, s, r, e in kv.MEETS if s and s.get("gpu_model") and "gpu:model=" in r]
    assert triples
    for spec_model, req, expected in triples:
        req_model = [kvp.split("=", 1)[1] for kvp in req.split(";") if kvp.startswith("gpu:model=")][0]
        ob, npat, nmod, words = orc.model_table([spec_model], [req_model])
        pb = product_table([spec_model], [req_model])[0]
        assert np.array_equal(ob, pb)


def test_first_feasible_is_the_allocation_for_solo_configurations():
    w = synth.make_workers(3000, price="zipf")
    a = synth.make_asks(5000, "skewed")
    bits, npat, nmod, words = synth.intern_tables(w, a)
    og = orc.soa_form_groups(w.a, w.b, a.asks, a.opts, bits, words)
    first = orc.soa_first_feasible(w.a, w.b, a.asks, a.opts, bits, words, threads=4)
    assigned = np.flatnonzero(first != abi.PM_NONE)
    order = assigned[np.argsort(first[assigned], kind="stable")]
    assert np.array_equal(og.members, order.astype(np.uint32)) and np.array_equal(og.cfg, first[order])
    ev = orc.soa_eval_matrix(w.a, w.b, a.asks, a.opts, bits, words, 0, len(a), 0, len(w), threads=4)
    assert np.array_equal(first, ev["col_first"])
    # prices ride in the high word of the packed cost: the row minimum is the cheapest feasible worker
    feas = ev["row_best"] != abi.PM_COST_INF
    assert feas.any() and ((ev["row_best"][feas] >> 32) >= 1).all()
    wk = (ev["row_best"][feas] & 0xFFFFFFFF).astype(np.int64)
    assert np.array_equal(w.b["ext_ask_price"][wk].astype(np.int64), ev["row_best"][feas] >> 32)
