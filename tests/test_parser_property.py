"""Property test: the product's requirement parser (csrc/pm_host.cpp) and the oracle's restatement of
ComputeRequirements::from_str (node.rs:180-374) accept/reject the same strings and, when they accept,
describe the same requirement — checked by evaluating both against random nodes (string predicate vs
interned-table predicate)."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from helpers import spec_to_orc_node, spec_to_rows
from oracle import pm_oracle as orc
from protocol_b200 import abi
from protocol_b200._lib import PrimeMatchError
from protocol_b200.engine import Interner

KEYS = ["gpu:count", "gpu:model", "gpu:memory_mb", "gpu:memory_mb_min", "gpu:memory_mb_max", "gpu:total_memory_min",
        "gpu:total_memory_max", "cpu:cores", "ram_mb", "storage_gb", "gpu_model", "bogus"]
MODELS = ["A100", "h100,h200", "RTX 4090", "rtx_3090", "", "nvidia", "a100, h100 ,l40s"]
NUMS = ["0", "1", "2", "4", "8", "24000", "40000", "80000", "160000", "4294967295", "4294967296", "abc", "", "+8", " 16 "]

pair = st.tuples(st.sampled_from(KEYS), st.one_of(st.sampled_from(NUMS), st.sampled_from(MODELS)))
req_string = st.lists(pair, min_size=0, max_size=7).map(
    lambda kv: ";".join(f"{k}={v}" if k != "bogus" or v else k for k, v in kv))

SPECS = [dict(gpu_count=c, gpu_model=m, gpu_mem=g, cpu_cores=k, ram=r, storage=s)
         for c in (None, 0, 1, 4, 8) for m in (None, "NVIDIA A100 80GB", "RTX 4090")
         for g in (None, 24000, 80000) for k in (None, 16) for r in (None, 64000) for s in (None, 500)][::7]


@settings(max_examples=400, deadline=None)
@given(req_string)
def test_product_parser_equals_oracle_parser(s):
    try:
        oreq = orc.Req(s)
        o_ok = True
    except ValueError:
        o_ok = False
    it = Interner()
    try:
        ask, opts = it.parse(s)
        p_ok = True
    except PrimeMatchError as e:
        assert e.status == abi.PM_E_PARSE
        p_ok = False
    assert o_ok == p_ok, s
    if not o_ok:
        return
    assert int(ask["n_opts"]) == oreq.n_gpu()
    ask = ask.copy()
    ask["opt_off"] = 0
    for sp in SPECS:
        a, b = spec_to_rows(sp, it)
        bits, npat, nmod, words = it.table()
        want = orc.meets(spec_to_orc_node(sp), oreq)
        got = orc.soa_compatible(a, b, ask, opts, bits, words)
        assert got == want, (s, sp)
