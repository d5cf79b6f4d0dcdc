"""Pins the oracle to the reference's own known-answer tests
(crates/shared/src/models/node.rs:659-1241, newest_task/mod.rs:29-54,
node_store.rs:419-465, node_groups/tests.rs:1381-1445) and pins the product's
host-side parser / interner / SoA predicate to the same vectors (CPU only)."""
import numpy as np
import pytest

import kat_vectors as kv
from helpers import TableBuilder, spec_to_orc_node
from oracle import pm_oracle as orc
from protocol_b200 import abi
from protocol_b200._lib import PrimeMatchError
from protocol_b200.engine import Interner, sort_configs


@pytest.mark.parametrize("name,line,specs,req,expected", kv.MEETS, ids=[m[0] for m in kv.MEETS])
def test_oracle_meets(name, line, specs, req, expected):
    node = spec_to_orc_node(specs)
    assert orc.meets(node, orc.Req(req)) is expected


@pytest.mark.parametrize("name,line,specs,req,expected", kv.MEETS, ids=[m[0] for m in kv.MEETS])
def test_host_tables_meets(name, line, specs, req, expected):
    """product parser + interner -> SoA rows; SoA restatement of the predicate agrees with the KAT."""
    tb = TableBuilder()
    tb.add_config(req)
    tb.add_node(specs)
    t = tb.tables()
    assert orc.soa_compatible(t["wa"][0], t["wb"][0], t["asks"][0], t["opts"], t["bits"], t["words"]) is expected


@pytest.mark.parametrize("name,line,s,ok", kv.PARSE_OK, ids=[p[0] for p in kv.PARSE_OK])
def test_parser_accepts(name, line, s, ok):
    if ok:
        orc.Req(s)
        Interner().parse(s)
    else:
        with pytest.raises(ValueError):
            orc.Req(s)
        with pytest.raises(PrimeMatchError) as ei:
            Interner().parse(s)
        assert ei.value.status == abi.PM_E_PARSE


PRESENT = {"count": abi.PM_O_COUNT, "model": abi.PM_O_MODEL, "memory_mb": abi.PM_O_MEM,
           "memory_mb_min": abi.PM_O_MEM_MIN, "memory_mb_max": abi.PM_O_MEM_MAX,
           "total_memory_min": abi.PM_O_TOT_MIN, "total_memory_max": abi.PM_O_TOT_MAX}


@pytest.mark.parametrize("name,line,s,want", kv.PARSE_STRUCT, ids=[p[0] for p in kv.PARSE_STRUCT])
def test_parser_structure(name, line, s, want):
    r = orc.Req(s)
    assert r.n_gpu() == len(want["gpu"])
    for i, g in enumerate(want["gpu"]):
        for f in kv.GPU_FIELDS:
            assert r.gpu(i, f) == g.get(f), (i, f)
    assert r.scalar("ram_mb") == want["ram_mb"]
    assert r.scalar("storage_gb") == want["storage_gb"]
    assert (r.scalar("cpu_cores") if r.has_cpu() else None) == want["cpu"]
    # product parser: same structure in table form
    it = Interner()
    ask, opts = it.parse(s)
    assert int(ask["n_opts"]) == len(want["gpu"]) and len(opts) == len(want["gpu"])
    for i, g in enumerate(want["gpu"]):
        present = sum(PRESENT[f] for f in g)
        assert int(opts[i]["present"]) == present
        for f, v in g.items():
            if f == "model":
                assert int(opts[i]["pattern_id"]) == it.pattern(v)
            else:
                assert int(opts[i][f]) == v
    flags = int(ask["flags"])
    assert bool(flags & abi.PM_A_REQ_RAM) == (want["ram_mb"] is not None)
    assert bool(flags & abi.PM_A_REQ_STORAGE) == (want["storage_gb"] is not None)
    assert bool(flags & abi.PM_A_REQ_CPU) == (want["cpu"] is not None)
    if want["ram_mb"] is not None:
        assert int(ask["ram_mb"]) == want["ram_mb"]
    if want["cpu"] is not None:
        assert int(ask["cpu_cores"]) == want["cpu"]


MODEL_PAIRS = [
    # (spec model, requirement model list, expected) — node.rs:463-484 and the tests above
    ("nvidia_a100_80gb_pcie", "A100", True), ("NVIDIA A100 80GB", "A100", True), ("RTX 3090", "A100", False),
    ("rtx4090", "RTX4090", True), ("RTX 4090", "RTX4090", True), ("RTX 4090", "rtx_4090", True),
    ("AMD Radeon RX 7900", "nvidia,rtx", False), ("NVIDIA H100", "a100,h100,h200", True),
    ("A100", "NVIDIA A100 80GB", True),   # requirement contains the spec
    ("anything", "", True),               # empty requirement model matches any Some(model)
    ("H100", " h100 , a100 ", True), ("nvidia rtx a6000", "a6000", True), ("Tesla V100-SXM2-32GB", "v100,a100", True),
    ("NVIDIA L40S", "l40", True), ("NVIDIA L40S", "l4 0", True),  # underscores are stripped: "l4_0" -> "l40"
    ("NVIDIA L40S", "l4x0", False),
]


@pytest.mark.parametrize("spec,req,expected", MODEL_PAIRS)
def test_model_clause(spec, req, expected):
    assert orc.model_matches(spec, req) is expected
    it = Interner()
    m, p = it.model(spec), it.pattern(req)
    bits, npat, nmod, words = it.table()
    assert bool((bits[p * words + (m >> 5)] >> (m & 31)) & 1) is expected


def test_interner_matches_oracle_on_catalogue():
    from protocol_b200 import synth

    it = Interner()
    models = [m for m, _ in synth.MODEL_CATALOGUE]
    pats = synth.PATTERN_CATALOGUE
    for m in models:
        it.model(m)
    for p in pats:
        it.pattern(p)
    bits, npat, nmod, words = it.table()
    assert (npat, nmod) == (len(pats), len(models))
    for pi, p in enumerate(pats):
        for mi, m in enumerate(models):
            assert bool((bits[pi * words + (mi >> 5)] >> (mi & 31)) & 1) == orc.model_matches(m, p), (p, m)


def test_compat_none_cases():
    """is_node_compatible_with_config, mod.rs:206-215."""
    no_specs = orc.make_node(specs=False)
    with_specs = spec_to_orc_node(kv.specs(1, "A100", 40000))
    req = orc.Req("gpu:count=1")
    assert orc.node_compatible(no_specs, None) is True      # (None, _) => true
    assert orc.node_compatible(with_specs, None) is True
    assert orc.node_compatible(no_specs, req) is False      # (Some, None) => false
    assert orc.node_compatible(with_specs, req) is True


def test_count_none_equals_zero_rule():
    """node.rs:447-461: spec count None is ok only for req count 0."""
    n_none = spec_to_orc_node(kv.specs(None, "A100", 40000))
    n_zero = spec_to_orc_node(kv.specs(0, "A100", 40000))
    for node in (n_none, n_zero):
        assert orc.meets(node, orc.Req("gpu:count=0")) is True
        assert orc.meets(node, orc.Req("gpu:count=1")) is False


def test_total_memory_wraps_like_release_build():
    """node.rs:509: u32 multiply; release builds wrap (prod-release.yml builds --release)."""
    node = spec_to_orc_node(kv.specs(65536, "X", 65536))      # 2^32 -> 0
    assert orc.meets(node, orc.Req("gpu:total_memory_max=10")) is True
    assert orc.meets(node, orc.Req("gpu:total_memory_min=1")) is False


def test_newest_task_plugin():
    """newest_task/mod.rs:29-54 and max_by_key's last-maximum rule."""
    assert orc.newest_task([1, 2]) == 1
    assert orc.newest_task([]) == orc.PM_NONE
    assert orc.newest_task([5, 5, 3]) == 1
    # get_all_tasks sorts created_at desc (task_store.rs:79) before the plugin sees the list
    order = orc.sort_tasks([1, 7, 7, 3])
    assert order == [1, 2, 3, 0]


def test_node_sorting():
    """node_store.rs:419-465: Dead, Discovered, Healthy -> Healthy, Discovered, Dead."""
    assert orc.sort_nodes_by_status([orc.DEAD, orc.DISCOVERED, orc.HEALTHY]) == [2, 1, 0]
    st = [orc.UNHEALTHY, orc.HEALTHY, orc.EJECTED, orc.DEAD, orc.HEALTHY, orc.DISCOVERED, orc.BANNED]
    assert orc.sort_nodes_by_status(st) == [1, 4, 5, 0, 2, 6, 3]


def test_get_idx_in_group():
    """node_groups/tests.rs:1381-1445: BTreeSet<String> order of the addresses."""
    a1 = "0x1234567890123456789012345678901234567890"
    a2 = "0x2234567890123456789012345678901234567890"
    a3 = "0x3234567890123456789012345678901234567890"
    assert orc.idx_in_group([a2, a1], a1) == 0
    assert orc.idx_in_group([a2, a1], a2) == 1
    assert orc.idx_in_group([a2, a1], a3) == -1
    # EIP-55 mixed case: 'A'..'F' sort before 'a'..'f' (byte order)
    assert orc.idx_in_group(["0xaB", "0xAb", "0xab"], "0xAb") == 0
    assert orc.idx_in_group(["0xaB", "0xAb", "0xab"], "0xab") == 2


def test_config_priority_sort():
    """mod.rs:150-164 on the configs of tests.rs:1630-1800 style setups + product pm_sort_configs."""
    r = orc.Req("gpu:count=1")
    cfgs = [("general-1", 1, 1, None), ("gpu-2", 2, 4, r), ("general-2", 2, 2, None), ("gpu-1", 1, 1, r),
            ("big", 4, 8, None)]
    perm = orc.sort_configs(cfgs)
    assert [cfgs[i][0] for i in perm] == ["big", "gpu-2", "general-2", "gpu-1", "general-1"]
    p2 = sort_configs([c[1] for c in cfgs], [c[3] is not None for c in cfgs])
    assert p2.tolist() == perm
    sorted_cfgs = [cfgs[i] for i in perm]
    avail = orc.available_configs(sorted_cfgs, [1, 0, 1, 1, 0])
    assert [sorted_cfgs[i][0] for i in avail] == ["big", "general-2", "gpu-1"]


def test_haversine_reference_cities():
    """mod.rs:218-231 with the coordinates of tests.rs:2920-2989."""
    d = orc.haversine_km(45.5186, -73.5545, 32.7942, -96.7475)
    assert 2400.0 < d < 2470.0
    assert orc.haversine_km(10.0, 20.0, 10.0, 20.0) == 0.0
    # pure-python restatement with the same operation order agrees to the last bit
    import math

    def hav(lat1, lon1, lat2, lon2):
        k = math.pi / 180.0
        a = math.sin((lat2 - lat1) * k / 2.0) ** 2 if False else None
        s1 = math.sin(((lat2 - lat1) * k) / 2.0)
        s2 = math.sin(((lon2 - lon1) * k) / 2.0)
        a = s1 * s1 + math.cos(lat1 * k) * math.cos(lat2 * k) * (s2 * s2)
        return 6371.0 * (2.0 * math.atan2(math.sqrt(a), math.sqrt(1.0 - a)))

    assert hav(45.5186, -73.5545, 32.7942, -96.7475) == d


# ---- node ids: Address::from_str + to_string (alloy-primitives 1.1.0; not under /root/reference) -----------------------
EIP55_VECTORS = [   # the test vectors of EIP-55 itself
    "0x52908400098527886E0F7030069857D2E4169EE7", "0x8617E340B3D01FA5F11F306F4090FD50E238070D",   # all caps
    "0xde709f2102306220921060314715629080e2fb77", "0x27b1fdb04752bbc536007a920d24acb045561c26",   # all lower
    "0x5aAeb6053F3E94C9b9A09f33669435E7Ef1BeAed", "0xfB6916095ca1df60bB79Ce92cE3Ea74c37c5d359",
    "0xdbF03B407c01E7cD3CBea99509d93f8DDDC8C6FB", "0xD1220A0cf47c7B9Be7A2E6BA89F429762e7b9aDb",
]


def _product_canonical(address: str):
    import ctypes as C

    from protocol_b200 import _lib
    out = C.create_string_buffer(43)
    return out.value.decode() if _lib.load().pm_address_canonical(address.encode(), out) == 0 else None


def test_keccak256_and_eip55_known_answers():
    """The oracle's Keccak-256 against the well-known digests, its EIP-55 casing against the vectors of the EIP, and the
    product's pm_address_canonical against both."""
    assert orc.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert orc.keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    for v in EIP55_VECTORS:
        for spelled in (v, v.lower(), "0x" + v[2:].upper(), v[2:], "0X" + v[2:].lower()):
            assert orc.eip55(spelled) == v
            assert _product_canonical(spelled) == v
    # the addresses the reference's own tests use (node_groups/tests.rs, monitor.rs): to_string() of a parsed Address
    assert _product_canonical("0x0000000000000000000000000000000000000000") == "0x0000000000000000000000000000000000000000"


def test_address_canonical_matches_oracle_on_random_addresses_and_rejects_non_addresses():
    import random
    rng = random.Random(55)
    for _ in range(3000):
        a = "0x" + "".join(rng.choice("0123456789abcdefABCDEF") for _ in range(40))
        assert _product_canonical(a) == orc.eip55(a) and orc.eip55(a) is not None
    for bad in ("", "0x", "0x123", "0x" + "g" * 40, "0x" + "1" * 39, "0x" + "1" * 41, "node-1", "0x" + "1" * 20 + " " + "1" * 19):
        assert orc.eip55(bad) is None and _product_canonical(bad) is None
