"""Allocator properties ported from the reference's node_groups/tests.rs onto the oracle
(orc.form_groups = faithful try_form_new_groups, mod.rs:478-628), and equivalence of the
SoA restatement (orc.soa_form_groups, the large-size checker) with the faithful loop on
randomised inputs.  CPU only."""
import numpy as np
import pytest

import kat_vectors as kv
from helpers import TableBuilder, spec_to_orc_node
from oracle import pm_oracle as orc
from protocol_b200 import abi, synth

A1 = "0x1234567890123456789012345678901234567890"
A2 = "0x2234567890123456789012345678901234567890"
A3 = "0x3234567890123456789012345678901234567890"
RTX = kv.specs(8, "RTX4090", 24)


def plain(addr, status=orc.HEALTHY, **kw):
    return orc.make_node(address=addr, status=status, specs=False, **kw)


def sizes(g):
    return sorted(len(m) for _, m in g.as_list())


def test_group_formation_and_dissolution_basics():
    """tests.rs:105-196: min=max=2, one node -> no group; two nodes -> one group; a Dead node is not
    a candidate on the next pass."""
    cfg = [("test-config", 2, 2, None)]
    assert len(orc.form_groups([plain(A1)], cfg, True)) == 0
    g = orc.form_groups([plain(A1), plain(A2)], cfg, True)
    assert g.as_list() == [(0, [0, 1])]
    g = orc.form_groups([plain(A1, status=orc.DEAD), plain(A2)], cfg, True)
    assert len(g) == 0


def test_requirements_single_node():
    """tests.rs:303-384."""
    req = orc.Req("gpu:count=8;gpu:model=RTX4090;")
    cfg = [("test-config-with-requirements", 1, 1, req)]
    assert len(orc.form_groups([plain(A1)], cfg, True)) == 0           # compute_specs None
    g = orc.form_groups([plain(A1), spec_to_orc_node(RTX, address=A2)], cfg, True)
    assert g.as_list() == [(0, [1])]


def test_requirements_multiple_nodes():
    """tests.rs:387-506: min=max=2 with requirements; the spec-less node never joins."""
    req = orc.Req("gpu:count=8;gpu:model=RTX4090;")
    cfg = [("c", 2, 2, req)]
    n1, n2, n3 = plain(A1), spec_to_orc_node(RTX, address=A2), spec_to_orc_node(RTX, address=A3)
    assert len(orc.form_groups([n1, n2], cfg, True)) == 0
    g = orc.form_groups([n1, n2, n3], cfg, True)
    assert g.as_list() == [(0, [1, 2])]


def test_max_size_any_two_of_three():
    """tests.rs:734-885: exactly one group of two; which two is unpinned by the reference
    (:846-861) — the determinisation rule (canonical order) picks the first two."""
    for prox in (False, True):
        g = orc.form_groups([plain(A1), plain(A2), plain(A3)], [("test-config", 2, 2, None)], prox)
        assert sizes(g) == [2]
        assert g.as_list() == [(0, [0, 1])]


def test_building_largest_possible_groups():
    """tests.rs:1630-1757: configs 1/2/3 -> three nodes land in ONE group of three (largest first)."""
    cfgs = [("small", 1, 1, None), ("medium", 2, 2, None), ("large", 3, 3, None)]
    order = orc.sort_configs(cfgs)
    sorted_cfgs = [cfgs[i] for i in order]
    assert [c[0] for c in sorted_cfgs] == ["large", "medium", "small"]
    g = orc.form_groups([plain(A1), plain(A2), plain(A3)], sorted_cfgs, True)
    assert g.as_list() == [(0, [0, 1, 2])]


def test_group_formation_priority_3_plus_1():
    """tests.rs:1803-1904: four nodes, configs {3,3} and {1,1}: one 3-group + one 1-group, not 4 solos."""
    cfgs = [("solo", 1, 1, None), ("trio", 3, 3, None)]
    sorted_cfgs = [cfgs[i] for i in orc.sort_configs(cfgs)]
    nodes = [plain(f"0x{i + 1}234567890123456789012345678901234567890") for i in range(4)]
    g = orc.form_groups(nodes, sorted_cfgs, True)
    assert sizes(g) == [1, 3]
    members = sorted(m for _, ms in g.as_list() for m in ms)
    assert members == [0, 1, 2, 3]                                      # exclusivity: each node once


def test_multiple_groups_same_configuration():
    """tests.rs:1907-2009: six nodes, min=max=2 -> exactly 3 groups of 2, 3 distinct ids."""
    nodes = [plain(f"0x{i + 1}234567890123456789012345678901234567890") for i in range(6)]
    g = orc.form_groups(nodes, [("pairs", 2, 2, None)], True)
    assert sizes(g) == [2, 2, 2]
    assert sorted(m for _, ms in g.as_list() for m in ms) == list(range(6))


def test_node_cannot_be_in_multiple_groups_and_assigned_are_skipped():
    """tests.rs:993-1212: nodes already present in node_to_group are not candidates."""
    nodes = [plain(A1, assigned=True), plain(A2), plain(A3)]
    g = orc.form_groups(nodes, [("c", 1, 2, None)], False)
    assert g.as_list() == [(0, [1, 2])]


def test_unhealthy_and_no_p2p_nodes_are_not_candidates():
    """mod.rs:492-497."""
    nodes = [plain(A1, status=orc.UNHEALTHY), plain(A2, p2p=False), plain(A3)]
    g = orc.form_groups(nodes, [("c", 1, 1, None)], False)
    assert g.as_list() == [(0, [2])]


MONTREAL, DALLAS = (45.5186, -73.5545), (32.7942, -96.7475)
M1, M2 = "0xB2631de00e6120969d34456b9c7Ee22352f13b02", "0x2C490CAdf3A8C2Ab67b00831973da8b9d18e5b6D"
D1, D2 = "0x7ec9d3bc276B74969341c03dc00B9f70c0EadFd5", "0x32d7cd9b8F6eA556a67E0c9386cdd911Da3AD3E5"
A6000 = kv.specs(1, "nvidia rtx a6000", 49140)


def test_proximity_never_mixes_montreal_and_dallas():
    """The coordinates and node specs of tests.rs:2861-3064, run through the formation pass with
    ProximityOptimizationPolicy enabled and min=max=2: interleaved arrival order still pairs by city."""
    nodes = [spec_to_orc_node(A6000, address=M1, location=MONTREAL), spec_to_orc_node(A6000, address=D1, location=DALLAS),
             spec_to_orc_node(A6000, address=M2, location=MONTREAL), spec_to_orc_node(A6000, address=D2, location=DALLAS)]
    g = orc.form_groups(nodes, [("2x40-48GB", 2, 2, None)], True)
    groups = [sorted(m) for _, m in g.as_list()]
    assert sorted(groups) == [[0, 2], [1, 3]]
    # without the policy the first two arrivals are paired regardless of distance
    g2 = orc.form_groups(nodes, [("2x40-48GB", 2, 2, None)], False)
    assert sorted(sorted(m) for _, m in g2.as_list()) == [[0, 1], [2, 3]]
    # members come out in BTreeSet<String> order of the address strings (mod.rs:63-69)
    for _, m in g.as_list():
        addrs = [[M1, D1, M2, D2][i] for i in m]
        assert addrs == sorted(addrs, key=lambda s: s.encode())


def test_seed_prefers_located_node_and_unlocated_sort_last():
    """mod.rs:526-530, 234-255."""
    nodes = [plain(A1), spec_to_orc_node(A6000, address=A2, location=DALLAS),
             spec_to_orc_node(A6000, address=A3, location=MONTREAL), plain("0x4234567890123456789012345678901234567890")]
    g = orc.form_groups(nodes, [("c", 3, 3, None)], True)
    # seed = node 1 (first with a location); nearest = node 2 (Montreal, located) before the unlocated
    # nodes 0 and 3 (distance f64::MAX, stable order) -> {1, 2, 0}
    assert [sorted(m) for _, m in g.as_list()] == [[0, 1, 2]]


def test_degenerate_sizes_follow_the_loop_literally():
    """min_group_size == 0 yields one trailing empty group per pass; max == 0 (first-fit) takes nobody
    but still creates that empty group (mod.rs:507,517,555-566,606)."""
    nodes = [plain(A1), plain(A2), plain(A3)]
    g = orc.form_groups(nodes, [("z", 0, 2, None)], False)
    assert [m for _, m in g.as_list()] == [[0, 1], [2], []]
    g = orc.form_groups(nodes, [("z", 0, 0, None)], False)
    assert [m for _, m in g.as_list()] == [[]]
    g = orc.form_groups(nodes, [("z", 0, 0, None)], True)    # the seed is inserted before the max check
    assert [m for _, m in g.as_list()] == [[0], [1], [2], []]


# ---------------------------------------------------------------------------------------------
def build_both(n_nodes, n_cfgs, seed, group_sizes, kind="mixed"):
    """Same synthetic swarm as (a) reference-shaped nodes+strings and (b) engine tables."""
    w = synth.make_workers(n_nodes, seed=seed, with_addresses=True)
    a = synth.make_asks(n_cfgs, kind, seed=seed + 1, group_sizes=group_sizes)
    bits, npat, nmod, words = synth.intern_tables(w, a)
    nodes = []
    for i in range(n_nodes):
        f = int(w.a["flags"][i])
        has = lambda b: bool(f & b)
        nodes.append(orc.make_node(
            address=w.addresses[i], status=int(w.status[i]), p2p=has(abi.PM_W_P2P), specs=has(abi.PM_W_HAS_SPECS),
            gpu_count=int(w.a["gpu_count"][i]) if has(abi.PM_W_HAS_GPU_COUNT) else None,
            gpu_model=w.model_strings[int(w.a["model_id"][i])] if has(abi.PM_W_HAS_GPU_MODEL) else None,
            gpu_mem=int(w.a["gpu_mem_mb"][i]) if has(abi.PM_W_HAS_GPU_MEM) else None,
            gpu_present=has(abi.PM_W_HAS_GPU), cpu_present=has(abi.PM_W_HAS_CPU),
            cpu_cores=int(w.b["cpu_cores"][i]) if has(abi.PM_W_HAS_CPU_CORES) else None,
            ram=int(w.b["ram_mb"][i]) if has(abi.PM_W_HAS_RAM) else None,
            storage=int(w.b["storage_gb"][i]) if has(abi.PM_W_HAS_STORAGE) else None,
            location=(float(w.lat[i]), float(w.lon[i])) if has(abi.PM_W_HAS_LOC) else None))
    reqs = [orc.Req(a.requirement_string(t)) for t in range(n_cfgs)]
    cfgs = [(f"cfg-{t}", int(a.asks["min_group_size"][t]), int(a.asks["max_group_size"][t]), reqs[t])
            for t in range(n_cfgs)]
    return w, a, (bits, words), nodes, cfgs


@pytest.mark.parametrize("seed", [11, 12, 13])
@pytest.mark.parametrize("proximity", [False, True], ids=["first_fit", "proximity"])
def test_soa_restatement_equals_faithful_loop(seed, proximity):
    """The SoA/interned allocator used as the large-size checker produces exactly the groups of the
    faithful string/AoS loop: same groups, same creation order, same BTreeSet member order."""
    sizes_ = [(1, 1), (2, 2), (2, 4), (3, 3), (1, 3), (4, 8)]
    w, a, (bits, words), nodes, cfgs = build_both(700, 60, seed, sizes_)
    # the faithful loop first applies get_nodes()' stable status sort; the SoA tables are given in that
    # order already only if statuses are grouped, so feed the SoA side the permuted table
    perm = np.array(orc.sort_nodes_by_status(w.status), dtype=np.int64)
    gf = orc.form_groups(nodes, cfgs, proximity)
    inv_rank = w.addr_rank[perm]
    gs = orc.soa_form_groups(w.a[perm], w.b[perm], a.asks, a.opts, bits, words, addr_rank=inv_rank,
                             lat=w.lat[perm], lon=w.lon[perm], proximity=proximity)
    assert np.array_equal(gf.cfg, gs.cfg)
    assert np.array_equal(gf.off, gs.off)
    assert np.array_equal(gf.members, perm[gs.members])     # map permuted rows back to input rows
    assert len(gf) > 20


def test_soa_predicate_equals_string_predicate_on_synthetic_pairs():
    w, a, (bits, words), nodes, cfgs = build_both(300, 80, 99, None)
    import ctypes as C
    for t in range(0, 80, 3):
        for i in range(0, 300, 7):
            s = orc.node_compatible(nodes[i], cfgs[t][3])
            v = orc.soa_compatible(w.a[i], w.b[i], a.asks[t], a.opts, bits, words)
            assert s == v, (t, i, a.requirement_string(t))


def test_faithful_loop_counts_the_quadratic_refilter():
    """forming G groups re-filters the remaining nodes G+1 times per configuration."""
    nodes = [plain(f"0x{i:040x}") for i in range(40)]
    g = orc.form_groups(nodes, [("solo", 1, 1, None)], False)
    assert len(g) == 40
    assert g.evals == sum(range(41))            # 40 + 39 + ... + 1 (+0 for the last, empty filter)


# ---------------------------------------------------------------------------------------------
# try_merge_solo_groups (mod.rs:631-971)
def test_proximity_merging_prevents_wrong_nodes_grouping():
    """tests.rs:2861-3064 verbatim: four solo groups (Montreal, Montreal, Dallas, Dallas in arrival
    order M1, M2, D1, D2) merge into exactly two 2-node groups, one per city."""
    nodes = [spec_to_orc_node(A6000, address=M1, location=MONTREAL), spec_to_orc_node(A6000, address=M2, location=MONTREAL),
             spec_to_orc_node(A6000, address=D1, location=DALLAS), spec_to_orc_node(A6000, address=D2, location=DALLAS)]
    cfgs = [("2x40-48GB", 2, 2, None), ("1x40-48GB", 1, 1, None)]          # priority order (min desc)
    for ids in (["1", "2", "3", "4"], ["9", "3", "7", "1"], ["a", "10", "2", "1f"]):   # any id order
        solos = [(ids[i], i, False) for i in range(4)]
        g = orc.merge_solo_groups(nodes, solos, cfgs, proximity=True)
        assert sorted(sorted(m) for _, m in g.as_list()) == [[0, 1], [2, 3]]
        assert all(c == 0 for c, _ in g.as_list())


def test_no_merge_when_policy_disabled_or_single_solo():
    """tests.rs:2636-2710 (TaskSwitchingPolicy.enabled = false) and mod.rs:640-644."""
    nodes = [plain(A1), plain(A2), plain(A3)]
    cfgs = [("c", 1, 3, None)]
    solos = [("1", 0, True), ("2", 1, True), ("3", 2, False)]
    assert len(orc.merge_solo_groups(nodes, solos, cfgs, task_switching_enabled=False)) == 0
    assert len(orc.merge_solo_groups(nodes, solos[:1], cfgs)) == 0
    g = orc.merge_solo_groups(nodes, solos, cfgs, proximity=False)
    assert g.as_list() == [(0, [0, 1, 2])]
    # prefer_larger_groups = false: a batch containing a group that holds a task is refused
    assert len(orc.merge_solo_groups(nodes, solos, cfgs, proximity=False, prefer_larger_groups=False)) == 0


def test_merge_only_compatible_groups():
    """tests.rs:2471-2635: only solo groups whose node meets the requirements are merged."""
    a100 = kv.specs(8, "A100", 80000)
    nodes = [spec_to_orc_node(a100, address=A1), plain(A2), spec_to_orc_node(a100, address=A3),
             spec_to_orc_node(kv.specs(8, "RTX 3090", 24000), address="0x4234567890123456789012345678901234567890")]
    req = orc.Req("gpu:count=8;gpu:model=A100")
    solos = [(str(i + 1), i, False) for i in range(4)]
    g = orc.merge_solo_groups(nodes, solos, [("a100-pair", 2, 2, req)], proximity=True)
    assert g.as_list() == [(0, [0, 2])]


def test_merge_chunks_and_leftovers():
    nodes = [plain(f"0x{i + 1:040x}") for i in range(7)]
    solos = [(f"{i + 1:x}", i, False) for i in range(7)]
    g = orc.merge_solo_groups(nodes, solos, [("trio", 3, 3, None), ("pair", 2, 2, None)], proximity=False)
    assert [(c, sorted(m)) for c, m in g.as_list()] == [(0, [0, 1, 2]), (0, [3, 4, 5])]   # the 7th stays solo
    g = orc.merge_solo_groups(nodes, solos, [("upto4", 1, 4, None)], proximity=False)
    assert [sorted(m) for _, m in g.as_list()] == [[0, 1, 2, 3], [4, 5, 6]]


@pytest.mark.parametrize("seed", [21, 22, 23, 24])
@pytest.mark.parametrize("where", ["cities", "scattered", "one_point"])
def test_latitude_pruned_proximity_equals_the_restated_loop(seed, where):
    """The checker's fast proximity mode (latitude-ordered candidates, a group looks outward from its seed only while
    R * |delta latitude| can still beat its k-th nearest) forms exactly the groups of the reference's loop restated
    literally (distances to every remaining worker, stable sort): same groups, creation order and members — with
    every worker of a city on one coordinate (all ties: canonical position decides), with scattered coordinates, and
    with everybody on one point; unlocated workers and degenerate sizes included."""
    sizes_ = [(1, 1), (2, 2), (2, 4), (3, 3), (1, 3), (4, 8), (0, 2), (0, 0), (2, 5)]
    w = synth.make_workers(3000, seed=seed, with_addresses=True)
    a = synth.make_asks(50, "mixed", seed=seed + 1, group_sizes=sizes_)
    bits, npat, nmod, words = synth.intern_tables(w, a)
    rng = np.random.default_rng(seed)
    lat, lon = w.lat.copy(), w.lon.copy()
    if where == "scattered":
        lat = lat + rng.normal(0, 3.0, len(lat)).clip(-20, 20)
        lon = lon + rng.normal(0, 5.0, len(lon))
        lat[rng.random(len(lat)) < 0.05] = 48.8566                     # and a cluster of exact latitude ties
    elif where == "one_point":
        lat[:] = 45.5
        lon[:] = -73.5
    args = (w.a, w.b, a.asks, a.opts, bits, words)
    want = orc.soa_form_groups(*args, addr_rank=w.addr_rank, lat=lat, lon=lon, proximity=True)
    got = orc.soa_form_groups(*args, addr_rank=w.addr_rank, lat=lat, lon=lon, proximity="banded")
    assert np.array_equal(want.cfg, got.cfg) and np.array_equal(want.off, got.off)
    assert np.array_equal(want.members, got.members)
    assert len(want) > 50
