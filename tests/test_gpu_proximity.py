"""Group formation under the reference's DEFAULT policy, ProximityOptimizationPolicy{enabled:true}
(crates/orchestrator/src/plugins/node_groups/mod.rs:85-89, 524-552, 218-255), on the all-SM cooperative sweep
(pm_proximity_grid.cuh, the default) and on the single-CTA sweep it replaced (PM_TUNE_PROX=2): groups, creation order
and member order equal the CPU checker's on every shape the sweep has a special path for — ties everywhere (city
clusters, one point), scattered coordinates, nobody located, groups larger than one selection pass (16), empty
groups (min 0), long hand-downs between configurations (ordered compaction), solo configurations after grouped ones.
"""
import os

import numpy as np
import pytest

from helpers import groups_equal, load_engine
from oracle import pm_oracle as orc
from protocol_b200 import abi, synth
from protocol_b200.engine import Engine, Interner

pytestmark = pytest.mark.gpu

KERNELS = [("0", "grid"), ("2", "single_cta")]


def swarm(n_workers, where, seed_shift=0):
    w = synth.make_workers(n_workers, seed=synth.SEED_WORKERS + seed_shift, with_addresses=n_workers <= 30000)
    rng = np.random.default_rng(seed_shift + 1)
    lat, lon = w.lat.copy(), w.lon.copy()
    if where == "scattered":      # distinct points: no near-ties between different libm's at kilometre scale
        lat = lat + rng.normal(0, 3.0, len(lat)).clip(-20, 20)
        lon = lon + rng.normal(0, 5.0, len(lon))
    elif where == "one_point":
        lat[:] = 45.5
        lon[:] = -73.5
    elif where == "nowhere":
        w.a["flags"] &= ~np.uint32(abi.PM_W_HAS_LOC)
    else:
        assert where == "cities"
    w.lat, w.lon = lat, lon
    return w


def tables_for(w, a):
    bits, npat, nmod, words = synth.intern_tables(w, a)
    return dict(asks=a.asks, opts=a.opts, wa=w.a, wb=w.b, bits=bits, n_patterns=npat, n_models=nmod, words=words,
                lat=w.lat, lon=w.lon)


def custom_asks(rows):
    """[(requirement string | None, min, max)] in priority order -> synth.Asks."""
    it = Interner()
    for i, m in enumerate(m for m, _ in synth.MODEL_CATALOGUE):
        assert it.model(m) == i
    for i, p in enumerate(synth.PATTERN_CATALOGUE):
        assert it.pattern(p) == i
    asks, opts = [], []
    for req, mn, mx in rows:
        if req is None:
            row, o = np.zeros(1, dtype=abi.ASK)[0], np.zeros(0, dtype=abi.GPU_OPT)
        else:
            row, o = it.parse(req)
        row = row.copy()
        row["opt_off"] = sum(len(x) for x in opts)
        row["min_group_size"], row["max_group_size"] = mn, mx
        asks.append(row)
        opts.append(o)
    it.close()
    allo = np.concatenate(opts).astype(abi.GPU_OPT) if sum(map(len, opts)) else np.zeros(0, dtype=abi.GPU_OPT)
    return synth.Asks(asks=np.array(asks, dtype=abi.ASK), opts=allo, pattern_strings=list(synth.PATTERN_CATALOGUE))


def run(kernel, w, a, checker="banded"):
    t = tables_for(w, a)
    os.environ["PM_TUNE_PROX"] = kernel   # read at pm_create
    try:
        eng = Engine(timing=True)
    finally:
        os.environ.pop("PM_TUNE_PROX", None)
    load_engine(eng, t, addr_rank=w.addr_rank, locations=True)
    eng.match(abi.PM_MODE_PROXIMITY)
    res = eng.fetch()
    eng.match(abi.PM_MODE_PROXIMITY)            # a second pass on the same engine reuses every buffer
    res2 = eng.fetch()
    eng.close()
    og = orc.soa_form_groups(t["wa"], t["wb"], t["asks"], t["opts"], t["bits"], t["words"], addr_rank=w.addr_rank,
                             lat=t["lat"], lon=t["lon"], proximity=checker)
    assert groups_equal(res, og), f"groups differ: engine {res.n_groups} vs checker {len(og)}"
    assert groups_equal(res2, og)
    wg = np.full(len(w), abi.PM_NONE, dtype=np.uint32)
    for g in range(len(og.cfg)):
        wg[og.members[og.off[g]:og.off[g + 1]]] = g
    assert np.array_equal(res.worker_group, wg)
    return res, og


@pytest.mark.parametrize("kernel", [k for k, _ in KERNELS], ids=[n for _, n in KERNELS])
@pytest.mark.parametrize("where", ["cities", "scattered", "one_point", "nowhere"])
def test_mixed_group_sizes(kernel, where):
    sizes = [(1, 1), (2, 2), (2, 4), (3, 3), (4, 8), (1, 3), (0, 2), (2, 5), (0, 1)]
    w = swarm(4000, where, seed_shift=7)
    a = synth.make_asks(120, "mixed", seed=synth.SEED_ASKS + 7, group_sizes=sizes)
    res, og = run(kernel, w, a, checker=True if where == "nowhere" else "banded")
    assert res.n_groups > 50


@pytest.mark.parametrize("kernel", [k for k, _ in KERNELS], ids=[n for _, n in KERNELS])
@pytest.mark.parametrize("where", ["cities", "scattered"])
def test_groups_larger_than_one_selection_pass(kernel, where):
    """max - 1 > 16 neighbours: several selection passes (and barriers) per group; min close to max leaves tails."""
    w = swarm(6000, where, seed_shift=3)
    a = custom_asks([("gpu:count=8", 30, 40), ("gpu:count=4", 17, 17), ("gpu:count=2", 18, 35), (None, 20, 33), (None, 2, 19)])
    res, og = run(kernel, w, a)
    sizes = np.diff(res.group_off)
    assert sizes.max() >= 35 and res.stats["n_bumped"] > 0


@pytest.mark.parametrize("kernel", [k for k, _ in KERNELS], ids=[n for _, n in KERNELS])
def test_long_hand_down_between_configurations(kernel):
    """A configuration that cannot reach its minimum hands ALL its candidates down (hundreds: the list of the next
    configuration is rebuilt by ordered compaction), a pair configuration leaves one, and a solo configuration with
    min 0 takes the rest and adds its empty group."""
    w = swarm(9000, "cities", seed_shift=11)
    n8 = int(((w.a["gpu_count"] == 8) & ((w.a["flags"] & 3) == 3) & ((w.a["flags"] & abi.PM_W_HAS_SPECS) != 0)).sum())
    assert n8 > 600
    a = custom_asks([("gpu:count=8", n8 + 50, n8 + 60),      # never forms: every 8-GPU candidate is handed down
                     ("gpu:count=8;gpu:model=H100", 3, 3),
                     ("gpu:count=8", 2, 2),
                     ("gpu:count=4", 2, 7),
                     (None, 0, 1)])
    res, og = run(kernel, w, a)
    assert res.stats["n_bumped"] >= n8 and (np.diff(res.group_off) == 0).sum() == 1


@pytest.mark.parametrize("kernel", [k for k, _ in KERNELS], ids=[n for _, n in KERNELS])
def test_pairs_then_solo_at_100k_nodes(kernel):
    """The shape of profiles/r01_host_paths.txt (pair-h100 2..2 + solo 1..1, default policy) at 100 000 nodes."""
    w = swarm(100_000, "cities", seed_shift=0)
    a = custom_asks([("gpu:count=8;gpu:count=4", 2, 2), (None, 1, 1)])   # 35 % of the candidates compete for pairs
    res, og = run(kernel, w, a)
    print(f"\n[proximity {dict(KERNELS)[kernel]}] 100k nodes: {res.n_groups} groups, ms_resolve = {res.stats['ms_resolve']:.2f}, "
          f"batches = {res.stats['n_tiles']}, groups formed in batches = {res.stats['n_rounds']}, one at a time = {res.stats['n_build_launches']}")
    assert (np.diff(res.group_off) == 2).sum() > 10_000


@pytest.mark.skipif(os.environ.get("PM_TEST_BIG") != "1", reason="1M-node timing run (PM_TEST_BIG=1): ~1 min of CPU checker")
def test_pairs_then_solo_at_1m_nodes():
    """1M nodes, every fourth an 8-GPU node competing for pair groups (the 40 s case of round 1)."""
    w = swarm(1_000_000, "cities", seed_shift=0)
    a = custom_asks([("gpu:count=8;gpu:count=4", 2, 2), (None, 1, 1)])
    res, og = run("0", w, a)
    print(f"\n[proximity grid] 1M nodes: {res.n_groups} groups ({int((np.diff(res.group_off) == 2).sum())} pairs), "
          f"ms_resolve = {res.stats['ms_resolve']:.1f}, batches = {res.stats['n_tiles']}, groups formed in batches = {res.stats['n_rounds']}, "
          f"one at a time = {res.stats['n_build_launches']}")
