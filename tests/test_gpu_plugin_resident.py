"""The management pass with RESIDENT tables (pm_plugin_try_form_new_groups keeps the worker table on the engine between
passes and uploads only rows touched since the previous one — pm_resize_workers / pm_update_workers) must publish
exactly the groups of a pass that re-uploads everything: a random history of node arrivals, status changes, deaths
(dissolution), merges and passes runs twice — once undisturbed, once on an engine whose table version is bumped before
every pass so that each pass takes the full-upload path — and the published groups are compared after every step.
Also: the engine-level delta calls against a plain upload, and the ADVICE r1 case of a merge pass that meets a solo
group whose node is not in the node table."""
import numpy as np
import pytest

from helpers import groups_equal, load_engine
from oracle import pm_oracle as orc
from protocol_b200 import abi, synth
from protocol_b200.engine import Engine
from protocol_b200.plugin import (ComputeSpecs, GpuSpecs, NodeGroupConfiguration, NodeGroupsPlugin, NodeStatus,
                                  OrchestratorNode, Task)

pytestmark = pytest.mark.gpu

MODELS = ["NVIDIA H100", "NVIDIA A100 80GB", "RTX 4090"]
CITIES = [(45.5186, -73.5545), (32.7942, -96.7475), (48.85, 2.35), None]


def addr(i):
    return "0x" + f"{(i * 2654435761) & 0xFFFFFFFF:08x}".upper() + f"{i:032x}"


def node(i, rng):
    spec = ComputeSpecs(gpu=GpuSpecs(count=int(rng.choice([1, 8])), model=MODELS[int(rng.integers(0, 3))], memory_mb=80000),
                        ram_mb=64000, storage_gb=1000) if rng.random() > 0.1 else None
    loc = CITIES[int(rng.integers(0, 4))]
    return OrchestratorNode(addr(i), status=NodeStatus.Healthy if rng.random() > 0.15 else NodeStatus.Discovered,
                            p2p_id=f"p2p-{i}" if rng.random() > 0.05 else None, compute_specs=spec, location=loc)


@pytest.mark.parametrize("proximity", [False, True], ids=["first_fit", "proximity"])
def test_resident_pass_equals_full_upload_pass(proximity):
    cfgs = [NodeGroupConfiguration("h100-trio", 3, 3, "gpu:count=8;gpu:model=H100"),
            NodeGroupConfiguration("pair", 2, 2, "gpu:count=8"), NodeGroupConfiguration("flex", 1, 4)]
    e1, e2 = Engine(), Engine()
    a = NodeGroupsPlugin(cfgs, engine=e1, proximity_enabled=proximity)
    b = NodeGroupsPlugin(cfgs, engine=e2, proximity_enabled=proximity)
    task = Task(allowed_topologies=["pair", "flex", "h100-trio"])   # one task (one id) known to both plugins
    for p in (a, b):
        p.add_task(task)
    rng = np.random.default_rng(17)
    n = 0
    for step in range(40):
        ops = []
        for _ in range(int(rng.integers(0, 25))):
            ops.append(("add", n, int(rng.integers(0, 1 << 30))))
            n += 1
        for _ in range(int(rng.integers(0, 6))):
            if n:
                ops.append(("status", int(rng.integers(0, n)), int(rng.choice([NodeStatus.Healthy, NodeStatus.Dead, NodeStatus.Unhealthy]))))
        for p in (a, b):
            for op in ops:
                if op[0] == "add":
                    p.add_node(node(op[1], np.random.default_rng(op[2])))
                else:
                    p.update_node_status(addr(op[1]), op[2])
        # b's engine is touched from outside before every pass: its plugin has to take the full-upload path
        e2.resize_workers(0)
        fa, fb = a.try_form_new_groups(), b.try_form_new_groups()
        assert fa == fb, f"step {step}: {fa} vs {fb} groups formed"
        if step % 3 == 2:
            assert a.try_merge_solo_groups() == b.try_merge_solo_groups()
        ga, gb = a.get_all_groups(), b.get_all_groups()
        assert ga == gb, f"step {step}: published groups differ"
    assert len(a.get_all_groups()) > 20
    for p in (a, b):
        p.close()
    e1.close()
    e2.close()


def test_delta_calls_equal_a_plain_upload():
    w = synth.make_workers(5000)
    a = synth.make_asks(200, "mixed", group_sizes=[(1, 1), (2, 3)])
    bits, npat, nmod, words = synth.intern_tables(w, a)
    t = dict(asks=a.asks, opts=a.opts, wa=w.a, wb=w.b, bits=bits, n_patterns=npat, n_models=nmod, words=words, lat=w.lat, lon=w.lon)
    og = orc.soa_form_groups(w.a, w.b, a.asks, a.opts, bits, words, lat=w.lat, lon=w.lon, proximity=True)
    eng = Engine()
    # start from a table of 3000 rows with other contents, grow to 5000, then overwrite every row through the delta call
    other = synth.make_workers(3000, seed=99)
    t0 = dict(t, wa=other.a, wb=other.b, lat=other.lat, lon=other.lon)
    load_engine(eng, t0, locations=True)
    eng.match(abi.PM_MODE_PROXIMITY)
    eng.resize_workers(5000)
    perm = np.random.default_rng(0).permutation(5000).astype(np.uint32)
    for lo in range(0, 5000, 1700):
        sel = perm[lo:lo + 1700]
        eng.update_workers(sel, w.a[sel], w.b[sel], w.lat[sel], w.lon[sel])
    eng.match(abi.PM_MODE_PROXIMITY)
    assert groups_equal(eng.fetch(), og)
    # shrink: rows past the end disappear, the kept rows stay
    eng.resize_workers(2500)
    eng.match(abi.PM_MODE_PROXIMITY)
    og2 = orc.soa_form_groups(w.a[:2500], w.b[:2500], a.asks, a.opts, bits, words, lat=w.lat[:2500], lon=w.lon[:2500], proximity=True)
    assert groups_equal(eng.fetch(), og2)
    eng.close()


def test_merge_pass_skips_solo_groups_of_unstored_nodes():
    """a restored solo group may name a node that is not in the node table yet; find_compatible_solo_groups
    treats it as compatible with nothing (mod.rs:733-741) — the merge pass must not fail, and must still merge the rest."""
    eng = Engine()
    p = NodeGroupsPlugin([NodeGroupConfiguration("merge-config", 1, 3)], engine=eng, proximity_enabled=False)
    p.add_task(Task(allowed_topologies=["merge-config"]))
    for i in range(3):
        p.add_node(OrchestratorNode(addr(i), p2p_id=f"p{i}"))
    p.restore_group("a1", "merge-config", [addr(0)])
    p.restore_group("a2", "merge-config", [addr(1)])
    p.restore_group("a3", "merge-config", [addr(77)])       # node 77 was never stored
    assert p.try_merge_solo_groups() == 1
    groups = {g["id"]: g for g in p.get_all_groups()}
    assert "a3" in groups and groups["a3"]["nodes"] == [addr(77)]
    merged = [g for g in groups.values() if len(g["nodes"]) == 2]
    assert len(merged) == 1 and sorted(merged[0]["nodes"]) == sorted([addr(0), addr(1)])
    p.close()
    eng.close()
