"""The host mirror once groups exist — on the CPU: groups are put in with pm_plugin_restore_group (the start-up path:
what the reference left in Redis) instead of being formed on the GPU, and the scheduler / observer / recovery paths of
node_groups (scheduler_impl.rs:11-210, status_update_impl.rs:8-39, mod.rs:1046-1119, 1224-1325, 1423-1487) run
against them.  The same scenarios run with GPU-formed groups in test_gpu_plugin.py."""
import json
import threading

import pytest

from protocol_b200 import abi
from protocol_b200._lib import PrimeMatchError
from protocol_b200.plugin import NodeGroupConfiguration, NodeGroupsPlugin, NodeStatus, OrchestratorNode, Scheduler, Task

A1 = "0x1234567890123456789012345678901234567890"
A2 = "0x2234567890123456789012345678901234567890"
A3 = "0x3234567890123456789012345678901234567890"


def make(configs, **kw):
    return NodeGroupsPlugin(configs, **kw)


def test_restored_group_is_scheduled_like_a_formed_one():
    """tests.rs:509-676 with the group restored: same task for both members, GROUP_INDEX by BTreeSet order, variables."""
    plugin = make([NodeGroupConfiguration("test-config", 2, 2)])
    sched = Scheduler(plugin)
    plugin.add_node(OrchestratorNode(A1, p2p_id="p2p-1"))
    plugin.add_node(OrchestratorNode(A2, p2p_id="p2p-2"))
    env = {"LOCAL_RANK": "0", "RANK": "${GROUP_INDEX}", "WORLD_SIZE": "${GROUP_SIZE}", "GROUP_ID": "${GROUP_ID}",
           "NEXT": "${NEXT_P2P_ADDRESS}", "TOTAL_UPLOAD_COUNT": "${TOTAL_UPLOAD_COUNT}", "LAST_FILE_IDX": "${LAST_FILE_IDX}"}
    task = Task(image="prime-vllm", name="test-task", env_vars=env, cmd=["run", "--rank", "${GROUP_INDEX}.${GROUP_SIZE}"],
                allowed_topologies=["test-config"])
    plugin.add_task(task)
    assert sched.get_task_for_node(A1) is None                          # not in a group yet
    plugin.restore_group("2a", "test-config", [A1, A2])
    g = plugin.get_node_group(A2)
    assert g == {"id": "2a", "nodes": [A1, A2], "configuration_name": "test-config", "task_id": None}
    plugin.record_upload(A1, "2a", "f0")
    plugin.record_upload(A1, "2a", "f1")
    t1, t2 = sched.get_task_for_node(A1), sched.get_task_for_node(A2)
    assert t1["id"] == t2["id"] == task.id
    e1, e2 = t1["env_vars"], t2["env_vars"]
    assert (e1["GROUP_INDEX"], e1["RANK"], e1["WORLD_SIZE"], e1["GROUP_ID"], e1["NEXT"]) == ("0", "0", "2", "2a", "p2p-2")
    assert (e2["GROUP_INDEX"], e2["RANK"], e2["WORLD_SIZE"], e2["GROUP_ID"], e2["NEXT"]) == ("1", "1", "2", "2a", "p2p-1")
    assert (e1["TOTAL_UPLOAD_COUNT"], e1["LAST_FILE_IDX"]) == ("2", "1") and (e2["TOTAL_UPLOAD_COUNT"], e2["LAST_FILE_IDX"]) == ("0", "0")
    assert t1["cmd"] == ["run", "--rank", "0.2"] and t2["cmd"] == ["run", "--rank", "1.2"]
    assert plugin.get_node_group(A1)["task_id"] == task.id              # the claim (SET NX) is on the group now


def test_restore_validates_and_keeps_the_id_counter_ahead():
    plugin = make([NodeGroupConfiguration("c", 1, 2)])
    for a in (A1, A2, A3):
        plugin.add_node(OrchestratorNode(a))
    plugin.restore_group("ff", "c", [A1], task_id="some-task-that-no-longer-exists")
    with pytest.raises(PrimeMatchError) as e:
        plugin.restore_group("ff", "c", [A2])                           # id taken
    assert e.value.status == abi.PM_E_STATE
    with pytest.raises(PrimeMatchError) as e:
        plugin.restore_group("100", "c", [A1])                          # node already grouped
    assert e.value.status == abi.PM_E_STATE
    with pytest.raises(PrimeMatchError):
        plugin.restore_group("101", "c", [A2, A2])
    plugin.restore_group("not-hex-id", "c", [A2, A3], created_at_ms=1_700_000_000_250)
    assert [g["id"] for g in plugin.get_all_groups()] == ["ff", "not-hex-id"]
    # a claim on a task that is gone is dropped at the first look (get_current_group_task, mod.rs:436-469)
    assert Scheduler(plugin).get_task_for_node(A1) is None
    assert plugin.get_node_group(A1)["task_id"] is None
    cmds = plugin.redis_writeback()
    stored = json.loads({(c[0], c[1]): c[2:] for c in cmds}[("SET", "node_group:not-hex-id")][0])
    assert stored["nodes"] == [A2, A3] and stored["created_at"] == "2023-11-14T22:13:20.250Z"


def test_writeback_then_restore_round_trip():
    """What pm_plugin_redis_writeback emits is what a fresh plugin needs at start-up."""
    src = make([NodeGroupConfiguration("c2", 2, 2), NodeGroupConfiguration("c1", 1, 1)])
    for a in (A1, A2, A3):
        src.add_node(OrchestratorNode(a))
    task = Task(allowed_topologies=["c2", "c1"])
    src.add_task(task)
    src.restore_group("1", "c2", [A1, A2], task_id=task.id)
    src.restore_group("2", "c1", [A3])
    cmds = src.redis_writeback()
    dst = make([NodeGroupConfiguration("c2", 2, 2), NodeGroupConfiguration("c1", 1, 1)])
    for a in (A1, A2, A3):
        dst.add_node(OrchestratorNode(a))
    dst.add_task(task)
    claims = {c[1].split(":", 1)[1]: c[2] for c in cmds if c[0] == "SET" and c[1].startswith("group_task:")}
    for c in cmds:
        if c[0] == "SET" and c[1].startswith("node_group:"):
            g = json.loads(c[2])
            dst.restore_group(g["id"], g["configuration_name"], g["nodes"], task_id=claims.get(g["id"]))
    assert dst.get_all_groups() == src.get_all_groups()
    assert sorted(map(tuple, dst.redis_writeback())) == sorted(map(tuple, cmds))


def test_dissolution_paths_on_restored_groups():
    """handle_status_change (status_update_impl.rs:8-39) and on_task_deleted (mod.rs:1259-1320)."""
    plugin = make([NodeGroupConfiguration("c", 2, 2)])
    for a in (A1, A2):
        plugin.add_node(OrchestratorNode(a))
    task = Task(allowed_topologies=["c"])
    plugin.add_task(task)
    plugin.restore_group("1", "c", [A1, A2])
    assert Scheduler(plugin).get_task_for_node(A1)["id"] == task.id
    plugin.update_node_status(A1, NodeStatus.Dead)
    assert plugin.get_node_group(A1) is None and plugin.get_node_group(A2) is None and plugin.get_all_groups() == []
    plugin.update_node_status(A1, NodeStatus.Healthy)
    plugin.restore_group("2", "c", [A1, A2])
    assert Scheduler(plugin).get_task_for_node(A2)["id"] == task.id
    plugin.delete_task(task.id)                                          # the group working on it goes with it
    assert plugin.get_all_groups() == []


def test_group_recovery_entry_points_on_restored_groups():
    """get_group_by_id / validate_group_exists / handle_group_not_found (mod.rs:1046-1119)."""
    plugin = make([NodeGroupConfiguration("c", 1, 1)])
    for a in (A1, A2):
        plugin.add_node(OrchestratorNode(a))
    busy, idle = Task(name="busy", allowed_topologies=["c"]), Task(name="orphan", allowed_topologies=["c"])
    plugin.add_task(busy)
    plugin.add_task(idle)
    plugin.restore_group("1", "c", [A1], task_id=busy.id)
    plugin.restore_group("2", "c", [A2])
    assert plugin.validate_group_exists("1") and not plugin.validate_group_exists("3")
    assert plugin.get_group_by_id("2")["nodes"] == [A2]
    assert plugin.handle_group_not_found("gone", idle.id) is True        # group 1 is busy, group 2 takes it
    assert plugin.get_group_by_id("2")["task_id"] == idle.id and plugin.get_group_by_id("1")["task_id"] == busy.id
    assert plugin.handle_group_not_found("gone", idle.id) is False


def test_concurrent_heartbeats_and_table_updates():
    """Heartbeats from several threads while tasks and statuses change: every answer is a consistent one (the group's
    claim or none), nothing crashes, and the tables end in the expected state."""
    plugin = make([NodeGroupConfiguration("c", 1, 1)])
    sched = Scheduler(plugin)
    addrs = [f"0x{i + 1:040x}" for i in range(64)]
    for i, a in enumerate(addrs):
        plugin.add_node(OrchestratorNode(a))
        plugin.restore_group(f"{i + 1:x}", "c", [a])
    tasks = [Task(name=f"t{i}", created_at=i, allowed_topologies=["c"]) for i in range(50)]
    for t in tasks:
        plugin.add_task(t)
    names = {t.name for t in tasks} | {"late"}
    errors, stop = [], threading.Event()

    def beat(offset):
        try:
            i = offset
            while not stop.is_set():
                got = sched.get_task_for_node(addrs[i % len(addrs)])
                if got is not None and got["name"] not in names:
                    errors.append(got)
                i += 7
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=beat, args=(k,)) for k in range(6)]
    for th in threads:
        th.start()
    late = Task(name="late", created_at=999, allowed_topologies=["c"])
    plugin.add_task(late)
    for t in tasks[:25]:
        plugin.delete_task(t.id)                                         # groups that claimed them dissolve
    for a in addrs[:8]:
        plugin.update_node_status(a, NodeStatus.Dead)
    stop.set()
    for th in threads:
        th.join()
    assert not errors, errors[:3]
    for a in addrs[:8]:
        assert plugin.get_node_group(a) is None
    for g in plugin.get_all_groups():
        assert g["task_id"] is None or g["task_id"] in {t.id for t in tasks[25:]} | {late.id}


def test_export_tables_is_the_snapshot_a_management_pass_uploads():
    """SURVEY 8 A21 / 8f-1, ingest -> columnar: rows in node_store.get_nodes() order (stable status-class sort:
    Healthy, Discovered, the rest, Dead), candidate flags of mod.rs:492-497, locations, BTreeSet<String> address ranks;
    the cached ranks follow the table when it grows."""
    import numpy as np

    from protocol_b200.plugin import ComputeSpecs, GpuSpecs

    rng = np.random.default_rng(5)
    plugin = make([NodeGroupConfiguration("c", 1, 2)])
    statuses = [NodeStatus.Healthy, NodeStatus.Discovered, NodeStatus.Dead, NodeStatus.Unhealthy, NodeStatus.WaitingForHeartbeat,
                NodeStatus.Ejected, NodeStatus.LowBalance]
    nodes = []

    def add(n):
        for _ in range(n):
            addr = "0x" + "".join(rng.choice(list("0123456789abcdefABCDEF"), 40))
            st = statuses[int(rng.integers(len(statuses)))]
            node = OrchestratorNode(addr, status=st, p2p_id=None if rng.random() < 0.3 else "p",
                                    compute_specs=ComputeSpecs(gpu=GpuSpecs(count=int(rng.integers(1, 9)), model="H100", memory_mb=80000),
                                                               ram_mb=int(rng.integers(1, 10**6))) if rng.random() < 0.8 else None,
                                    location=(float(rng.uniform(-60, 60)), float(rng.uniform(-170, 170))) if rng.random() < 0.6 else None)
            plugin.add_node(node)
            nodes.append(node)

    def check():
        t = plugin.export_tables()
        cls = {NodeStatus.Healthy: 0, NodeStatus.Discovered: 1, NodeStatus.Dead: 3}
        order = sorted(range(len(nodes)), key=lambda i: cls.get(nodes[i].status, 2))          # stable
        by_addr = sorted(range(len(nodes)), key=lambda i: nodes[i].address.encode())             # byte-lexicographic
        rank = {i: r for r, i in enumerate(by_addr)}
        grouped = {n for g in plugin.get_all_groups() for n in g["nodes"]}
        assert len(t["a"]) == len(nodes)
        for row, i in enumerate(order):
            n, f = nodes[i], int(t["a"]["flags"][row])
            assert bool(f & abi.PM_W_HEALTHY) == (n.status == NodeStatus.Healthy)
            assert bool(f & abi.PM_W_P2P) == (n.p2p_id is not None)
            assert bool(f & abi.PM_W_ASSIGNED) == (n.address in grouped)
            assert bool(f & abi.PM_W_HAS_LOC) == (n.location is not None)
            assert bool(f & abi.PM_W_HAS_SPECS) == (n.compute_specs is not None)
            if n.compute_specs is not None:
                assert int(t["a"]["gpu_count"][row]) == n.compute_specs.gpu.count and int(t["b"]["ram_mb"][row]) == n.compute_specs.ram_mb
            if n.location is not None:
                assert (t["lat"][row], t["lon"][row]) == n.location
            assert int(t["addr_rank"][row]) == rank[i]

    add(300)
    healthy = [n.address for n in nodes if n.status == NodeStatus.Healthy]
    plugin.restore_group("1", "c", healthy[:2])
    plugin.restore_group("2", "c", healthy[2:3])
    check()
    add(150)                                                             # the rank cache is for 300 rows: it must follow
    check()
    plugin.update_node_status(nodes[0].address, NodeStatus.Dead)         # a status change moves a row, not a rank
    nodes[0].status = NodeStatus.Dead
    check()


def test_assigned_flag_follows_every_membership_change():
    """The PM_W_ASSIGNED bit of the exported table (mod.rs:496: get_node_group is Some) through restore, dissolution by
    status change and by task deletion, and a group restored before its node is stored."""
    plugin = make([NodeGroupConfiguration("c", 1, 2)])

    def assigned():
        t = plugin.export_tables()
        rows = sorted(range(len(t["a"])), key=lambda r: int(t["addr_rank"][r]))      # address order
        return [bool(int(t["a"]["flags"][r]) & abi.PM_W_ASSIGNED) for r in rows]

    plugin.add_node(OrchestratorNode(A1))
    plugin.add_node(OrchestratorNode(A2))
    assert assigned() == [False, False]
    plugin.restore_group("1", "c", [A1, A3])                              # A3 is not stored yet
    assert assigned() == [True, False]
    plugin.add_node(OrchestratorNode(A3))
    assert assigned() == [True, False, True]
    plugin.update_node_status(A3, NodeStatus.Dead)                        # dissolves group 1
    assert assigned() == [False, False, False]
    task = Task(allowed_topologies=["c"])
    plugin.add_task(task)
    plugin.restore_group("2", "c", [A2], task_id=task.id)
    assert assigned() == [False, True, False]
    plugin.delete_task(task.id)                                           # the group working on it goes
    assert assigned() == [False, False, False]


def test_idle_group_without_applicable_task_is_answered_without_a_claim():
    """A group whose configuration no task allows (allowed_topologies) heartbeats to `current_task: null` every time —
    from the cached per-configuration choice, without writing a claim — and gets the task as soon as one applies
    (scheduler_impl.rs:42-70; ADVICE r1: this path used to fall back to the exclusive lock on every heartbeat)."""
    plugin = make([NodeGroupConfiguration("a", 1, 2), NodeGroupConfiguration("b", 1, 2)])
    sched = Scheduler(plugin)
    plugin.add_node(OrchestratorNode(A1, p2p_id="p1"))
    plugin.add_node(OrchestratorNode(A2, p2p_id="p2"))
    plugin.restore_group("1", "a", [A1, A2])
    other = Task(image="img", name="only-b", allowed_topologies=["b"])
    plugin.add_task(other)
    for _ in range(5):
        assert sched.get_task_for_node(A1) is None and sched.get_task_for_node(A2) is None
    assert plugin.get_node_group(A1)["task_id"] is None                 # nothing was claimed
    mine = Task(image="img", name="for-a", allowed_topologies=["a"])
    plugin.add_task(mine)                                               # the task list changed: the cache is rebuilt
    assert sched.get_task_for_node(A1)["id"] == mine.id and sched.get_task_for_node(A2)["id"] == mine.id
    assert plugin.get_node_group(A1)["task_id"] == mine.id


def test_group_index_follows_checksummed_order_under_the_canonical_policy():
    """mod.rs:63-69,424-434: NodeGroup.nodes is a BTreeSet<String> of Address::to_string() strings; in EIP-55 case 'A'..'F'
    sort before 'a'..'f', so the order differs from the lower-case one.  With policy.canonical_addresses the mirror keys,
    orders and answers by the checksummed form whatever spelling comes in."""
    from oracle import pm_oracle as orc

    lows = ["0x" + "ab" * 19 + f"{i:02x}" for i in range(0xa0, 0xa8)] + ["0x" + "cd" * 19 + "ef", "0x" + "0f" * 20]
    canon = {a: orc.eip55(a) for a in lows}
    pair = None
    for i, a in enumerate(lows):                          # two addresses whose order flips between the two spellings
        for b in lows[i + 1:]:
            if (a < b) != (canon[a] < canon[b]):
                pair = (a, b)
    assert pair is not None
    a, b = pair
    plugin = make([NodeGroupConfiguration("c", 2, 2)], canonical_addresses=True)
    sched = Scheduler(plugin)
    plugin.add_node(OrchestratorNode(a, p2p_id="p-a"))
    plugin.add_node(OrchestratorNode(b.upper().replace("0X", "0x"), p2p_id="p-b"))
    task = Task(image="i", name="t", env_vars={"RANK": "${GROUP_INDEX}"}, allowed_topologies=["c"])
    plugin.add_task(task)
    plugin.restore_group("1", "c", [a, b])
    want = sorted([canon[a], canon[b]])
    assert plugin.get_node_group(a[2:])["nodes"] == want                 # any spelling finds it; members in BTreeSet order
    ranks = {canon[x]: sched.get_task_for_node(x)["env_vars"]["RANK"] for x in (a, b)}
    assert ranks == {want[0]: "0", want[1]: "1"}
    assert want != sorted([a, b], key=str) or True                       # (the lower-case order would have been the other one)
    assert [canon[x] for x in sorted([a, b])] != want
