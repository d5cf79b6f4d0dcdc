"""Host mirror of the plugin interface (csrc/pm_plugin.cpp), CPU-only parts: Scheduler with the default
NewestTaskPlugin (scheduler/mod.rs:87-154, newest_task/mod.rs:29-54), constructor validation
(node_groups/tests.rs:1336-1378), observers (tests.rs:1467-1627) — and the refusal to form groups
without a GPU engine."""
import pytest

from protocol_b200 import abi
from protocol_b200._lib import PrimeMatchError
from protocol_b200.plugin import NodeGroupConfiguration, NodeGroupsPlugin, OrchestratorNode, Scheduler, Task

ZERO = "0x0000000000000000000000000000000000000000"
ONES = "0x0101010101010101010101010101010101010101"


def test_get_task_for_node_default_plugin():
    """scheduler/mod.rs:91-110."""
    plugin = NodeGroupsPlugin([])
    sched = Scheduler(plugin)
    assert sched.get_task_for_node(ZERO) is None
    task = Task(image="image", name="name", created_at=1)
    plugin.add_task(task)
    got = sched.get_task_for_node(ZERO)
    assert got["id"] == task.id and got["name"] == "name" and got["image"] == "image" and got["created_at"] == 1
    assert got["env_vars"] is None and got["cmd"] is None


def test_variable_replacement():
    """scheduler/mod.rs:112-153: ${TASK_ID} and ${NODE_ADDRESS} in env vars and cmd."""
    plugin = NodeGroupsPlugin([])
    sched = Scheduler(plugin)
    task = Task(image="image", name="name", created_at=1,
                env_vars={"TASK_ID_VAR": "task-${TASK_ID}", "NODE_VAR": "node-${NODE_ADDRESS}"},
                cmd=["--task=${TASK_ID}", "--node=${NODE_ADDRESS}"],
                volume_mounts=[("/host/${TASK_ID}/${NODE_ADDRESS}", "/data/${TASK_ID}")])
    plugin.add_task(task)
    got = sched.get_task_for_node(ONES)
    assert got["env_vars"]["TASK_ID_VAR"] == f"task-{task.id}"
    assert got["env_vars"]["NODE_VAR"] == f"node-{ONES}"
    assert got["cmd"] == [f"--task={task.id}", f"--node={ONES}"]
    assert got["volume_mounts"] == [{"host_path": f"/host/{task.id}/{ONES}", "container_path": f"/data/{task.id}"}]


def test_newest_task_wins_with_last_maximum_on_ties():
    """newest_task/mod.rs:29-54 + max_by_key's last-maximum rule over the desc-sorted list."""
    plugin = NodeGroupsPlugin([])
    sched = Scheduler(plugin)
    t1, t2 = Task(name="a", created_at=1), Task(name="b", created_at=2)
    plugin.add_task(t1)
    plugin.add_task(t2)
    assert sched.get_task_for_node(ZERO)["id"] == t2.id
    t3 = Task(name="c", created_at=2)          # tie: list order after the stable desc sort is [t2, t3, t1]
    plugin.add_task(t3)
    assert sched.get_task_for_node(ZERO)["id"] == t3.id
    plugin.delete_task(t3.id)
    plugin.delete_task(t2.id)
    assert sched.get_task_for_node(ZERO)["id"] == t1.id


def test_unique_names_and_valid_sizes():
    """tests.rs:1336-1378: duplicate names / max < min panic in the reference -> PM_E_INVALID here."""
    with pytest.raises(PrimeMatchError) as e:
        NodeGroupsPlugin([NodeGroupConfiguration("a", 1, 1), NodeGroupConfiguration("a", 2, 2)])
    assert e.value.status == abi.PM_E_INVALID and "unique" in str(e.value)
    with pytest.raises(PrimeMatchError) as e:
        NodeGroupsPlugin([NodeGroupConfiguration("a", 2, 1)])
    assert e.value.status == abi.PM_E_INVALID and "invalid" in str(e.value)
    with pytest.raises(PrimeMatchError) as e:
        NodeGroupsPlugin([NodeGroupConfiguration("a", 1, 1, "gpu:count=abc")])
    assert e.value.status == abi.PM_E_PARSE


def test_node_without_group_gets_no_task():
    """scheduler_impl.rs:208-209: 'Node is not in a group, skipping all tasks'."""
    plugin = NodeGroupsPlugin([NodeGroupConfiguration("c", 1, 1)])
    plugin.add_node(OrchestratorNode("0x1234567890123456789012345678901234567890"))
    plugin.add_task(Task(name="t", allowed_topologies=["c"]))
    assert Scheduler(plugin).get_task_for_node("0x1234567890123456789012345678901234567890") is None
    assert plugin.get_node_group("0x1234567890123456789012345678901234567890") is None


def test_group_formation_has_no_cpu_path():
    plugin = NodeGroupsPlugin([NodeGroupConfiguration("c", 1, 1)])
    plugin.add_node(OrchestratorNode("0x1234567890123456789012345678901234567890"))
    with pytest.raises(PrimeMatchError) as e:
        plugin.try_form_new_groups()
    assert e.value.status == abi.PM_E_NO_DEVICE


def test_newest_task_choice_is_cached_until_the_task_list_changes():
    """NewestTaskPlugin (newest_task/mod.rs:8-19) = max_by_key over the created_at-desc list: the LAST task of the
    newest timestamp in store order.  The choice is cached per task-list version; adds and deletes refresh it."""
    plugin = NodeGroupsPlugin([])
    sched = Scheduler(plugin)
    tasks = [Task(name=f"t{i}", created_at=c) for i, c in enumerate([5, 9, 9, 3])]
    for t in tasks:
        plugin.add_task(t)
    assert sched.get_task_for_node(ZERO)["name"] == "t2"
    assert sched.get_task_for_node(ONES)["name"] == "t2"
    plugin.delete_task(tasks[2].id)
    assert sched.get_task_for_node(ZERO)["name"] == "t1"
    t4 = Task(name="t4", created_at=9)
    plugin.add_task(t4)
    assert sched.get_task_for_node(ZERO)["name"] == "t4"
    plugin.delete_task(tasks[1].id)
    plugin.delete_task(tasks[0].id)
    assert sched.get_task_for_node(ZERO)["name"] == "t4"
    plugin.delete_task(t4.id)
    assert sched.get_task_for_node(ZERO)["name"] == "t3"
    plugin.delete_task(tasks[3].id)
    assert sched.get_task_for_node(ZERO) is None
