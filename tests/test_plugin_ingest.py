"""Discovery ingest: the reference's DiscoveryMonitor tests
(crates/orchestrator/src/discovery/monitor.rs:449-800) against the host mirror, plus the
status-transition table of sync_single_node_with_discovery (:236-420).  CPU only."""
from protocol_b200.plugin import ComputeSpecs, DiscoveryNode, NodeGroupsPlugin, NodeStatus, OrchestratorNode

A1 = "0x1234567890123456789012345678901234567890"
A2 = "0x2234567890123456789012345678901234567890"
A3 = "0x3234567890123456789012345678901234567890"
SPECS = ComputeSpecs(ram_mb=1024, storage_gb=10)
NOW = 1_700_000_000_000


def test_sync_single_node_with_discovery():
    """monitor.rs:449-560: store has the node Ejected; discovery says validated + whitelisted (inactive)
    -> the node is marked Dead so it can recover."""
    plugin = NodeGroupsPlugin([])
    plugin.add_node(OrchestratorNode(A1, status=NodeStatus.Ejected, compute_specs=SPECS, ip_address="127.0.0.1", port=8080))
    assert plugin.get_node(A1)["status"] == "Ejected"
    plugin.sync_discovery([DiscoveryNode(A1, "127.0.0.1", 8080, SPECS, is_validated=True, is_provider_whitelisted=True,
                                         is_active=False)], NOW)
    assert plugin.get_node(A1)["status"] == "Dead"


def test_first_seen_timestamp_set_on_new_node():
    """monitor.rs:562-674."""
    plugin = NodeGroupsPlugin([])
    dn = DiscoveryNode(A1, "192.168.1.100", 8080, SPECS, is_validated=True, is_provider_whitelisted=True, is_active=True)
    assert plugin.sync_discovery([dn], NOW) == 1
    node = plugin.get_node(A1)
    assert node["first_seen_ms"] == NOW and node["status"] == "Discovered" and node["ip_address"] == "192.168.1.100"
    dn.ip_address = "192.168.1.101"
    assert plugin.sync_discovery([dn], NOW + 60_000) == 0
    node = plugin.get_node(A1)
    assert node["first_seen_ms"] == NOW and node["ip_address"] == "192.168.1.101" and node["status"] == "Discovered"


def test_sync_node_with_same_endpoint():
    """monitor.rs:676-800: a new node on the endpoint of a healthy node is not added; another port is."""
    plugin = NodeGroupsPlugin([])
    plugin.add_node(OrchestratorNode(A1, status=NodeStatus.Healthy, compute_specs=SPECS, ip_address="127.0.0.1", port=8080))
    mk = lambda a, port: DiscoveryNode(a, "127.0.0.1", port, SPECS, is_validated=True, is_provider_whitelisted=True, is_active=True)
    assert plugin.sync_discovery([mk(A2, 8080)], NOW) == 0
    assert plugin.get_node(A2) is None
    assert plugin.sync_discovery([mk(A3, 8081)], NOW) == 1
    assert plugin.get_node(A3) is not None


def test_transitions_table():
    plugin = NodeGroupsPlugin([])
    mk = lambda a, **kw: DiscoveryNode(a, "10.0.0.1", 1000 + int(a[2]), SPECS, is_validated=True, **kw)
    # validated but not whitelisted -> Ejected (:270-281)
    plugin.add_node(OrchestratorNode(A1, status=NodeStatus.Healthy, ip_address="10.0.0.1", port=1001, last_status_change_ms=NOW))
    plugin.sync_discovery([mk(A1, is_provider_whitelisted=False, is_active=True)], NOW + 1000)
    assert plugin.get_node(A1)["status"] == "Ejected"
    # healthy but no longer active on chain: grace period of 5 minutes (:300-338)
    plugin.add_node(OrchestratorNode(A2, status=NodeStatus.Healthy, ip_address="10.0.0.1", port=1002, last_status_change_ms=NOW))
    plugin.sync_discovery([mk(A2, is_provider_whitelisted=True, is_active=False)], NOW + 60_000)
    assert plugin.get_node(A2)["status"] == "Healthy"
    plugin.sync_discovery([mk(A2, is_provider_whitelisted=True, is_active=False)], NOW + 6 * 60_000)
    assert plugin.get_node(A2)["status"] == "Dead"
    # dead node updated on discovery after its death -> Discovered again, with the new specs (:364-389)
    last_change = plugin.get_node(A2)["last_status_change_ms"]
    plugin.sync_discovery([DiscoveryNode(A2, "10.0.0.1", 1002, ComputeSpecs(ram_mb=2048), is_validated=True,
                                         is_provider_whitelisted=True, is_active=True, last_updated_ms=last_change + 1)],
                          last_change + 10)
    n2 = plugin.get_node(A2)
    assert n2["status"] == "Discovered" and n2["ram_mb"] == 2048
    # zero balance -> LowBalance (:391-402); not validated -> ignored; duplicates by id -> first wins (:203-210)
    plugin.add_node(OrchestratorNode(A3, status=NodeStatus.Healthy, ip_address="10.0.0.1", port=1003, last_status_change_ms=NOW))
    plugin.sync_discovery([mk(A3, is_provider_whitelisted=True, is_active=True, latest_balance=0)], NOW + 1000)
    assert plugin.get_node(A3)["status"] == "LowBalance"
    a4 = "0x4234567890123456789012345678901234567890"
    assert plugin.sync_discovery([DiscoveryNode(a4, "10.0.0.9", 9, SPECS, is_validated=False)], NOW) == 0
    first = DiscoveryNode(a4, "10.0.0.9", 9, SPECS, is_validated=True, is_provider_whitelisted=True, is_active=True)
    dup = DiscoveryNode(a4, "10.0.0.10", 10, SPECS, is_validated=True, is_provider_whitelisted=True, is_active=True)
    assert plugin.sync_discovery([first, dup], NOW) == 1
    assert plugin.get_node(a4)["ip_address"] == "10.0.0.9"


def test_endpoint_index_scales_linearly():
    """The reference re-reads every node for every synced node (Theta(N^2), monitor.rs:218-234);
    20k nodes over 2k shared endpoints sync in well under a second through the endpoint index."""
    import time

    plugin = NodeGroupsPlugin([])
    n = 20_000
    nodes = [DiscoveryNode(f"0x{i:040x}", f"10.1.{(i % 2000) // 250}.{i % 250}", 8000 + (i % 3), SPECS, is_validated=True,
                           is_provider_whitelisted=True, is_active=True) for i in range(n)]
    t0 = time.perf_counter()
    assert plugin.sync_discovery(nodes, NOW, max_healthy_nodes_with_same_endpoint=1) == n   # nobody healthy yet
    plugin.sync_discovery(nodes, NOW + 1000)
    assert time.perf_counter() - t0 < 5.0


def test_sync_from_discovery_wire_format():
    """The JSON the discovery service returns (ApiResponse<Vec<DiscoveryNode>>, node.rs:10-23, 552-570)."""
    import json

    import pytest

    from protocol_b200 import abi
    from protocol_b200._lib import PrimeMatchError

    body = json.dumps({"success": True, "data": [
        {"id": A1, "provider_address": A1, "ip_address": "192.168.1.100", "port": 8080, "compute_pool_id": 1,
         "compute_specs": {"gpu": {"count": 8, "model": "NVIDIA H100", "memory_mb": 80000, "indices": [0, 1]},
                           "cpu": {"cores": 64, "model": "x"}, "ram_mb": 1024, "storage_gb": 10, "storage_path": "/var/lib"},
         "is_validated": True, "is_active": True, "is_provider_whitelisted": True, "is_blacklisted": False,
         "last_updated": "2024-05-01T12:00:00.250Z", "created_at": None,
         "location": {"latitude": 45.5186, "longitude": -73.5545, "city": "Montr\\u00e9al", "region": None, "country": "CA"},
         "latest_balance": "0x0"},
        {"id": A2, "provider_address": A2, "ip_address": "192.168.1.101", "port": 8081, "compute_pool_id": 1,
         "compute_specs": None, "is_validated": False, "is_active": True},
        {"id": A3, "provider_address": A3, "ip_address": "192.168.1.102", "port": 8082, "compute_pool_id": 1,
         "compute_specs": {"gpu": None, "cpu": None, "ram_mb": 2048, "storage_gb": None, "storage_path": "/x"},
         "is_validated": True, "is_active": True, "is_provider_whitelisted": True, "latest_balance": "1000000000000000000"},
    ]})
    plugin = NodeGroupsPlugin([])
    assert plugin.sync_discovery_json(body, NOW) == 2                  # the unvalidated node is skipped
    n1 = plugin.get_node(A1)
    assert n1["ip_address"] == "192.168.1.100" and n1["port"] == 8080 and n1["status"] == "Discovered"
    assert n1["has_location"] and n1["has_compute_specs"] and n1["ram_mb"] == 1024
    assert plugin.get_node(A2) is None
    assert plugin.get_node(A3)["ram_mb"] == 2048 and not plugin.get_node(A3)["has_location"]
    # second sync: the zero balance now applies to the existing node (monitor.rs:391-402)
    assert plugin.sync_discovery_json(body, NOW + 1000) == 0
    assert plugin.get_node(A1)["status"] == "LowBalance" and plugin.get_node(A3)["status"] == "Discovered"
    assert plugin.sync_discovery_json("[]", NOW) == 0
    for bad in ("{", '{"data": 3}', '[{"id": 1}]', '[] trailing'):
        with pytest.raises(PrimeMatchError) as e:
            plugin.sync_discovery_json(bad, NOW)
        assert e.value.status == abi.PM_E_PARSE


def _wire(node_id, ip="10.0.0.1", port=8080, **extra):
    d = {"id": node_id, "provider_address": node_id, "ip_address": ip, "port": port, "compute_pool_id": 1,
         "compute_specs": {"gpu": {"count": 8, "model": "NVIDIA H100", "memory_mb": 80000, "indices": [0, 1, [2, {"x": "]}\\\""}]]},
                           "cpu": {"cores": 64, "model": "x\\\"y"}, "ram_mb": 1024, "storage_gb": 10, "storage_path": "/var/{lib}"},
         "is_validated": True, "is_active": True, "is_provider_whitelisted": True, "is_blacklisted": False,
         "location": {"latitude": 45.5, "longitude": -73.5, "city": "Montr\u00e9al"}}
    d.update(extra)
    return d


def test_streaming_reader_skips_what_it_does_not_need():
    """The body is read in one streaming pass: members the monitor does not use (nested, with quotes, braces and
    escapes inside strings) are skipped in place, 'data' need not come first, and a key written with an escape
    still matches."""
    import json

    body = json.dumps({"meta": {"pages": [1, 2, {"next": None}], "note": "a \"quoted\" } ] string"}, "success": True,
                       "data": [_wire(A1), _wire(A3, ip="10.0.0.3")], "trailer": [[[]]]})
    body = body.replace('"id": "' + A1, '"i\\u0064": "' + A1)        # "i\u0064" == "id"
    plugin = NodeGroupsPlugin([])
    assert plugin.sync_discovery_json("  \n" + body + " \t\n", NOW) == 2
    n1 = plugin.get_node(A1)
    assert n1["ip_address"] == "10.0.0.1" and n1["has_location"] and n1["has_compute_specs"] and n1["ram_mb"] == 1024
    assert plugin.get_node(A3)["ip_address"] == "10.0.0.3"


def test_streaming_reader_field_rules():
    """First occurrence of a duplicated key counts (as `Value::get` did); wrongly typed members are absent, not errors:
    a string port is 0, a string latitude means no location, compute_specs: null means no specs."""
    import json

    a = json.dumps(_wire(A1, port="8080", location={"latitude": "45.5", "longitude": -73.5}, compute_specs=None))
    a = a[:-1] + ', "ip_address": "9.9.9.9", "is_validated": false}'      # later duplicates are ignored
    b = json.dumps(_wire(A2, latest_balance="0x0000"))
    plugin = NodeGroupsPlugin([])
    assert plugin.sync_discovery_json('{"data": [' + a + "," + b + '], "data": [{"broken": 1}]}', NOW) == 2
    n1 = plugin.get_node(A1)
    assert n1["ip_address"] == "10.0.0.1" and n1["port"] == 0 and not n1["has_location"] and not n1["has_compute_specs"]
    assert plugin.get_node(A2)["status"] == "Discovered"
    assert plugin.sync_discovery_json("[" + b + "]", NOW + 1000) == 0      # zero balance applies to the existing node
    assert plugin.get_node(A2)["status"] == "LowBalance"


def test_malformed_body_applies_nothing():
    """`response.json()` failing in the reference means the fetch does nothing: a body that breaks after ten thousand
    good nodes (inside a member that is only skipped) leaves the table untouched."""
    import json

    import pytest

    from protocol_b200 import abi
    from protocol_b200._lib import PrimeMatchError

    good = ",".join(json.dumps(_wire(f"0x{i:040x}", ip=f"10.{i >> 8}.{i & 255}.1")) for i in range(1, 10001))
    plugin = NodeGroupsPlugin([])
    for bad_tail in (', {"id": "x", "ip_address": "y", "junk": [1, 2}]', ', 7]', ', {"id": "x"}]', "]]", ", {\"id\": \"x\", \"ip_address\": \"y\", \"s\": \"\\q\"}]"):
        with pytest.raises(PrimeMatchError) as e:
            plugin.sync_discovery_json("[" + good + bad_tail, NOW)
        assert e.value.status == abi.PM_E_PARSE
        assert plugin.get_node(f"0x{1:040x}") is None
    assert plugin.sync_discovery_json("[" + good + "]", NOW) == 10000


def test_duplicate_ids_across_chunks_and_unstored_nodes():
    """'Remove duplicates based on node ID' (monitor.rs:203-210) holds across the chunks the body is applied in
    (8192 nodes each), and for an id that was not stored the first time (endpoint already taken by a healthy node)."""
    import json

    plugin = NodeGroupsPlugin([])
    plugin.add_node(OrchestratorNode(A1, status=NodeStatus.Healthy, compute_specs=SPECS, ip_address="10.9.9.9", port=1))
    ids = [f"0x{i:040x}" for i in range(100, 9100)]
    nodes = [_wire(a, ip=f"10.{(i >> 8) & 255}.{i & 255}.7") for i, a in enumerate(ids)]
    blocked = "0x" + "b" * 40
    nodes.insert(5, _wire(blocked, ip="10.9.9.9", port=1))                 # same endpoint as the healthy A1: not stored
    nodes.append(_wire(ids[0], ip="1.1.1.1"))                              # duplicate of the first id, 9000 nodes later
    nodes.append(_wire(blocked, ip="10.9.9.10", port=2))                   # duplicate of the unstored id: still skipped
    assert plugin.sync_discovery_json(json.dumps(nodes), NOW) == len(ids)
    assert plugin.get_node(ids[0])["ip_address"] == "10.0.0.7"
    assert plugin.get_node(blocked) is None
    # the next fetch is a new one: the id is free again
    assert plugin.sync_discovery_json(json.dumps([_wire(blocked, ip="10.9.9.10", port=2)]), NOW + 1000) == 1


def test_json_and_struct_ingest_agree_on_random_fetches():
    """Differential property: the same fetch given as C structs (pm_plugin_sync_discovery) and as the discovery
    service's JSON body (pm_plugin_sync_discovery_json, with members the monitor does not read mixed in) leaves the
    same node table — presence bits, interned model, location, status — over several rounds of random changes."""
    import json

    from hypothesis import given, settings
    from hypothesis import strategies as st

    from protocol_b200.plugin import GpuSpecs

    models = [None, "NVIDIA H100 80GB HBM3", "NVIDIA GeForce RTX 4090", "A100", "weird \"quoted\" \\ model"]
    gpu = st.one_of(st.none(), st.builds(GpuSpecs, count=st.one_of(st.none(), st.integers(0, 16)), model=st.sampled_from(models),
                                         memory_mb=st.one_of(st.none(), st.integers(0, 200_000))))
    specs = st.one_of(st.none(), st.builds(ComputeSpecs, gpu=gpu, cpu_cores=st.one_of(st.none(), st.integers(0, 256)),
                                           cpu_present=st.one_of(st.none(), st.booleans()),
                                           ram_mb=st.one_of(st.none(), st.integers(0, 2**32 - 1)),
                                           storage_gb=st.one_of(st.none(), st.integers(0, 100_000))))
    node = st.builds(DiscoveryNode, id=st.sampled_from([f"0x{i:040x}" for i in range(1, 9)]),
                     ip_address=st.sampled_from(["10.0.0.1", "10.0.0.2", "192.168.100.200"]), port=st.integers(0, 65535),
                     compute_specs=specs, is_validated=st.booleans(), is_active=st.booleans(), is_provider_whitelisted=st.booleans(),
                     is_blacklisted=st.booleans(), last_updated_ms=st.one_of(st.none(), st.integers(NOW - 10**6, NOW + 10**6)),
                     location=st.one_of(st.none(), st.tuples(st.floats(-90, 90), st.floats(-180, 180))),
                     latest_balance=st.one_of(st.none(), st.sampled_from([0, 1, 10**18])))

    def wire(n):
        d = {"id": n.id, "provider_address": n.id, "ip_address": n.ip_address, "port": n.port, "compute_pool_id": 7,
             "extra": {"nested": [1, {"deep": "]}"}], "s": "a\\\"b"}, "is_validated": n.is_validated, "is_active": n.is_active,
             "is_provider_whitelisted": n.is_provider_whitelisted, "is_blacklisted": n.is_blacklisted}
        s = n.compute_specs
        if s is None:
            d["compute_specs"] = None
        else:
            cpu_present = s.cpu_present if s.cpu_present is not None else s.cpu_cores is not None
            d["compute_specs"] = {
                "gpu": None if s.gpu is None else {"count": s.gpu.count, "model": s.gpu.model, "memory_mb": s.gpu.memory_mb, "indices": [0, 1]},
                "cpu": {"cores": s.cpu_cores, "model": "cpu"} if cpu_present else None,
                "ram_mb": s.ram_mb, "storage_gb": s.storage_gb, "storage_path": "/x"}
        if n.last_updated_ms is not None:
            ms = n.last_updated_ms
            import datetime
            d["last_updated"] = datetime.datetime.fromtimestamp(ms / 1000, datetime.timezone.utc).strftime("%Y-%m-%dT%H:%M:%S.") + f"{ms % 1000:03d}Z"
        if n.location is not None:
            d["location"] = {"latitude": n.location[0], "longitude": n.location[1], "city": None}
        if n.latest_balance is not None:
            d["latest_balance"] = hex(n.latest_balance) if n.latest_balance % 2 == 0 else str(n.latest_balance)
        return d

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.lists(node, max_size=12), min_size=1, max_size=4))
    def check(fetches):
        a, b = NodeGroupsPlugin([]), NodeGroupsPlugin([])
        for r, fetch in enumerate(fetches):
            now = NOW + r * 400_000
            na = a.sync_discovery(fetch, now)
            nb = b.sync_discovery_json(json.dumps({"success": True, "data": [wire(n) for n in fetch]}), now)
            assert na == nb
            for i in range(1, 9):
                assert a.get_node(f"0x{i:040x}") == b.get_node(f"0x{i:040x}")

    check()


def test_entry_without_id_or_ip_is_skipped_and_the_rest_of_the_fetch_applies():
    """The monitor logs a node it cannot sync and goes on with the others (monitor.rs:425-429); the struct entry point
    used to return PM_E_INVALID at the first such entry and drop the rest of the fetch."""
    import ctypes as C

    from protocol_b200 import abi

    plugin = NodeGroupsPlugin([])
    arr = (abi.PmDiscoveryNode * 3)()
    keep = []
    for i, (addr, ip) in enumerate([("0x" + "1" * 40, "10.0.0.1"), ("0x" + "2" * 40, None), ("0x" + "3" * 40, "10.0.0.3")]):
        d = arr[i]
        keep.append((addr.encode(), ip.encode() if ip else None))
        d.node.address = keep[-1][0]
        d.ip_address = keep[-1][1]
        d.port = 8000 + i
        d.is_validated, d.is_active, d.is_provider_whitelisted = 1, 1, 1
        d.last_updated_ms = -1
    n_new = C.c_uint32()
    rc = plugin._lib.pm_plugin_sync_discovery(plugin._h, arr, 3, NOW, 1, C.byref(n_new))
    assert rc == abi.PM_OK and n_new.value == 2
    assert plugin.get_node("0x" + "1" * 40) is not None and plugin.get_node("0x" + "3" * 40) is not None
    assert plugin.get_node("0x" + "2" * 40) is None


def test_canonical_addresses_policy_parses_and_checksums_every_id():
    """the reference keys everything by Address::to_string() of the PARSED id (monitor.rs:240; BTreeSet member
    order mod.rs:63-69).  With policy.canonical_addresses the mirror does the same: a lower-case id in a discovery body, an
    upper-case one in a heartbeat and the checksummed one in a lookup are one node; an id that is not an address is skipped
    like the reference's parse error, the rest of the fetch applies."""
    import json

    import pytest

    from oracle import pm_oracle as orc
    from protocol_b200 import abi
    from protocol_b200._lib import PrimeMatchError

    low = ["0x" + f"{i:02x}" * 20 for i in (0xab, 0xcd, 0x1f)]
    canon = [orc.eip55(a) for a in low]
    assert all(c != a for c, a in zip(canon, low))                       # the checksum case really differs
    body = json.dumps([_wire(low[0], ip="10.0.0.1"), _wire("not-an-address", ip="10.0.0.9"), _wire(low[1].upper().replace("0X", "0x"), ip="10.0.0.2"),
                       _wire(canon[2], ip="10.0.0.3")])
    plugin = NodeGroupsPlugin([], canonical_addresses=True)
    assert plugin.sync_discovery_json(body, NOW) == 3
    for a, c in zip(low, canon):
        for spelled in (a, c, a[2:], "0x" + a[2:].upper()):
            node = plugin.get_node(spelled)
            assert node is not None and node["address"] == c
    assert plugin.sync_discovery_json(body, NOW + 1000) == 0             # the same three nodes, however they are spelled
    with pytest.raises(PrimeMatchError) as e:
        plugin.get_node("not-an-address")
    assert e.value.status == abi.PM_E_INVALID
    with pytest.raises(PrimeMatchError):
        plugin.add_node(OrchestratorNode("0x1234", status=NodeStatus.Healthy))
    # opaque strings without the policy (the default): three spellings are three nodes
    plain = NodeGroupsPlugin([])
    assert plain.sync_discovery_json(json.dumps([_wire(low[0]), _wire(canon[0], ip="10.0.0.2"), _wire("not-an-address", ip="10.0.0.3")]), NOW) == 3
    assert plain.get_node(low[0])["address"] == low[0] and plain.get_node(canon[0])["address"] == canon[0]


def test_canonical_policy_is_the_opaque_mirror_fed_with_checksummed_ids():
    """Differential property: a mirror with policy.canonical_addresses fed ids in random spellings (lower, upper, mixed, with
    and without 0x) ends in the same node table as an opaque mirror fed Address::to_string() of the same ids — what the
    Rust shim would pass — over several fetches through both ingest entry points."""
    import json
    import random

    from oracle import pm_oracle as orc

    rng = random.Random(7)
    lows = ["0x" + "".join(rng.choice("0123456789abcdef") for _ in range(40)) for _ in range(10)]

    def respell(a):
        h = "".join(c.upper() if rng.random() < 0.5 else c for c in a[2:])
        return rng.choice(["0x", "0X", ""]) + h

    opaque, canon = NodeGroupsPlugin([]), NodeGroupsPlugin([], canonical_addresses=True)
    for r in range(6):
        picked = rng.sample(lows, rng.randint(1, len(lows)))
        mk = lambda node_id, i: DiscoveryNode(node_id, f"10.0.{i}.1", 8000 + i, SPECS, is_validated=True, is_provider_whitelisted=rng.random() < 0.8,
                                              is_active=rng.random() < 0.7)
        state = rng.getstate()
        a_nodes = [mk(orc.eip55(a), lows.index(a)) for a in picked]
        rng.setstate(state)                                   # the same whitelisted / active draws for both mirrors
        b_nodes = [mk("placeholder", lows.index(a)) for a in picked]
        for n, a in zip(b_nodes, picked):
            n.id = respell(a)
        now = NOW + r * 400_000
        if r % 2:
            assert opaque.sync_discovery(a_nodes, now) == canon.sync_discovery(b_nodes, now)
        else:
            wire = lambda n: _wire(n.id, ip=n.ip_address, port=n.port, is_provider_whitelisted=n.is_provider_whitelisted, is_active=n.is_active)
            assert (opaque.sync_discovery_json(json.dumps([wire(n) for n in a_nodes]), now)
                    == canon.sync_discovery_json(json.dumps([wire(n) for n in b_nodes]), now))
        for a in lows:
            assert opaque.get_node(orc.eip55(a)) == canon.get_node(respell(a))
