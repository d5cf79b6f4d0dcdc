"""Shared test helpers: build reference-shaped nodes, engine tables and compare results."""
from __future__ import annotations

import numpy as np

from oracle import pm_oracle as orc
from protocol_b200 import abi
from protocol_b200.engine import Interner


def spec_to_orc_node(s: dict, address="0x0", status=orc.HEALTHY, p2p=True, location=None, assigned=False,
                     specs=True):
    return orc.make_node(address=address, status=status, p2p=p2p, specs=specs, location=location,
                         assigned=assigned, gpu_count=s.get("gpu_count"), gpu_model=s.get("gpu_model"),
                         gpu_mem=s.get("gpu_mem"), cpu_cores=s.get("cpu_cores"), ram=s.get("ram"),
                         storage=s.get("storage"), gpu_present=s.get("gpu_present"),
                         cpu_present=s.get("cpu_present"))


def spec_to_rows(s: dict | None, interner: Interner, healthy=True, p2p=True, assigned=False, has_loc=False):
    """ComputeSpecs kwargs (None == compute_specs: None) -> (WORKER_A row, WORKER_B row)."""
    a = np.zeros(1, dtype=abi.WORKER_A)
    b = np.zeros(1, dtype=abi.WORKER_B)
    f = 0
    if healthy:
        f |= abi.PM_W_HEALTHY
    if p2p:
        f |= abi.PM_W_P2P
    if assigned:
        f |= abi.PM_W_ASSIGNED
    if has_loc:
        f |= abi.PM_W_HAS_LOC
    if s is not None:
        f |= abi.PM_W_HAS_SPECS
        gpu_present = s.get("gpu_present")
        if gpu_present is None:
            gpu_present = any(s.get(k) is not None for k in ("gpu_count", "gpu_model", "gpu_mem"))
        if gpu_present:
            f |= abi.PM_W_HAS_GPU
            if s.get("gpu_count") is not None:
                f |= abi.PM_W_HAS_GPU_COUNT
                a["gpu_count"] = s["gpu_count"]
            if s.get("gpu_model") is not None:
                f |= abi.PM_W_HAS_GPU_MODEL
                a["model_id"] = interner.model(s["gpu_model"])
            if s.get("gpu_mem") is not None:
                f |= abi.PM_W_HAS_GPU_MEM
                a["gpu_mem_mb"] = s["gpu_mem"]
        cpu_present = s.get("cpu_present")
        if cpu_present is None:
            cpu_present = s.get("cpu_cores") is not None
        if cpu_present:
            f |= abi.PM_W_HAS_CPU
            if s.get("cpu_cores") is not None:
                f |= abi.PM_W_HAS_CPU_CORES
                b["cpu_cores"] = s["cpu_cores"]
        if s.get("ram") is not None:
            f |= abi.PM_W_HAS_RAM
            b["ram_mb"] = s["ram"]
        if s.get("storage") is not None:
            f |= abi.PM_W_HAS_STORAGE
            b["storage_gb"] = s["storage"]
    a["flags"] = f
    return a[0], b[0]


class TableBuilder:
    """Accumulates configs (requirement strings) and nodes into engine tables through the
    product's own parser/interner."""

    def __init__(self):
        self.it = Interner()
        self.asks = []
        self.opts = []
        self.wa = []
        self.wb = []
        self.lat = []
        self.lon = []

    def add_config(self, req: str | None, mn: int = 1, mx: int = 1):
        if req is None:
            row = np.zeros(1, dtype=abi.ASK)[0]
            o = np.zeros(0, dtype=abi.GPU_OPT)
        else:
            row, o = self.it.parse(req)
        row = row.copy()
        row["opt_off"] = sum(len(x) for x in self.opts)
        row["min_group_size"] = mn
        row["max_group_size"] = mx
        self.asks.append(row)
        self.opts.append(o)
        return len(self.asks) - 1

    def add_node(self, s: dict | None, healthy=True, p2p=True, assigned=False, location=None):
        a, b = spec_to_rows(s, self.it, healthy=healthy, p2p=p2p, assigned=assigned, has_loc=location is not None)
        self.wa.append(a)
        self.wb.append(b)
        self.lat.append(location[0] if location else 0.0)
        self.lon.append(location[1] if location else 0.0)
        return len(self.wa) - 1

    def tables(self):
        asks = np.array(self.asks, dtype=abi.ASK) if self.asks else np.zeros(0, dtype=abi.ASK)
        opts = np.concatenate(self.opts).astype(abi.GPU_OPT) if self.opts and sum(map(len, self.opts)) else np.zeros(0, dtype=abi.GPU_OPT)
        wa = np.array(self.wa, dtype=abi.WORKER_A) if self.wa else np.zeros(0, dtype=abi.WORKER_A)
        wb = np.array(self.wb, dtype=abi.WORKER_B) if self.wb else np.zeros(0, dtype=abi.WORKER_B)
        bits, npat, nmod, words = self.it.table()
        return dict(asks=asks, opts=opts, wa=wa, wb=wb, bits=bits, n_patterns=npat, n_models=nmod, words=words,
                    lat=np.array(self.lat, dtype=np.float64), lon=np.array(self.lon, dtype=np.float64))


def load_engine(eng, t: dict, addr_rank=None, locations=False):
    eng.set_asks(t["asks"], t["opts"])
    eng.set_model_table(t["bits"], t["n_patterns"], t["n_models"], t["words"])
    eng.set_workers(t["wa"], t["wb"])
    if addr_rank is not None:
        eng.set_addr_rank(addr_rank)
    if locations:
        eng.set_locations(t["lat"], t["lon"])


def groups_equal(res, og) -> bool:
    """Engine MatchResult vs oracle Groups: same groups, same creation order, same member order."""
    return (len(res.group_ask) == len(og.cfg) and np.array_equal(res.group_ask, og.cfg)
            and np.array_equal(res.group_off, og.off) and np.array_equal(res.group_members, og.members))
