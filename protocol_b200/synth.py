"""Deterministic synthetic swarms: splitmix64, fixed seeds.

Produces the engine's plain-data tables directly (vectorised) plus, for the
oracle's string-shaped path, the model / requirement strings they came from.
Distributions follow SURVEY.md 8(d); the model catalogue is seeded with the
strings the reference's own tests use (crates/shared/src/models/node.rs:744,
771-777,817-822,884-890,1033,1069,1217; node_groups/tests.rs:2914-2915).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import abi

SEED_WORKERS = 0xB2000001
SEED_ASKS = 0xB2000002
SEED_EXT = 0xB2000003

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """n outputs of splitmix64 started at `seed` (+ a per-column stream offset)."""
    with np.errstate(over="ignore"):
        base = np.uint64((seed + stream * 0xD1342543DE82EF95) & 0xFFFFFFFFFFFFFFFF)
        x = base + (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def _unit(seed: int, n: int, stream: int) -> np.ndarray:
    return (splitmix64(seed, n, stream) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def _choice(seed: int, n: int, stream: int, values, probs=None) -> np.ndarray:
    u = _unit(seed, n, stream)
    values = np.asarray(values)
    if probs is None:
        idx = np.minimum((u * len(values)).astype(np.int64), len(values) - 1)
    else:
        idx = np.searchsorted(np.cumsum(probs), u, side="right")
        idx = np.minimum(idx, len(values) - 1)
    return values[idx]


# (model string, memory_mb)
MODEL_CATALOGUE = [
    ("nvidia rtx a6000", 49140),
    ("NVIDIA H100", 80000),
    ("NVIDIA A100 80GB", 80000),
    ("NVIDIA A100", 40000),
    ("RTX 4090", 24000),
    ("RTX 3090", 24000),
    ("RTX 3060", 12000),
    ("nvidia_a100_80gb_pcie", 40000),
    ("NVIDIA_A100_80GB_PCIE", 40000),
    ("rtx4090", 24000),
    ("AMD Radeon RX 7900", 20000),
    ("NVIDIA H200", 141000),
    ("NVIDIA H100 80GB HBM3", 80000),
    ("NVIDIA GeForce RTX 4090", 24564),
    ("NVIDIA L40S", 46068),
    ("Tesla V100-SXM2-32GB", 32768),
]

PATTERN_CATALOGUE = [
    "a100,h100,h200", "H100", "A100", "RTX4090", "rtx 3090, rtx 4090", "nvidia",
    "a6000", "h100,h200", "rtx", "v100,a100", "l40s", "amd",
]

# 64 cities; the first two are the reference's proximity test (tests.rs:2920-2989)
_CITY_SEED = 0xC171E5
CITIES = np.zeros((64, 2), dtype=np.float64)
CITIES[0] = (45.5186, -73.5545)   # Montreal
CITIES[1] = (32.7942, -96.7475)   # Dallas
CITIES[2:, 0] = _unit(_CITY_SEED, 62, 1) * 120.0 - 55.0
CITIES[2:, 1] = _unit(_CITY_SEED, 62, 2) * 340.0 - 170.0


@dataclass
class Workers:
    a: np.ndarray            # abi.WORKER_A
    b: np.ndarray            # abi.WORKER_B
    lat: np.ndarray
    lon: np.ndarray
    model_strings: list      # catalogue order == model_id order
    status: np.ndarray       # orc_status ordinal (2 == Healthy)
    addresses: list | None = None
    addr_rank: np.ndarray | None = None

    def __len__(self):
        return len(self.a)


def zipf_levels(seed: int, n: int, stream: int, s: float = 1.1, levels: int = 1024) -> np.ndarray:
    """Zipf(s) over `levels` price levels: level k (1-based) with probability ~ k^-s."""
    p = np.arange(1, levels + 1, dtype=np.float64) ** (-s)
    cdf = np.cumsum(p / p.sum())
    return (np.minimum(np.searchsorted(cdf, _unit(seed, n, stream), side="right"), levels - 1) + 1).astype(np.uint32)


def wide_model_catalogue(n_models: int):
    """A permissionless pool's model strings: free-form NVML names (node.rs:463-484 normalises and substring-matches
    them), far more than 32 distinct ones.  The 16 catalogue entries come first (ids stay valid), then variants."""
    out = list(MODEL_CATALOGUE)
    fam = ["NVIDIA H100", "NVIDIA A100", "NVIDIA GeForce RTX 4090", "NVIDIA RTX A6000", "Tesla V100", "NVIDIA L40S",
           "AMD Instinct MI300X", "NVIDIA H200", "NVIDIA GeForce RTX 3090", "NVIDIA B200", "Quadro RTX 8000"]
    mem = [80000, 40000, 24564, 49140, 32768, 46068, 192000, 141000, 24576, 183000, 49152]
    suffix = ["", " PCIe", " SXM", " NVL", "-SXM4", " 80GB HBM3", " Ti"]
    seen = {m for m, _ in out}
    i = 0
    while len(out) < n_models:
        f = i % len(fam)
        k = i // len(fam)
        name = fam[f] + suffix[k % len(suffix)] + (f" rev{k // len(suffix)}" if k >= len(suffix) else "")
        i += 1
        if name in seen:
            continue
        seen.add(name)
        out.append((name, mem[f]))
    return out[:n_models]


def make_workers(n: int, seed: int = SEED_WORKERS, with_addresses: bool = False,
                 healthy_frac: float = 0.90, price: str | None = None, catalogue=None) -> Workers:
    """price: None (0: the reference has no prices), 'loguniform' (cfg3: 10..2000) or 'zipf' (cfg5: Zipf(1.1) over
    1024 levels) for the extension column; catalogue: [(model string, memory_mb)], default the 16-entry one."""
    catalogue = MODEL_CATALOGUE if catalogue is None else catalogue
    a = np.zeros(n, dtype=abi.WORKER_A)
    b = np.zeros(n, dtype=abi.WORKER_B)
    a["gpu_count"] = _choice(seed, n, 1, [1, 2, 4, 8], [0.40, 0.25, 0.20, 0.15])
    model = _choice(seed, n, 2, np.arange(len(catalogue)))
    a["model_id"] = model
    a["gpu_mem_mb"] = np.array([m for _, m in catalogue], dtype=np.uint32)[model]
    if price == "loguniform":
        b["ext_ask_price"] = np.exp(np.log(10) + _unit(SEED_EXT, n, 1) * np.log(200)).astype(np.uint32)
    elif price == "zipf":
        b["ext_ask_price"] = zipf_levels(SEED_EXT, n, 2)
    elif price is not None:
        raise ValueError(price)
    b["cpu_cores"] = _choice(seed, n, 3, [8, 16, 32, 64, 128])
    b["ram_mb"] = _choice(seed, n, 4, [32000, 64000, 128000, 256000, 512000, 1024000])
    b["storage_gb"] = _choice(seed, n, 5, [250, 500, 1000, 2000, 4000, 8000])
    no_specs = _unit(seed, n, 6) < 0.02
    no_mem = _unit(seed, n, 7) < 0.01
    healthy = _unit(seed, n, 8) < healthy_frac
    p2p = _unit(seed, n, 9) < 0.98
    no_loc = _unit(seed, n, 10) < 0.05
    city = _choice(seed, n, 11, np.arange(len(CITIES)))

    flags = np.full(n, abi.PM_W_HAS_SPECS | abi.PM_W_HAS_GPU | abi.PM_W_HAS_GPU_COUNT | abi.PM_W_HAS_GPU_MEM
                    | abi.PM_W_HAS_GPU_MODEL | abi.PM_W_HAS_CPU | abi.PM_W_HAS_CPU_CORES | abi.PM_W_HAS_RAM
                    | abi.PM_W_HAS_STORAGE, dtype=np.uint32)
    flags[no_mem] &= ~np.uint32(abi.PM_W_HAS_GPU_MEM)
    flags[no_specs] = 0
    flags[healthy] |= abi.PM_W_HEALTHY
    flags[p2p] |= abi.PM_W_P2P
    flags[~no_loc] |= abi.PM_W_HAS_LOC
    a["flags"] = flags
    # absent fields hold zeros, like a default-constructed row
    a["gpu_mem_mb"][no_mem | no_specs] = 0
    for col in ("gpu_count", "model_id"):
        a[col][no_specs] = 0
    for col in ("cpu_cores", "ram_mb", "storage_gb"):
        b[col][no_specs] = 0
    lat = np.where(no_loc, 0.0, CITIES[city, 0])
    lon = np.where(no_loc, 0.0, CITIES[city, 1])
    # non-healthy rows spread over the other NodeStatus ordinals
    other = _choice(seed, n, 12, [0, 1, 3, 4, 5, 6, 7])
    status = np.where(healthy, 2, other).astype(np.uint32)
    w = Workers(a=a, b=b, lat=lat, lon=lon, model_strings=[m for m, _ in catalogue], status=status)
    if with_addresses:
        h1 = splitmix64(seed, n, 13)
        h2 = splitmix64(seed, n, 14)
        h3 = splitmix64(seed, n, 15)
        case = splitmix64(seed, n, 16)
        addrs = []
        for i in range(n):
            hexs = f"{int(h1[i]):016x}{int(h2[i]):016x}{int(h3[i]) & 0xFFFFFFFF:08x}"
            c = int(case[i])
            # pseudo EIP-55 casing: opaque to the engine, exercises byte-order of 'A'-'F' vs 'a'-'f'
            hexs = "".join(ch.upper() if (ch.isalpha() and (c >> k) & 1) else ch for k, ch in enumerate(hexs))
            addrs.append("0x" + hexs)
        w.addresses = addrs
        order = sorted(range(n), key=lambda i: addrs[i].encode())
        rank = np.empty(n, dtype=np.uint32)
        rank[np.array(order, dtype=np.int64)] = np.arange(n, dtype=np.uint32)
        w.addr_rank = rank
    return w


@dataclass
class Asks:
    asks: np.ndarray         # abi.ASK, priority order
    opts: np.ndarray         # abi.GPU_OPT, CSR in ask order
    pattern_strings: list    # pattern_id order
    ext_price_cap: np.ndarray | None = None
    names: list = field(default_factory=list)

    def __len__(self):
        return len(self.asks)

    def requirement_string(self, t: int) -> str:
        """The ComputeRequirements string this row came from (for the string oracle)."""
        a = self.asks[t]
        parts = []
        for o in range(int(a["n_opts"])):
            q = self.opts[int(a["opt_off"]) + o]
            p = int(q["present"])
            if p & abi.PM_O_COUNT:
                parts.append(f"gpu:count={int(q['count'])}")
            if p & abi.PM_O_MODEL:
                parts.append(f"gpu:model={self.pattern_strings[int(q['pattern_id'])]}")
            if p & abi.PM_O_MEM:
                parts.append(f"gpu:memory_mb={int(q['memory_mb'])}")
            if p & abi.PM_O_MEM_MIN:
                parts.append(f"gpu:memory_mb_min={int(q['memory_mb_min'])}")
            if p & abi.PM_O_MEM_MAX:
                parts.append(f"gpu:memory_mb_max={int(q['memory_mb_max'])}")
            if p & abi.PM_O_TOT_MIN:
                parts.append(f"gpu:total_memory_min={int(q['total_memory_min'])}")
            if p & abi.PM_O_TOT_MAX:
                parts.append(f"gpu:total_memory_max={int(q['total_memory_max'])}")
        f = int(a["flags"])
        if f & abi.PM_A_REQ_CPU_CORES:
            parts.append(f"cpu:cores={int(a['cpu_cores'])}")
        if f & abi.PM_A_REQ_RAM:
            parts.append(f"ram_mb={int(a['ram_mb'])}")
        if f & abi.PM_A_REQ_STORAGE:
            parts.append(f"storage_gb={int(a['storage_gb'])}")
        return ";".join(parts)


def make_asks(n: int, kind: str = "mixed", seed: int = SEED_ASKS, group_sizes=None) -> Asks:
    """kind: 'uniform1' (cfg1), 'mixed' (cfg2/4), 'skewed' (cfg5: + 10% infeasible gpu:count=3).

    group_sizes: None -> min=max=1; else a list of (min,max) pairs drawn uniformly,
    after which rows are put into the reference's priority order (mod.rs:150-164)."""
    asks = np.zeros(n, dtype=abi.ASK)
    asks["flags"] = abi.PM_A_HAS_REQ
    if kind == "uniform1":
        n_opts = np.ones(n, dtype=np.uint32)
    else:
        n_opts = np.where(_unit(seed, n, 1) < 0.10, 2, 1).astype(np.uint32)
    asks["n_opts"] = n_opts
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(n_opts, out=off[1:])
    asks["opt_off"] = off[:-1]
    total = int(off[-1])
    opts = np.zeros(total, dtype=abi.GPU_OPT)
    owner = np.repeat(np.arange(n), n_opts)
    which = np.arange(total) - off[:-1][owner]          # 0 or 1 within the ask
    opts["present"] = abi.PM_O_COUNT
    if kind == "uniform1":
        opts["count"] = 1
    else:
        opts["count"] = _choice(seed, total, 2, [1, 2, 4, 8])
        with_model = _unit(seed, total, 3) < 0.50
        opts["present"][with_model] |= abi.PM_O_MODEL
        opts["pattern_id"][with_model] = _choice(seed, total, 4, np.arange(len(PATTERN_CATALOGUE)))[with_model]
        with_min = _unit(seed, total, 5) < 0.30
        opts["present"][with_min] |= abi.PM_O_MEM_MIN
        opts["memory_mb_min"][with_min] = _choice(seed, total, 6, [16000, 24000, 40000, 48000, 80000])[with_min]
        with_tot = _unit(seed, total, 7) < 0.05
        opts["present"][with_tot] |= abi.PM_O_TOT_MIN
        opts["total_memory_min"][with_tot] = _choice(seed, total, 8, [48000, 160000, 320000])[with_tot]
        # a second OR-option asks for a different count (as node.rs:764)
        second = which == 1
        opts["count"][second] = _choice(seed, total, 9, [1, 2, 4, 8])[second]
        scalar = _unit(seed, n, 10)
        ram = scalar < 0.20
        asks["flags"][ram] |= abi.PM_A_REQ_RAM
        asks["ram_mb"][ram] = _choice(seed, n, 11, [32000, 64000, 128000, 256000])[ram]
        sto = (scalar >= 0.20) & (scalar < 0.30)
        asks["flags"][sto] |= abi.PM_A_REQ_STORAGE
        asks["storage_gb"][sto] = _choice(seed, n, 12, [250, 500, 1000, 2000])[sto]
        cpu = (scalar >= 0.30) & (scalar < 0.40)
        asks["flags"][cpu] |= abi.PM_A_REQ_CPU | abi.PM_A_REQ_CPU_CORES
        asks["cpu_cores"][cpu] = _choice(seed, n, 13, [8, 16, 32, 64])[cpu]
        if kind == "skewed":
            infeasible = _unit(seed, n, 14) < 0.10
            opts["count"][infeasible[owner]] = 3     # no worker has 3 GPUs
    asks["min_group_size"] = 1
    asks["max_group_size"] = 1
    out = Asks(asks=asks, opts=opts, pattern_strings=list(PATTERN_CATALOGUE))
    if group_sizes is not None:
        gs = np.asarray(group_sizes, dtype=np.uint32)
        pick = _choice(seed, n, 15, np.arange(len(gs)))
        asks["min_group_size"] = gs[pick, 0]
        asks["max_group_size"] = gs[pick, 1]
        out = reorder_asks(out, priority_perm(asks))
    out.names = [f"cfg-{i}" for i in range(n)]
    return out


def priority_perm(asks: np.ndarray) -> np.ndarray:
    """mod.rs:150-164: stable by min_group_size desc, then with-requirements first."""
    key_req = (asks["flags"] & abi.PM_A_HAS_REQ) == 0
    idx = np.arange(len(asks))
    return np.lexsort((idx, key_req, -asks["min_group_size"].astype(np.int64)))


def reorder_asks(x: Asks, perm: np.ndarray) -> Asks:
    asks = x.asks[perm].copy()
    n_opts = asks["n_opts"].astype(np.int64)
    off = np.zeros(len(asks) + 1, dtype=np.int64)
    np.cumsum(n_opts, out=off[1:])
    src = np.concatenate([np.arange(int(o), int(o) + int(k)) for o, k in zip(asks["opt_off"], n_opts)]) \
        if len(asks) else np.zeros(0, dtype=np.int64)
    opts = x.opts[src.astype(np.int64)].copy() if len(src) else x.opts[:0].copy()
    asks["opt_off"] = off[:-1]
    cap = x.ext_price_cap[perm].copy() if x.ext_price_cap is not None else None
    return Asks(asks=asks, opts=opts, pattern_strings=x.pattern_strings, ext_price_cap=cap,
                names=[x.names[i] for i in perm] if x.names else [])


def intern_tables(workers: Workers, asks: Asks):
    """Intern the catalogues through the product interner; returns (bits, n_pat, n_models, words).

    model_id / pattern_id in the synthetic tables are catalogue positions, and the
    interner hands out ids in first-seen order, so interning the catalogues in
    order reproduces them."""
    from .engine import Interner

    it = Interner()
    for i, m in enumerate(workers.model_strings):
        assert it.model(m) == i
    for i, p in enumerate(asks.pattern_strings):
        assert it.pattern(p) == i
    out = it.table()
    it.close()
    return out
