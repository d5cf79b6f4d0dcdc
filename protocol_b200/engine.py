"""Python host wrapper over the C ABI (include/prime_match.h).

Mirrors, for the hot path only, the reference's NodeGroupsPlugin surface:
`Engine.match()` is one `try_form_new_groups` pass
(crates/orchestrator/src/plugins/node_groups/mod.rs:478-628) over the resident
worker and ask tables.  All compute happens in the CUDA library; this module is
plumbing (numpy views over pinned/host buffers, error mapping).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import abi
from ._lib import PrimeMatchError, load


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


class Interner:
    """pm_interner: model-string interning + acceptance table (node.rs:463-484)."""

    def __init__(self):
        self._lib = load()
        self._h = self._lib.pm_interner_create()
        if not self._h:
            raise MemoryError("pm_interner_create failed")

    def close(self):
        if self._h:
            self._lib.pm_interner_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def model(self, s: str) -> int:
        return self._lib.pm_intern_model(self._h, s.encode())

    def pattern(self, s: str) -> int:
        return self._lib.pm_intern_pattern(self._h, s.encode())

    def table(self):
        bits = C.POINTER(C.c_uint32)()
        npat, nmod, words = C.c_uint32(), C.c_uint32(), C.c_uint32()
        rc = self._lib.pm_interner_table(self._h, C.byref(bits), C.byref(npat), C.byref(nmod), C.byref(words))
        if rc != abi.PM_OK:
            raise PrimeMatchError(rc, "pm_interner_table")
        n = max(npat.value, 1) * words.value
        arr = np.ctypeslib.as_array(bits, shape=(n,)).copy()
        return arr, npat.value, nmod.value, words.value

    def parse(self, s: str, max_opts: int = 64):
        """ComputeRequirements::from_str -> (ask_row: np.void ASK, opts: np.ndarray GPU_OPT)."""
        ask = abi.PmAsk()
        opts = (abi.PmGpuOpt * max_opts)()
        n = C.c_uint32()
        err = C.create_string_buffer(256)
        rc = self._lib.pm_parse_requirements(s.encode(), self._h, C.byref(ask), opts, max_opts, C.byref(n), err, 256)
        if rc != abi.PM_OK:
            raise PrimeMatchError(rc, err.value.decode())
        a = np.zeros(1, dtype=abi.ASK)
        for f in abi.ASK.names:
            a[f] = getattr(ask, f)
        o = np.zeros(n.value, dtype=abi.GPU_OPT)
        for i in range(n.value):
            for f in abi.GPU_OPT.names:
                o[f][i] = getattr(opts[i], f)
        return a[0], o


def sort_configs(min_group_size, has_requirements) -> np.ndarray:
    """Priority order of NodeGroupsPlugin::new_with_policy (mod.rs:150-164)."""
    lib = load()
    mn = np.ascontiguousarray(min_group_size, dtype=np.uint32)
    hr = np.ascontiguousarray(has_requirements, dtype=np.uint8)
    perm = np.empty(len(mn), dtype=np.uint32)
    rc = lib.pm_sort_configs(_ptr(mn), _ptr(hr), len(mn), _ptr(perm))
    if rc != abi.PM_OK:
        raise PrimeMatchError(rc, "pm_sort_configs")
    return perm


@dataclass
class MatchResult:
    worker_group: np.ndarray
    worker_ask: np.ndarray
    group_ask: np.ndarray
    group_off: np.ndarray
    group_members: np.ndarray
    ask_best: np.ndarray
    ask_count: np.ndarray
    stats: dict

    @property
    def n_groups(self) -> int:
        return len(self.group_ask)

    def groups(self):
        """[(ask, [workers...])] in creation order."""
        return [
            (int(self.group_ask[g]), self.group_members[self.group_off[g]:self.group_off[g + 1]].tolist())
            for g in range(self.n_groups)
        ]


class Engine:
    def __init__(self, device: int = 0, cost_tile_bytes: int = 0, timing: bool = False,
                 shard_first: int = 0, shard_count: int = 0, stream: int = 0):
        self._lib = load()
        cfg = abi.PmCfg(abi.PM_ABI_VERSION, device, abi.PM_CFG_TIMING if timing else 0, 0,
                        cost_tile_bytes, shard_first, shard_count, stream or None)
        h = C.c_void_p()
        rc = self._lib.pm_create(C.byref(cfg), C.byref(h))
        if rc != abi.PM_OK:
            raise PrimeMatchError(rc, (self._lib.pm_last_error(None) or b"").decode())
        self._h = h
        self.n_workers = 0
        self.n_asks = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != abi.PM_OK:
            raise PrimeMatchError(rc, (self._lib.pm_last_error(self._h) or b"").decode())

    # ---- tables -----------------------------------------------------------
    def set_asks(self, asks: np.ndarray, opts: np.ndarray):
        asks = np.ascontiguousarray(asks, dtype=abi.ASK)
        opts = np.ascontiguousarray(opts, dtype=abi.GPU_OPT)
        self._check(self._lib.pm_set_asks(self._h, _ptr(asks), len(asks), _ptr(opts), len(opts)))
        self.n_asks = len(asks)

    def set_model_table(self, bits: np.ndarray, n_patterns: int, n_models: int, words: int):
        bits = np.ascontiguousarray(bits, dtype=np.uint32)
        self._check(self._lib.pm_set_model_table(self._h, _ptr(bits), n_patterns, n_models, words))

    def set_workers(self, a: np.ndarray, b: np.ndarray):
        a = np.ascontiguousarray(a, dtype=abi.WORKER_A)
        b = np.ascontiguousarray(b, dtype=abi.WORKER_B)
        assert len(a) == len(b)
        self._check(self._lib.pm_set_worker_count(self._h, len(a)))
        self._check(self._lib.pm_upsert_workers(self._h, _ptr(a), _ptr(b), 0, len(a)))
        self.n_workers = len(a)
        self.sync()

    def upsert_workers(self, a: np.ndarray, b: np.ndarray, first: int = 0, sync: bool = True):
        """a/b must stay alive (and ideally pinned) until sync()."""
        self._check(self._lib.pm_upsert_workers(self._h, _ptr(a), _ptr(b), first, len(a)))
        if sync:
            self.sync()

    def set_locations(self, lat: np.ndarray, lon: np.ndarray):
        lat = np.ascontiguousarray(lat, dtype=np.float64)
        lon = np.ascontiguousarray(lon, dtype=np.float64)
        self._check(self._lib.pm_set_worker_locations(self._h, _ptr(lat), _ptr(lon), 0, len(lat)))
        self.sync()

    def set_addr_rank(self, rank: np.ndarray):
        rank = np.ascontiguousarray(rank, dtype=np.uint32)
        self._check(self._lib.pm_set_worker_addr_rank(self._h, _ptr(rank), 0, len(rank)))
        self.sync()

    def resize_workers(self, n: int):
        self._check(self._lib.pm_resize_workers(self._h, n))
        self.n_workers = n

    def update_workers(self, idx, a, b, lat=None, lon=None):
        """Scatter changed rows into the resident table (pm_update_workers)."""
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        a = np.ascontiguousarray(a, dtype=abi.WORKER_A)
        b = np.ascontiguousarray(b, dtype=abi.WORKER_B)
        la = np.ascontiguousarray(lat, dtype=np.float64) if lat is not None else None
        lo = np.ascontiguousarray(lon, dtype=np.float64) if lon is not None else None
        self._check(self._lib.pm_update_workers(self._h, _ptr(idx), _ptr(a), _ptr(b), _ptr(la) if la is not None else None,
                                                _ptr(lo) if lo is not None else None, len(idx)))
        self.sync()

    def table_version(self) -> int:
        return int(self._lib.pm_table_version(self._h))

    def set_flags(self, idx: np.ndarray, flags: np.ndarray):
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        flags = np.ascontiguousarray(flags, dtype=np.uint32)
        self._check(self._lib.pm_set_flags(self._h, _ptr(idx), _ptr(flags), len(idx)))
        self.sync()

    # ---- north-star extension columns (PM_MODE_AUCTION only) ---------------
    def set_price_caps(self, caps: np.ndarray):
        caps = np.ascontiguousarray(caps, dtype=np.uint32)
        self._check(self._lib.pm_set_ask_price_caps(self._h, _ptr(caps), len(caps)))

    def set_worker_reputation(self, reputation: np.ndarray, first: int = 0):
        """north_star worker column `reputation` (u32 per worker); read by PM_MODE_AUCTION through the per-ask floor."""
        reputation = np.ascontiguousarray(reputation, dtype=np.uint32)
        self._check(self._lib.pm_set_worker_reputation(self._h, _ptr(reputation), first, len(reputation)))

    def set_min_reputation(self, floors: np.ndarray):
        floors = np.ascontiguousarray(floors, dtype=np.uint32)
        self._check(self._lib.pm_set_ask_min_reputation(self._h, _ptr(floors), len(floors)))

    def set_auction_params(self, cost_scale: int = 1, eps_start: int = 1, eps_div: int = 4):
        self._check(self._lib.pm_set_auction_params(self._h, cost_scale, eps_start, eps_div))

    # ---- the pass ---------------------------------------------------------
    def match(self, mode: int = abi.PM_MODE_FIRST_FIT):
        self._check(self._lib.pm_match(self._h, mode))

    def match_local(self, mode: int = abi.PM_MODE_FIRST_FIT):
        self._check(self._lib.pm_match_local(self._h, mode))

    def match_finish(self, mode: int = abi.PM_MODE_FIRST_FIT):
        self._check(self._lib.pm_match_finish(self._h, mode))

    def sync(self):
        self._check(self._lib.pm_stream_sync(self._h))

    def stats(self) -> dict:
        st = abi.PmStats()
        self._check(self._lib.pm_get_stats(self._h, C.byref(st)))
        return st.as_dict()

    def device_buffer(self, which: int):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self._lib.pm_device_buffer(self._h, which, C.byref(p), C.byref(n)))
        return p.value, n.value

    def cost_tile(self, t0: int, nt: int, n_cols: int | None = None) -> np.ndarray:
        """The materialised int64 cost rows [t0, t0+nt) x this engine's worker range."""
        n_cols = self.n_workers if n_cols is None else n_cols
        out = np.empty((nt, n_cols), dtype=np.int64)
        self._check(self._lib.pm_build_cost_tile(self._h, t0, nt, _ptr(out)))
        return out

    def fetch(self, copy: bool = True) -> MatchResult:
        r = abi.PmResult()
        self._check(self._lib.pm_fetch_result(self._h, C.byref(r)))
        return _result_views(r, copy)

    # ---- multi-GPU, one process per GPU -------------------------
    def set_shard(self, first: int, count: int):
        self._check(self._lib.pm_set_shard(self._h, first, count))

    def attach_comm(self, comm: "Comm | None"):
        """With a communicator attached, match() evaluates this rank's share of the workers, exchanges once
        (one packed all-gather inside the library) and runs the replicated resolution sweep."""
        self._check(self._lib.pm_attach_comm(self._h, comm._h if comm is not None else None))
        self._comm = comm   # keep it alive


def _result_views(r: "abi.PmResult", copy: bool) -> MatchResult:
    def view(ptr, n, dt):
        if n == 0:
            return np.zeros(0, dtype=dt)
        arr = np.ctypeslib.as_array(ptr, shape=(n,))
        return arr.copy() if copy else arr

    return MatchResult(
        worker_group=view(r.worker_group, r.n_workers, np.uint32),
        worker_ask=view(r.worker_ask, r.n_workers, np.uint32),
        group_ask=view(r.group_ask, r.n_groups, np.uint32),
        group_off=view(r.group_off, r.n_groups + 1, np.uint32),
        group_members=view(r.group_members, r.n_members, np.uint32),
        ask_best=view(r.ask_best, r.n_asks, np.int64),
        ask_count=view(r.ask_count, r.n_asks, np.uint32),
        stats=r.stats.as_dict(),
    )


class Comm:
    """pm_comm: this rank's end of the communicator the sharded pass exchanges through (NCCL inside the library)."""

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * abi.PM_COMM_ID_BYTES)()
        lib = load()
        rc = lib.pm_comm_unique_id(buf)
        if rc != abi.PM_OK:
            raise PrimeMatchError(rc, (lib.pm_last_error(None) or b"").decode())
        return bytes(buf)

    def __init__(self, unique_id: bytes, n_ranks: int, rank: int, device: int):
        self._lib = load()
        assert len(unique_id) == abi.PM_COMM_ID_BYTES
        buf = (C.c_uint8 * abi.PM_COMM_ID_BYTES).from_buffer_copy(unique_id)
        h = C.c_void_p()
        rc = self._lib.pm_comm_create(buf, n_ranks, rank, device, C.byref(h))
        if rc != abi.PM_OK:
            raise PrimeMatchError(rc, (self._lib.pm_last_error(None) or b"").decode())
        self._h = h
        self.n_ranks, self.rank = n_ranks, rank

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pm_comm_destroy(self._h)
            self._h = None


class MultiEngine:
    """pm_multi: one process driving several GPUs (the form the orchestrator's single management loop calls)."""

    def __init__(self, devices, cost_tile_bytes: int = 0, timing: bool = False):
        self._lib = load()
        cfg = abi.PmCfg(abi.PM_ABI_VERSION, 0, abi.PM_CFG_TIMING if timing else 0, 0, cost_tile_bytes, 0, 0, None)
        devs = (C.c_int32 * len(devices))(*devices)
        h = C.c_void_p()
        rc = self._lib.pm_multi_create(C.byref(cfg), devs, len(devices), C.byref(h))
        if rc != abi.PM_OK:
            raise PrimeMatchError(rc, (self._lib.pm_multi_last_error(None) or b"").decode())
        self._h = h
        self.n = len(devices)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pm_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != abi.PM_OK:
            raise PrimeMatchError(rc, (self._lib.pm_multi_last_error(self._h) or b"").decode())

    def set_asks(self, asks, opts):
        asks = np.ascontiguousarray(asks, dtype=abi.ASK)
        opts = np.ascontiguousarray(opts, dtype=abi.GPU_OPT)
        self._check(self._lib.pm_multi_set_asks(self._h, _ptr(asks), len(asks), _ptr(opts), len(opts)))

    def set_model_table(self, bits, n_patterns, n_models, words):
        bits = np.ascontiguousarray(bits, dtype=np.uint32)
        self._check(self._lib.pm_multi_set_model_table(self._h, _ptr(bits), n_patterns, n_models, words))

    def set_workers(self, a, b):
        a = np.ascontiguousarray(a, dtype=abi.WORKER_A)
        b = np.ascontiguousarray(b, dtype=abi.WORKER_B)
        self._check(self._lib.pm_multi_set_worker_count(self._h, len(a)))
        self._check(self._lib.pm_multi_upsert_workers(self._h, _ptr(a), _ptr(b), 0, len(a)))
        self._check(self._lib.pm_multi_sync(self._h))

    def set_locations(self, lat, lon):
        lat = np.ascontiguousarray(lat, dtype=np.float64)
        lon = np.ascontiguousarray(lon, dtype=np.float64)
        self._check(self._lib.pm_multi_set_worker_locations(self._h, _ptr(lat), _ptr(lon), 0, len(lat)))
        self._check(self._lib.pm_multi_sync(self._h))

    def set_addr_rank(self, rank):
        rank = np.ascontiguousarray(rank, dtype=np.uint32)
        self._check(self._lib.pm_multi_set_worker_addr_rank(self._h, _ptr(rank), 0, len(rank)))
        self._check(self._lib.pm_multi_sync(self._h))

    def match(self, mode: int = abi.PM_MODE_FIRST_FIT):
        self._check(self._lib.pm_multi_match(self._h, mode))

    def fetch(self, copy: bool = True) -> MatchResult:
        r = abi.PmResult()
        self._check(self._lib.pm_multi_fetch_result(self._h, C.byref(r)))
        return _result_views(r, copy)

    def stats(self, i: int = 0) -> dict:
        st = abi.PmStats()
        e = self._lib.pm_multi_engine(self._h, i)
        rc = self._lib.pm_get_stats(e, C.byref(st))
        if rc != abi.PM_OK:
            raise PrimeMatchError(rc, "pm_get_stats")
        return st.as_dict()


def pinned_empty(n: int, dtype) -> np.ndarray:
    """numpy array over pm_alloc_pinned memory (freed when the array dies)."""
    lib = load()
    dt = np.dtype(dtype)
    nbytes = max(n * dt.itemsize, 1)
    p = lib.pm_alloc_pinned(nbytes)
    if not p:
        raise MemoryError("pm_alloc_pinned failed")
    buf = (C.c_char * nbytes).from_address(p)
    arr = np.frombuffer(buf, dtype=dt, count=n)

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr

        def __del__(self):
            try:
                lib.pm_free_pinned(self.ptr)
            except Exception:
                pass

    # keep the owner alive as long as any view of `buf` lives
    buf._pm_owner = _Owner(p)
    return arr
