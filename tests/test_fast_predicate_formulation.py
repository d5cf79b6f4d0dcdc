"""The arithmetic of the engine's fast predicate (csrc/pm_device.cuh DevOptF, pm_kernels.cuh eval_opt_fast /
pm_ask_convert), restated in numpy u32 arithmetic and held against the CPU oracle pair by pair: every clause ends as a
sign bit, clauses are OR-ed, OR-options AND-ed, and the packed int64 cost is formed as (sign >> 31) * difference + base.
This pins the FORMULATION on the CPU (key layout, masked equality, 31-model acceptance words, the negation that turns
"non-zero" into a sign, the neutral option row of asks without GPU options); the CUDA code itself is checked against
the oracle by the -m gpu parity tests."""
import numpy as np
import pytest

from oracle import pm_oracle as orc
from protocol_b200 import abi, synth

U = np.uint32
CAND, NEVER, TOTINV = U(1 << 31), U(1 << 30), U(1 << 29)


def key_bits(f):
    f = f.astype(np.uint32)
    return ((f >> U(3)) & U(0x3FF)) | np.where(f & CAND, U(1 << 10), U(0)) | np.where(f & NEVER, U(1 << 11), U(0)) | \
        np.where(f & TOTINV, U(1 << 12), U(0))


def workers_reg(wa, wb):
    fl = wa["flags"].astype(np.uint32)
    hc, hm = (fl & U(abi.PM_W_HAS_GPU_COUNT)) != 0, (fl & U(abi.PM_W_HAS_GPU_MEM)) != 0
    cand = (fl & U(abi.PM_W_HEALTHY | abi.PM_W_P2P | abi.PM_W_ASSIGNED)) == U(abi.PM_W_HEALTHY | abi.PM_W_P2P)
    flags = (fl & U(0x1FFFFFFF)) | np.where(cand, CAND, U(0)) | np.where(hc & hm, U(0), TOTINV)
    count_eff = np.where(hc, wa["gpu_count"], 0).astype(np.uint32)
    mid = np.where((fl & U(abi.PM_W_HAS_GPU_MODEL)) != 0, wa["model_id"], 0).astype(np.uint32)
    with np.errstate(over="ignore"):
        tot = (wa["gpu_count"].astype(np.uint32) * wa["gpu_mem_mb"].astype(np.uint32)).astype(np.uint32)
    return dict(key=key_bits(flags) | (count_eff << U(13)), mem_eff=np.where(hm, wa["gpu_mem_mb"], 0).astype(np.uint32),
                tot=tot, tot_keep=np.where(hc & hm, U(0xFFFFFFFF), U(0)), cores=wb["cpu_cores"].astype(np.uint32),
                ram=wb["ram_mb"].astype(np.uint32), storage=wb["storage_gb"].astype(np.uint32),
                mword=mid // U(31), mmask=(U(1) << (mid % U(31))).astype(np.uint32), price=wb["ext_ask_price"].astype(np.uint32))


def device_bits(bits, npat, nmod, words):
    dwords = max((nmod + 30) // 31, 1)
    tbl = np.zeros((npat + 1, dwords), dtype=np.uint32)
    tbl[0, :] = 0x7FFFFFFF
    b = np.asarray(bits, dtype=np.uint32).reshape(max(npat, 1), words)
    m = np.arange(nmod)
    for p in range(npat):
        acc = ((b[p, m >> 5] >> (m & 31).astype(np.uint32)) & 1).astype(np.uint32)
        np.bitwise_or.at(tbl[p + 1], m // 31, acc << (m % 31).astype(np.uint32))
    return tbl, dwords


def convert_ask(a, opts):
    """pm_ask_convert for one ask -> (cpu, ram, storage, [DevOptF dicts]) (fast form only)."""
    f = int(a["flags"])
    has_req = bool(f & abi.PM_A_HAS_REQ)
    need = 1 << 31
    cpu = ram = sto = 0
    if has_req:
        need |= abi.PM_W_HAS_SPECS
        if f & abi.PM_A_REQ_CPU:
            need |= abi.PM_W_HAS_CPU
        if (f & abi.PM_A_REQ_CPU) and (f & abi.PM_A_REQ_CPU_CORES):
            need |= abi.PM_W_HAS_CPU_CORES
            cpu = int(a["cpu_cores"])
        if f & abi.PM_A_REQ_RAM:
            need |= abi.PM_W_HAS_RAM
            ram = int(a["ram_mb"])
        if f & abi.PM_A_REQ_STORAGE:
            need |= abi.PM_W_HAS_STORAGE
            sto = int(a["storage_gb"])
        if int(a["n_opts"]):
            need |= abi.PM_W_HAS_GPU
    if int(a["max_group_size"]) == 0:
        need |= 1 << 30
    rows = []
    n_eff = int(a["n_opts"]) if has_req else 0
    for o in range(n_eff):
        q = opts[int(a["opt_off"]) + o]
        pr = int(q["present"])
        xneed = 0
        mem_lo, mem_hi, tot_lo, tot_hi = 0, 0xFFFFFFFF, 0, 0xFFFFFFFF
        if pr & abi.PM_O_MEM:
            mem_lo = max(mem_lo, int(q["memory_mb"]))
        if pr & abi.PM_O_MEM_MIN:
            mem_lo = max(mem_lo, int(q["memory_mb_min"]))
        if pr & abi.PM_O_MEM_MAX:
            mem_hi = int(q["memory_mb_max"])
        if pr & (abi.PM_O_MEM | abi.PM_O_MEM_MIN | abi.PM_O_MEM_MAX):
            xneed |= abi.PM_W_HAS_GPU_MEM
        if pr & abi.PM_O_TOT_MIN:
            tot_lo = int(q["total_memory_min"])
        if pr & abi.PM_O_TOT_MAX:
            tot_hi = int(q["total_memory_max"])
        if mem_lo > mem_hi:
            xneed |= 1 << 30
            mem_lo, mem_hi = 0, 0xFFFFFFFF
        if tot_lo > tot_hi:
            xneed |= 1 << 29
            tot_lo, tot_hi = 0, 0xFFFFFFFF
        prow = 0
        if pr & abi.PM_O_MODEL:
            xneed |= abi.PM_W_HAS_GPU_MODEL
            prow = int(q["pattern_id"]) + 1
        m = int(key_bits(np.array([need | xneed], dtype=np.uint32))[0])
        v = m
        if pr & abi.PM_O_COUNT:
            m |= 0xFFFF << 13
            v |= (int(q["count"]) & 0xFFFF) << 13
        rows.append(dict(m=m, v=v, mem_lo=mem_lo, mem_hi=min(mem_hi, 0x7FFFFFFF), tot_lo=tot_lo, tot_hi=min(tot_hi, 0x7FFFFFFF), prow=prow))
    if not rows:
        m = int(key_bits(np.array([need], dtype=np.uint32))[0])
        rows.append(dict(m=m, v=m, mem_lo=0, mem_hi=0x7FFFFFFF, tot_lo=0, tot_hi=0x7FFFFFFF, prow=0))
    return cpu, ram, sto, rows


def fast_cost_matrix(wa, wb, asks, opts, bits, npat, nmod, words):
    w = workers_reg(wa, wb)
    tbl, dwords = device_bits(bits, npat, nmod, words)
    W = len(wa)
    gw = np.arange(W, dtype=np.uint32)
    out = np.empty((len(asks), W), dtype=np.int64)
    with np.errstate(over="ignore"):
        for t, a in enumerate(asks):
            cpu, ram, sto, rows = convert_ask(a, opts)
            um = None
            for q in rows:
                word = tbl[q["prow"], w["mword"]]
                z = ((w["key"] & U(q["m"])) ^ U(q["v"])) | (~word & w["mmask"])
                assert (z < U(1 << 31)).all()
                r = (w["mem_eff"] - U(q["mem_lo"])) | (U(q["mem_hi"]) - w["mem_eff"]) | \
                    (((w["tot"] - U(q["tot_lo"])) | (U(q["tot_hi"]) - w["tot"])) & w["tot_keep"])
                u = (U(0) - z) | r
                um = u if um is None else (um & u)
            fail = um | (w["cores"] - U(cpu)) | (w["ram"] - U(ram)) | (w["storage"] - U(sto))
            b = fail >> U(31)
            lo = b * (U(0xFFFFFFFF) - gw) + gw
            hi = b * (U(0x7FFFFFFF) - w["price"]) + w["price"]
            out[t] = (hi.astype(np.uint64) << np.uint64(32) | lo.astype(np.uint64)).astype(np.int64)
    return out


@pytest.mark.parametrize("kind,price", [("mixed", None), ("skewed", "zipf"), ("uniform1", "loguniform")])
def test_fast_formulation_equals_the_oracle_on_synthetic_tables(kind, price):
    w = synth.make_workers(1500, price=price)
    a = synth.make_asks(160, kind, group_sizes=[(1, 1), (0, 0), (2, 3)])   # max_group_size == 0 rows are masked
    bits, npat, nmod, words = synth.intern_tables(w, a)
    ev = orc.soa_eval_matrix(w.a, w.b, a.asks, a.opts, bits, words, 0, len(a), 0, len(w), want_cost=True)
    got = fast_cost_matrix(w.a, w.b, a.asks, a.opts, bits, npat, nmod, words)
    assert np.array_equal(got, ev["cost"])


def test_fast_formulation_with_many_models_and_every_clause():
    """> 31 models (several acceptance words), exact / min / max / total memory clauses, empty intervals, missing fields."""
    from helpers import TableBuilder
    import kat_vectors as kv

    tb = TableBuilder()
    reqs = ["gpu:count=8;gpu:model=H100", "gpu:memory_mb=40000", "gpu:memory_mb_min=30000;gpu:memory_mb_max=50000",
            "gpu:total_memory_min=100000;gpu:total_memory_max=400000", "gpu:count=2;gpu:count=4;gpu:model=a100",
            "cpu:cores=32;ram_mb=64000;storage_gb=500", "gpu:model=zz-nothing", "gpu:total_memory_min=1", None,
            "gpu:count=0", "gpu:memory_mb_min=90000;gpu:memory_mb_max=100"]
    for r in reqs:
        try:
            tb.add_config(r)
        except Exception:
            pass   # the parser rejects min > max: that row is simply not part of the table
    rng = np.random.default_rng(5)
    models = [m for m, _ in synth.wide_model_catalogue(70)]
    for i in range(600):
        s = dict(gpu_count=int(rng.choice([1, 2, 4, 8])) if rng.random() > .1 else None,
                 gpu_model=models[int(rng.integers(0, 70))] if rng.random() > .1 else None,
                 gpu_mem=int(rng.choice([24000, 40000, 80000])) if rng.random() > .1 else None,
                 cpu_cores=int(rng.choice([8, 32, 64])) if rng.random() > .1 else None,
                 ram=int(rng.choice([32000, 64000, 256000])) if rng.random() > .1 else None,
                 storage=int(rng.choice([250, 1000])) if rng.random() > .1 else None)
        tb.add_node(s if rng.random() > .05 else None, healthy=rng.random() > .1, p2p=rng.random() > .05, assigned=rng.random() < .05)
    t = tb.tables()
    assert t["n_models"] > 31
    ev = orc.soa_eval_matrix(t["wa"], t["wb"], t["asks"], t["opts"], t["bits"], t["words"], 0, len(t["asks"]), 0, len(t["wa"]), want_cost=True)
    got = fast_cost_matrix(t["wa"], t["wb"], t["asks"], t["opts"], t["bits"], t["n_patterns"], t["n_models"], t["words"])
    assert np.array_equal(got, ev["cost"])
