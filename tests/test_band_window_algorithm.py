"""The window formulation used by the experimental device sweep (protocol_b200/csrc/pm_proximity_band.cuh), restated in
numpy and held against the checker: per configuration the located candidates are ranked by (latitude, list position);
a group gathers the live entries of ranks [rs - half, rs + half] around its seed, takes its k nearest by
(distance, list position), and accepts when the window is the whole order or the latitude band the window covers
completely has a lower bound R * band above the k-th distance — else the window grows fourfold; taken entries stay as
tombstones and are squeezed out when they outnumber the live ones.  This is the ALGORITHM of the kernel (same accept
rule, same tie rule, same compaction trigger), not its CUDA mechanics: it guards the logic until the kernel has run."""
import numpy as np
import pytest

from oracle import pm_oracle as orc
from protocol_b200 import abi, synth

R = 6371.0
DEG = 3.14159265358979323846264338327950288 / 180.0


def haversine(lat1, lon1, lat2, lon2):
    """calculate_distance (mod.rs:218-231), same operation order in f64 (numpy does not contract to FMA)."""
    lat1r, lat2r = lat1 * DEG, lat2 * DEG
    dlat, dlon = (lat2 - lat1) * DEG, (lon2 - lon1) * DEG
    s1, s2 = np.sin(dlat * 0.5), np.sin(dlon * 0.5)
    a = s1 * s1 + (np.cos(lat1r) * np.cos(lat2r)) * (s2 * s2)
    return R * (2.0 * np.arctan2(np.sqrt(a), np.sqrt(1.0 - a)))


def form_groups_window(wa, wb, asks, opts, bits, words, lat, lon, half0=4, compact_min=8):
    """Returns (cfg, off, members-with-each-group-sorted-by-index).  half0 / compact_min are tiny so that window growth
    and compaction happen on test-sized inputs (the kernel uses 512 and 2048)."""
    W = len(wa)
    cand = (wa["flags"] & (abi.PM_W_HEALTHY | abi.PM_W_P2P | abi.PM_W_ASSIGNED)) == (abi.PM_W_HEALTHY | abi.PM_W_P2P)
    located = (wa["flags"] & abi.PM_W_HAS_LOC) != 0
    taken_global = ~cand
    cfg, off, members = [], [0], []
    for c in range(len(asks)):
        mn, mx = int(asks["min_group_size"][c]), int(asks["max_group_size"][c])
        remaining_ids = np.flatnonzero(~taken_global)
        if len(remaining_ids) == 0 and mn > 0:
            continue
        lst = np.array([w for w in remaining_ids if orc.soa_compatible(wa[w], wb[w], asks[c], opts, bits, words)], dtype=np.int64)
        n = len(lst)
        if n == 0 and mn > 0:
            continue
        taken = np.zeros(n, bool)
        loc = located[lst] if n else np.zeros(0, bool)
        # latitude order of the located entries: (latitude, list position)
        lat_ord = np.array(sorted(np.flatnonzero(loc), key=lambda i: (lat[lst[i]], i)), dtype=np.int64)
        lat_key = lat[lst[lat_ord]] if len(lat_ord) else np.zeros(0)
        rank_of = np.full(n, -1, np.int64)
        rank_of[lat_ord] = np.arange(len(lat_ord))
        nloc, dead_loc = len(lat_ord), 0
        remaining, ploc, pany = n, 0, 0
        while True:
            if remaining < mn:
                break
            seed, seed_loc = n, False
            if remaining:
                while ploc < n and not (loc[ploc] and not taken[ploc]):
                    ploc += 1
                if ploc < n:
                    seed, seed_loc = ploc, True
                else:
                    while pany < n and taken[pany]:
                        pany += 1
                    seed = pany
            have_seed = seed < n
            size = (min(mx, remaining) if mx else 1) if have_seed else 0
            if size < mn:
                break
            grp = []
            if have_seed:
                taken[seed] = True
                grp.append(int(lst[seed]))
                k, n_sel = size - 1, 0
                if k and seed_loc:
                    slat, slon = lat[lst[seed]], lon[lst[seed]]
                    rs, half = int(rank_of[seed]), half0
                    while True:
                        lo, hi = max(rs - half, 0), min(nloc - 1, rs + half)
                        win = [int(i) for i in lat_ord[lo:hi + 1] if not taken[i]]
                        all_in = lo == 0 and hi == nloc - 1
                        if len(win) < k and not all_in:
                            half *= 4
                            continue
                        band = np.inf
                        if lo != 0:
                            band = min(band, slat - lat_key[lo])
                        if hi != nloc - 1:
                            band = min(band, lat_key[hi] - slat)
                        d = {i: float(haversine(slat, slon, lat[lst[i]], lon[lst[i]])) for i in win}
                        picked = sorted(win, key=lambda i: (d[i], i))[:min(k, len(win))]
                        accept = all_in
                        if not accept and len(picked) == k:
                            lb = R * (band * DEG) * (1.0 - 1e-9) - 1e-9
                            accept = lb > d[picked[-1]]
                        if not accept:
                            half *= 4
                            continue
                        for i in picked:
                            taken[i] = True
                            grp.append(int(lst[i]))
                        n_sel = len(picked)
                        break
                    dead_loc += 1 + n_sel
                elif seed_loc:
                    dead_loc += 1
                if k > n_sel:                                      # located ones used up, or a seed without location
                    q = pany
                    while len(grp) < 1 + k and q < n:
                        if not taken[q]:
                            taken[q] = True
                            grp.append(int(lst[q]))
                        q += 1
            cfg.append(c)
            members.extend(sorted(grp))
            off.append(len(members))
            remaining -= size
            if size == 0:
                break
            if mx > 1 and nloc > compact_min and dead_loc * 2 > nloc:   # squeeze the tombstones out
                keep = np.array([i for i in lat_ord[:nloc] if not taken[i]], dtype=np.int64)
                lat_ord[:len(keep)] = keep
                lat_key[:len(keep)] = lat[lst[keep]]
                rank_of[keep] = np.arange(len(keep))
                nloc, dead_loc = len(keep), 0
        taken_global[lst[taken]] = True
    return np.array(cfg), np.array(off), np.array(members)


@pytest.mark.parametrize("seed", [31, 32, 33])
@pytest.mark.parametrize("where", ["cities", "scattered", "one_point"])
def test_window_formulation_forms_the_checkers_groups(seed, where):
    sizes = [(1, 1), (2, 2), (2, 4), (3, 3), (1, 3), (4, 8), (0, 2), (0, 0), (2, 5)]
    w = synth.make_workers(600, seed=seed)
    a = synth.make_asks(24, "mixed", seed=seed + 1, group_sizes=sizes)
    bits, npat, nmod, words = synth.intern_tables(w, a)
    rng = np.random.default_rng(seed)
    lat, lon = w.lat.copy(), w.lon.copy()
    if where == "scattered":
        lat = lat + rng.normal(0, 3.0, len(lat)).clip(-20, 20)
        lon = lon + rng.normal(0, 5.0, len(lon))
        lat[rng.random(len(lat)) < 0.05] = 48.8566
    elif where == "one_point":
        lat[:] = 45.5
        lon[:] = -73.5
    want = orc.soa_form_groups(w.a, w.b, a.asks, a.opts, bits, words, lat=lat, lon=lon, proximity=True)
    cfg, off, members = form_groups_window(w.a, w.b, a.asks, a.opts, bits, words, lat, lon)
    assert np.array_equal(want.cfg, cfg) and np.array_equal(want.off, off)
    assert np.array_equal(want.members, members)
    assert len(want) > 20
