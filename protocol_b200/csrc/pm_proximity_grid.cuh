// pm_proximity_grid.cuh — proximity group formation on ALL SMs (cooperative launch).
//
// Reference: NodeGroupsPlugin::try_form_new_groups with ProximityOptimizationPolicy{enabled:true} — the reference's
// DEFAULT policy (crates/orchestrator/src/plugins/node_groups/mod.rs:85-89), proximity branch :524-552,
// calculate_distance :218-231, sort_nodes_by_proximity :234-255 — run every 10 s by the management loop (:180-203).
//
// The loop is sequential per group by construction: a group's seed is the first remaining located node in canonical
// order, its members are the max-1 nearest remaining nodes, and the next group depends on who was removed.
// pm_proximity_sweep (pm_proximity.cuh) runs that loop in ONE CTA: every group costs a distance pass and max-1
// arg-mins over all remaining candidates of the configuration — 40 s for 250k candidates / 125k pair groups.
// Here the per-group work is spread over the whole chip and a group costs ONE grid barrier:
//
//   every thread of the grid owns a strided slice of the configuration's candidate list; per group
//     1. every CTA finds the seed by itself (same data, same answer — no exchange),
//     2. owners compute haversine(seed, candidate) for their live candidates and the CTA selects its k best
//        by (distance, list position) — k = max-1, in chunks of kPgTopK,
//     3. the per-CTA partial lists go to global memory, ONE grid.sync(),
//     4. every CTA merges all partials to the same k winners and marks them taken in its view of the list
//        (all CTAs write the same bits); CTA 0 also writes the group tables.
//   Rule that keeps the redundant control flow identical on every CTA: the list is only modified right after a grid
//   barrier, and only by decisions every CTA took from data that was complete before that barrier.
//
// Configurations of solo groups (max == 1) need no loop at all: located candidates first, then the others, each its
// own group (the seed rule, mod.rs:526-530) — an ordered partition, two barriers.  A configuration's candidate list
// is the sorted base segment merged with the workers handed down by earlier configurations; a long hand-down
// (e.g. the 750k workers a pair configuration leaves to the solo one) is rebuilt by an ordered compaction of
// {w : cur[w] == c} over the whole table instead of walking the linked list.
//
// Results are bit-identical to pm_proximity_sweep and to the oracle (tests/test_gpu_parity.py, test_gpu_proximity.py);
// the single-CTA kernel stays as the fallback when a cooperative launch is not possible.
#pragma once
#include <cooperative_groups.h>

#include "pm_proximity.cuh"

namespace pm {

namespace cg = cooperative_groups;

constexpr int kPgThreads = 512;
constexpr int kPgWarps = kPgThreads / 32;
constexpr uint32_t kPgTopK = 16;        // neighbours selected per grid barrier (larger groups take several)
constexpr uint32_t kPgXsSmem = 128;     // hand-downs up to this long are walked and sorted in shared memory
constexpr double kPgMax = 1.7976931348623157e308;   // f64::MAX: distance of a candidate without location
constexpr uint32_t kPgBatchK = 4;       // seed-parallel phase: neighbours listed per seed (group size - 1 plus spares)
constexpr uint32_t kPgHash = 4096;      // slots of the per-CTA "taken in this batch" set (<= 296 * 4 keys)
constexpr uint32_t kPgMaxSeeds = 304;   // >= co-resident CTAs of a B200 (148 SMs x 2)
constexpr uint32_t kPgUnroll = 8;       // candidates per thread in flight while the list streams from L2

struct GridProxParams {
  ProxParams p;
  double* part_d;        // [2][grid][kPgTopK] per-CTA partial selections, double-buffered by pass parity
  uint32_t* part_i;      // [2][grid][kPgTopK]
  uint32_t* cta_cnt;     // [2 * grid] per-CTA counts for the ordered compactions
  uint32_t* gctl;        // [8] [4..6] diagnostics (batches, groups formed in batches, groups formed one at a time); [0] leftover count  [1],[2] "a located candidate has |latitude| > 90", by configuration parity
  double* clat;          // [W] candidate-ordered copies of latitude / longitude / cos(latitude in radians), written with
  double* clon;          //     the list: the seed-parallel phase streams them instead of gathering through the index
  double* ccos;
  uint32_t n_workers;
};

struct PgShared {
  uint32_t u[8];
  uint32_t warp_cnt[kPgWarps];
  uint32_t warp_cnt2[kPgWarps];
  double red_d[kPgWarps];
  uint32_t red_i[kPgWarps];
  uint32_t picks[kPgThreads];     // positions chosen in the current pass (<= kPgTopK, or a chunk of list-order picks)
  uint32_t xs[kPgXsSmem];
  uint32_t seeds[kPgMaxSeeds];                 // seed-parallel phase: list positions of the batch's seeds
  uint32_t grp[kPgMaxSeeds * kPgBatchK];       // resolved groups: [seed, up to 3 neighbours] positions
  uint32_t cand[kPgMaxSeeds * kPgBatchK];      // the batch's neighbour lists, staged for the replay
  uint32_t taken[kPgHash];                     // positions taken by earlier groups of the batch (open addressing)
};

__device__ __noinline__ bool pg_set_has(const uint32_t* tab, uint32_t key) {
  uint32_t h = (key * 2654435761u) >> 20;
  for (;;) {
    const uint32_t v = tab[h];
    if (v == key) return true;
    if (v == kNone) return false;
    h = (h + 1u) & (kPgHash - 1u);
  }
}
__device__ __noinline__ void pg_set_put(uint32_t* tab, uint32_t key) {
  uint32_t h = (key * 2654435761u) >> 20;
  for (;;) {
    const uint32_t old = atomicCAS(&tab[h], kNone, key);
    if (old == kNone || old == key) return;
    h = (h + 1u) & (kPgHash - 1u);
  }
}

// calculate_distance (mod.rs:218-231) with the two cosines already taken: same operations in the same order as
// haversine_km, cos(lat1_rad) and cos(lat2_rad) are the cached values of exactly those expressions.
__device__ __noinline__ double haversine_km_cached(double lat1, double lon1, double cos1, double lat2, double lon2, double cos2) {
  const double kRadsPerDeg = 3.14159265358979323846264338327950288 / 180.0;
  const double delta_lat = __dmul_rn(__dsub_rn(lat2, lat1), kRadsPerDeg);
  const double delta_lon = __dmul_rn(__dsub_rn(lon2, lon1), kRadsPerDeg);
  const double s1 = sin(__dmul_rn(delta_lat, 0.5));
  const double s2 = sin(__dmul_rn(delta_lon, 0.5));
  double a = __dadd_rn(__dmul_rn(s1, s1), __dmul_rn(__dmul_rn(cos1, cos2), __dmul_rn(s2, s2)));
  if (a > 1.0) a = 1.0;   // antipodal rounding pushes a above 1 (NaN in the reference, no total order): clamped, DESIGN.md determinisation rule 7
  const double c = __dmul_rn(2.0, atan2(sqrt(a), sqrt(__dsub_rn(1.0, a))));
  return __dmul_rn(6371.0, c);
}

__device__ __forceinline__ uint32_t pg_ld(const uint32_t* p) { return __ldcg(p); }

// keep the kPgBatchK smallest (d, i) pairs, sorted; `mono` = positions arrive in increasing order (ties keep the holder)
template <bool MONO>
__device__ __forceinline__ void pg_top_insert(double (&bd)[kPgBatchK], uint32_t (&bi)[kPgBatchK], double d, uint32_t i) {
  const int L = (int)kPgBatchK - 1;
  if (!(bi[L] == kNone || d < bd[L] || (!MONO && d == bd[L] && i < bi[L]))) return;
  bd[L] = d;
  bi[L] = i;
#pragma unroll
  for (int q = L; q > 0; --q) {
    const bool up = bi[q - 1] == kNone || bd[q] < bd[q - 1] || (!MONO && bd[q] == bd[q - 1] && bi[q] < bi[q - 1]);
    if (up) {
      const double td = bd[q]; bd[q] = bd[q - 1]; bd[q - 1] = td;
      const uint32_t ti = bi[q]; bi[q] = bi[q - 1]; bi[q - 1] = ti;
    }
  }
}

// haversine distance >= R * |delta latitude| (a >= sin^2(dlat / 2) when both latitudes are real); the margins are
// those of the checker's latitude-pruned mode (oracle/): nine orders of magnitude above f64 rounding
__device__ __forceinline__ double pg_lat_bound(double slat, double lat) {
  const double dl = fabs(lat - slat) * (3.14159265358979323846264338327950288 / 180.0);
  return 6371.0 * dl * (1.0 - 1e-9) - 1e-9;
}

// smallest position >= start whose entry satisfies (e & mask) == want, or n.  Uniform over the CTA and — because every
// CTA reads the same list — over the grid.
__device__ __noinline__ uint32_t pg_find_first(PgShared& sh, const uint32_t* list, uint32_t n, uint32_t start,
                                                  uint32_t mask, uint32_t want) {
  for (uint32_t base = start; base < n; base += kPgThreads) {
    if (threadIdx.x == 0) sh.u[7] = kNone;
    __syncthreads();
    const uint32_t i = base + threadIdx.x;
    if (i < n && (pg_ld(list + i) & mask) == want) atomicMin(&sh.u[7], i);
    __syncthreads();
    const uint32_t f = sh.u[7];
    __syncthreads();
    if (f != kNone) return f;
  }
  return n;
}

// lexicographic (distance, position) minimum over the CTA; kNone when no thread has a candidate.  Uniform result in
// *out_d / return value.
__device__ __noinline__ uint32_t pg_block_argmin(PgShared& sh, double bd, uint32_t bi, double* out_d) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, bd, off);
    const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
    if (oi != kNone && (bi == kNone || od < bd || (od == bd && oi < bi))) { bd = od; bi = oi; }
  }
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (lane == 0) { sh.red_d[warp] = bd; sh.red_i[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    bd = lane < (uint32_t)kPgWarps ? sh.red_d[lane] : kPgMax;
    bi = lane < (uint32_t)kPgWarps ? sh.red_i[lane] : kNone;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, bd, off);
      const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (oi != kNone && (bi == kNone || od < bd || (od == bd && oi < bi))) { bd = od; bi = oi; }
    }
    if (lane == 0) { sh.red_d[0] = bd; sh.u[6] = bi; }
  }
  __syncthreads();
  const uint32_t r = sh.u[6];
  *out_d = sh.red_d[0];
  __syncthreads();
  return r;
}

// pop the CTA-wide smallest head of the threads' sorted lists; returns its position (kNone: all empty), *out_d its key
__device__ __forceinline__ uint32_t pg_pop_min(PgShared& sh, double (&bd)[kPgBatchK], uint32_t (&bi)[kPgBatchK], double* out_d) {
  const uint32_t wi = pg_block_argmin(sh, bd[0], bi[0], out_d);
  if (wi != kNone && bi[0] == wi) {
#pragma unroll
    for (uint32_t q = 0; q + 1 < kPgBatchK; ++q) { bd[q] = bd[q + 1]; bi[q] = bi[q + 1]; }
    bd[kPgBatchK - 1] = kPgMax;
    bi[kPgBatchK - 1] = kNone;
  }
  return wi;
}

// exclusive rank of this thread's flag inside the CTA and the CTA total
__device__ __noinline__ uint32_t pg_block_rank(PgShared& sh, bool flag, uint32_t* total) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t b = __ballot_sync(0xffffffffu, flag);
  if (lane == 0) sh.warp_cnt[warp] = (uint32_t)__popc(b);
  __syncthreads();
  uint32_t before = 0, tot = 0;
#pragma unroll
  for (int q = 0; q < kPgWarps; ++q) {
    const uint32_t v = sh.warp_cnt[q];
    if ((uint32_t)q < warp) before += v;
    tot += v;
  }
  __syncthreads();
  *total = tot;
  return before + (uint32_t)__popc(b & ((1u << lane) - 1u));
}

// two flags ranked in one pass (located / not located)
__device__ __noinline__ void pg_block_rank2(PgShared& sh, bool f0, bool f1, uint32_t* r0, uint32_t* r1,
                                               uint32_t* t0, uint32_t* t1) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t b0 = __ballot_sync(0xffffffffu, f0), b1 = __ballot_sync(0xffffffffu, f1);
  if (lane == 0) { sh.warp_cnt[warp] = (uint32_t)__popc(b0); sh.warp_cnt2[warp] = (uint32_t)__popc(b1); }
  __syncthreads();
  uint32_t bef0 = 0, bef1 = 0, tot0 = 0, tot1 = 0;
#pragma unroll
  for (int q = 0; q < kPgWarps; ++q) {
    const uint32_t v0 = sh.warp_cnt[q], v1 = sh.warp_cnt2[q];
    if ((uint32_t)q < warp) { bef0 += v0; bef1 += v1; }
    tot0 += v0; tot1 += v1;
  }
  __syncthreads();
  const uint32_t below = (1u << lane) - 1u;
  *r0 = bef0 + (uint32_t)__popc(b0 & below);
  *r1 = bef1 + (uint32_t)__popc(b1 & below);
  *t0 = tot0;
  *t1 = tot1;
}

// sum over the CTA
__device__ __noinline__ uint32_t pg_block_sum(PgShared& sh, uint32_t v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (lane == 0) sh.warp_cnt[warp] = v;
  __syncthreads();
  uint32_t tot = 0;
#pragma unroll
  for (int q = 0; q < kPgWarps; ++q) tot += sh.warp_cnt[q];
  __syncthreads();
  return tot;
}

// sum of cnt[0 .. upto) and of cnt[0 .. all), both CTA-uniform
__device__ __noinline__ void pg_prefix_of_ctas(PgShared& sh, const uint32_t* cnt, uint32_t upto, uint32_t all,
                                                  uint32_t* before, uint32_t* total) {
  uint32_t a = 0, t = 0;
  for (uint32_t b = threadIdx.x; b < all; b += kPgThreads) {
    const uint32_t v = pg_ld(cnt + b);
    if (b < upto) a += v;
    t += v;
  }
  *before = pg_block_sum(sh, a);
  *total = pg_block_sum(sh, t);
}

// one list entry and its candidate-ordered coordinate copies
__device__ __noinline__ void pg_put_entry(const GridProxParams& gp, uint32_t pos, uint32_t w, bool coords, uint32_t flag_slot) {
  const ProxParams& p = gp.p;
  const bool loc = (p.ev.wa[w].w & PM_W_HAS_LOC) != 0u;
  p.list[pos] = w | (loc ? kLocBit : 0u);
  if (coords && loc) {
    const double la = p.lat[w];
    if (!(fabs(la) <= 90.0)) atomicOr(gp.gctl + flag_slot, 1u);   // the latitude bound of the pruning needs real latitudes
    gp.clat[pos] = la;
    gp.clon[pos] = p.lon[w];
    gp.ccos[pos] = cos(__dmul_rn(la, 3.14159265358979323846264338327950288 / 180.0));
  }
}

__global__ void __launch_bounds__(kPgThreads) pm_proximity_grid(GridProxParams gp) {
  __shared__ PgShared sh;
  cg::grid_group grid = cg::this_grid();
  const ProxParams& p = gp.p;
  const uint32_t T = p.ev.n_asks, W = gp.n_workers;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t cta = blockIdx.x, ncta = gridDim.x;
  const uint32_t gtid = cta * kPgThreads + tid, nthr = ncta * kPgThreads;
  const bool lead = cta == 0;   // the CTA that writes the group tables
  uint32_t g = 0, mpos = 0, c_lo = 0, parity = 0, cfgno = 0;
  uint32_t n_batches = 0, n_batch_groups = 0, n_single = 0;   // diagnostics (pm_stats: n_tiles, n_rounds, n_build_launches)
  bool overflow = false;

  while (c_lo < T && !overflow) {
    // ---- next configuration that has members (or produces the empty group of min_group_size == 0)
    if (tid == 0) sh.u[0] = kNone;
    __syncthreads();
    {
      const uint32_t cc = c_lo + tid;
      if (cc < T && (pg_ld(p.base_len + cc) + pg_ld(p.xcount + cc) != 0u || p.amin[cc] == 0u)) atomicMin(&sh.u[0], cc);
    }
    __syncthreads();
    const uint32_t c = sh.u[0];
    __syncthreads();
    if (c == kNone) { c_lo += kPgThreads; continue; }
    const uint32_t mn = p.amin[c], mx = p.amax[c];
    const uint32_t bl = pg_ld(p.base_len + c), xc = pg_ld(p.xcount + c);
    const uint32_t n = bl + xc;
    const uint32_t* base = p.order + p.seg_start[c];
    // groups of 2..4 (1..3 neighbours) go through the seed-parallel phase first
    const bool batchable = mx >= 2u && mx - 1u < kPgBatchK && ncta >= 2u;
    // "a located candidate has |latitude| > 90" flag of this configuration; two slots used alternately, so that the lead
    // CTA can clear the next configuration's slot without racing the writers of this one
    const uint32_t flag_slot = 1u + (cfgno & 1u);

    // ---- the configuration's candidate list in canonical order, located candidates tagged
    if (xc <= kPgXsSmem) {
      // short hand-down: every CTA walks and sorts it by itself in shared memory, then the merge is spread over the grid
      if (xc) {
        if (tid == 0) {
          uint32_t j = 0;
          for (uint32_t x = pg_ld(p.xhead + c); x != kNone && j < kPgXsSmem; x = pg_ld(p.xnext + x)) sh.xs[j++] = x;
        }
        for (uint32_t j = xc + tid; j < kPgXsSmem; j += kPgThreads) sh.xs[j] = kNone;
        __syncthreads();
        for (uint32_t k = 2; k <= kPgXsSmem; k <<= 1)
          for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            if (tid < kPgXsSmem) {
              const uint32_t i = tid, l = i ^ j;
              if (l > i) {
                const uint32_t a = sh.xs[i], b = sh.xs[l];
                const bool up = (i & k) == 0;
                if ((a > b) == up) { sh.xs[i] = b; sh.xs[l] = a; }
              }
            }
            __syncthreads();
          }
      }
      for (uint32_t i = gtid; i < bl; i += nthr) {
        const uint32_t w = base[i];
        pg_put_entry(gp, i + lower_bound_u32(sh.xs, xc, w), w, mx >= 2u, flag_slot);
      }
      for (uint32_t j = gtid; j < xc; j += nthr) {
        const uint32_t w = sh.xs[j];
        pg_put_entry(gp, j + lower_bound_u32(base, bl, w), w, mx >= 2u, flag_slot);
      }
    } else {
      // long hand-down: the list is {w : cur[w] == c} in index order — an ordered compaction over the whole table
      const uint32_t chunk = ((W + ncta - 1) / ncta + kPgThreads - 1) / kPgThreads * kPgThreads;
      const uint32_t w_lo = min(W, cta * chunk), w_hi = min(W, w_lo + chunk);
      uint32_t mine = 0;
      for (uint32_t w = w_lo + tid; w < w_hi; w += kPgThreads) mine += pg_ld(p.cur + w) == c ? 1u : 0u;
      const uint32_t cnt = pg_block_sum(sh, mine);
      if (tid == 0) gp.cta_cnt[cta] = cnt;
      __threadfence();
      grid.sync();
      uint32_t off = 0, tot = 0;
      pg_prefix_of_ctas(sh, gp.cta_cnt, cta, ncta, &off, &tot);
      for (uint32_t w0 = w_lo; w0 < w_hi; w0 += kPgThreads) {
        const uint32_t w = w0 + tid;
        const bool f = w < w_hi && pg_ld(p.cur + w) == c;
        uint32_t tile = 0;
        const uint32_t r = pg_block_rank(sh, f, &tile);
        if (f) pg_put_entry(gp, off + r, w, mx >= 2u, flag_slot);
        off += tile;
      }
    }
    __threadfence();
    grid.sync();
    const bool prune = batchable && pg_ld(gp.gctl + flag_slot) == 0u;
    if (lead && tid == 0) gp.gctl[1u + ((cfgno + 1u) & 1u)] = 0u;   // nobody touches that slot before the next configuration
    ++cfgno;

    if (mx == 1u) {
      // ---- solo groups: located candidates first, then the others, each its own group (mod.rs:526-530)
      const uint32_t chunk = ((n + ncta - 1) / ncta + kPgThreads - 1) / kPgThreads * kPgThreads;
      const uint32_t i_lo = min(n, cta * chunk), i_hi = min(n, i_lo + chunk);
      const uint32_t extra = mn == 0u ? 1u : 0u;                       // the trailing empty group (:564,:606)
      if ((uint64_t)g + n + extra > (uint64_t)p.group_cap) {
        overflow = true;
      } else {
        uint32_t mine = 0;
        for (uint32_t i = i_lo + tid; i < i_hi; i += kPgThreads) mine += (pg_ld(p.list + i) & kLocBit) ? 1u : 0u;
        const uint32_t cnt = pg_block_sum(sh, mine);
        if (tid == 0) gp.cta_cnt[ncta + cta] = cnt;
        __threadfence();
        grid.sync();
        uint32_t loc_before = 0, n_loc = 0;
        pg_prefix_of_ctas(sh, gp.cta_cnt + ncta, cta, ncta, &loc_before, &n_loc);
        uint32_t unloc_before = i_lo - loc_before;
        for (uint32_t i0 = i_lo; i0 < i_hi; i0 += kPgThreads) {
          const uint32_t i = i0 + tid;
          const bool in = i < i_hi;
          const uint32_t e = in ? pg_ld(p.list + i) : 0u;
          uint32_t r_loc, r_un, t_loc, t_un;
          pg_block_rank2(sh, in && (e & kLocBit), in && !(e & kLocBit), &r_loc, &r_un, &t_loc, &t_un);
          if (in) {
            const uint32_t gi = (e & kLocBit) ? loc_before + r_loc : n_loc + unloc_before + r_un;
            const uint32_t w = e & kIdxMask;
            p.group_ask[g + gi] = c;
            p.group_off[g + gi] = mpos + gi;
            p.members[mpos + gi] = w;
            p.worker_group[w] = g + gi;
            p.worker_ask[w] = c;
          }
          loc_before += t_loc;
          unloc_before += t_un;
        }
        g += n;
        mpos += n;
        if (extra) {
          if (lead && tid == 0) { p.group_ask[g] = c; p.group_off[g] = mpos; }
          ++g;
        }
      }
      if (lead && tid == 0) { p.base_len[c] = 0; p.xcount[c] = 0; p.xhead[c] = kNone; }
      __threadfence();
      grid.sync();
      c_lo = c + 1;
      continue;
    }

    // ---- the group loop of mod.rs:507-609 for this configuration (max_group_size >= 2)
    uint32_t remaining = n, ploc = 0, pany = 0;

    // Seed-parallel phase.  The next B seeds are the first B live located candidates (canonical order).  CTA j lists,
    // for seed j, its kPgBatchK nearest live candidates as of the START of the batch — over ALL candidates, so a batch is
    // one pass of distance work per CTA and one grid barrier for up to `grid` groups.  Every CTA then replays the
    // sequential loop over the batch: a seed already taken by an earlier group of the batch is no seed any more; a
    // group's members are the first k entries of its list not taken earlier in the batch — exactly the k nearest of what
    // is left, because the list is the head of the order "by (distance, position)" over a superset.  A list that runs
    // out of spare entries ends the batch there (the first group of a batch can never run out).  Full-size groups only:
    // the last few of a configuration, larger groups and unlocated seeds go through the per-group loop below.
    if (batchable) {
      const uint32_t k = mx - 1u;
      for (;;) {
        const uint32_t room = p.group_cap > g ? p.group_cap - g : 0u;
        uint32_t B = min(min(ncta, (uint32_t)kPgMaxSeeds), min(remaining / mx, room));
        if (B < 2u || ploc >= n) break;
        // the batch's seeds: first B live located positions from ploc (same scan on every CTA)
        uint32_t nb = 0, pos = ploc;
        while (nb < B && pos < n) {
          const uint32_t i = pos + tid;
          const bool f = i < n && (pg_ld(p.list + i) & (kTakenBit | kLocBit)) == kLocBit;
          uint32_t tile = 0;
          const uint32_t r = pg_block_rank(sh, f, &tile);
          if (f && nb + r < B) sh.seeds[nb + r] = i;
          nb = min(B, nb + tile);
          pos += kPgThreads;
        }
        __syncthreads();
        if (nb == 0u) { ploc = n; break; }   // nobody located is left: the per-group loop takes the rest in list order
        parity ^= 1u;
        if (cta < nb) {
          const uint32_t sp = sh.seeds[cta];
          const double slat = __ldcg(gp.clat + sp), slon = __ldcg(gp.clon + sp), scos = __ldcg(gp.ccos + sp);
          double bd[kPgBatchK];
          uint32_t bi[kPgBatchK];
          // Pruning.  An upper bound T of the CTA's final 4th-best distance comes from a cheap first pass: every thread
          // keeps its 4 candidates with the smallest LATITUDE lower bound, their exact distances are taken, and the
          // 4th smallest of those over the CTA is T (exact distances of real candidates: the true 4th best cannot be
          // larger).  The second pass then computes a haversine only where the lower bound does not already exceed T —
          // a candidate with bound > T is farther than T and can neither enter the list nor win a tie.
          double T = kPgMax;
          // Both passes stream the whole list from L2; kPgUnroll candidates per thread are in flight at a time (the
          // loads of a chunk are issued before anything is decided), otherwise every candidate costs a full L2 round trip.
          // (Staggering the CTAs' starting offsets was measured and made it slower: 2.15 s -> 2.92 s at 1M nodes.)
          // Code size matters here: the first version of these loops inlined the haversine and the insertion network at
          // every unrolled site, the kernel grew to 13k instructions (208 KB) and 58 % of its warp-stall samples were
          // instruction-cache misses (profiles/r02_proximity_1m_stalls.txt).  Hence: loads unrolled, processing rolled,
          // one out-of-line haversine.
          const uint32_t n_chunks = (n + kPgThreads * kPgUnroll - 1) / (kPgThreads * kPgUnroll);
          if (prune) {
            double lv[kPgBatchK];
            uint32_t li[kPgBatchK];
#pragma unroll
            for (uint32_t q = 0; q < kPgBatchK; ++q) { lv[q] = kPgMax; li[q] = kNone; bd[q] = kPgMax; bi[q] = kNone; }
            for (uint32_t ck = 0; ck < n_chunks; ++ck) {
              const uint32_t base = ck * (kPgThreads * kPgUnroll) + tid;
              uint32_t e[kPgUnroll];
              double la[kPgUnroll];
#pragma unroll
              for (uint32_t u = 0; u < kPgUnroll; ++u) {
                const uint32_t i = base + u * kPgThreads;
                e[u] = (i < n && i != sp) ? pg_ld(p.list + i) : kTakenBit;
              }
#pragma unroll
              for (uint32_t u = 0; u < kPgUnroll; ++u)
                la[u] = (e[u] & (kTakenBit | kLocBit)) == kLocBit ? __ldcg(gp.clat + base + u * kPgThreads) : 0.0;
#pragma unroll 1
              for (uint32_t u = 0; u < kPgUnroll; ++u)
                if ((e[u] & (kTakenBit | kLocBit)) == kLocBit) pg_top_insert<true>(lv, li, pg_lat_bound(slat, la[u]), base + u * kPgThreads);
            }
#pragma unroll 1
            for (uint32_t q = 0; q < kPgBatchK; ++q) {
              const uint32_t i = li[q];
              if (i != kNone)
                pg_top_insert<false>(bd, bi, haversine_km_cached(slat, slon, scos, __ldcg(gp.clat + i), __ldcg(gp.clon + i), __ldcg(gp.ccos + i)), i);
            }
            uint32_t got = 0;
            double wd = kPgMax;
            for (uint32_t r = 0; r < kPgBatchK; ++r)
              if (pg_pop_min(sh, bd, bi, &wd) != kNone) ++got;
            if (got == kPgBatchK) T = wd;       // fewer than 4 located candidates exist: no pruning, unlocated ones may be needed
          }
#pragma unroll
          for (uint32_t q = 0; q < kPgBatchK; ++q) { bd[q] = kPgMax; bi[q] = kNone; }
          for (uint32_t ck = 0; ck < n_chunks; ++ck) {
            const uint32_t base = ck * (kPgThreads * kPgUnroll) + tid;
            uint32_t e[kPgUnroll];
            double la[kPgUnroll];
#pragma unroll
            for (uint32_t u = 0; u < kPgUnroll; ++u) {
              const uint32_t i = base + u * kPgThreads;
              e[u] = (i < n && i != sp) ? pg_ld(p.list + i) : kTakenBit;
            }
#pragma unroll
            for (uint32_t u = 0; u < kPgUnroll; ++u)
              la[u] = (e[u] & (kTakenBit | kLocBit)) == kLocBit ? __ldcg(gp.clat + base + u * kPgThreads) : 0.0;
#pragma unroll 1
            for (uint32_t u = 0; u < kPgUnroll; ++u) {
              if (e[u] & kTakenBit) continue;
              const uint32_t i = base + u * kPgThreads;
              double d = kPgMax;
              if (e[u] & kLocBit) {
                if (pg_lat_bound(slat, la[u]) > T) continue;
                d = haversine_km_cached(slat, slon, scos, la[u], __ldcg(gp.clon + i), __ldcg(gp.ccos + i));
              } else if (T != kPgMax) {
                continue;                        // at least 4 located candidates exist: one without location cannot be among the 4 nearest
              }
              pg_top_insert<true>(bd, bi, d, i);   // positions only grow in this thread: an equal distance never displaces a holder
            }
          }
          uint32_t* my_i = gp.part_i + ((size_t)parity * ncta + cta) * kPgTopK;
          for (uint32_t r = 0; r < kPgBatchK; ++r) {   // merge the per-thread lists: the winner pops its head
            double wd;
            const uint32_t wi = pg_pop_min(sh, bd, bi, &wd);
            if (tid == 0) my_i[r] = wi;
          }
        }
        __threadfence();
        grid.sync();
        // replay of the sequential loop over the batch, identically on every CTA (warp 0; the set is per CTA)
        for (uint32_t h = tid; h < kPgHash; h += kPgThreads) sh.taken[h] = kNone;
        {   // the batch's lists into shared memory in one go (a load per seed inside the replay would serialise L2 round trips)
          const uint32_t* all_i = gp.part_i + (size_t)parity * ncta * kPgTopK;
          for (uint32_t t = tid; t < nb * kPgBatchK; t += kPgThreads) sh.cand[t] = __ldcg(all_i + (size_t)(t / kPgBatchK) * kPgTopK + (t % kPgBatchK));
        }
        __syncthreads();
        if (warp == 0) {
          uint32_t formed = 0, next = nb;
          for (uint32_t j = 0; j < nb; ++j) {
            const uint32_t sp = sh.seeds[j];
            if (pg_set_has(sh.taken, sp)) continue;                        // became a member of an earlier group
            const uint32_t cand = lane < kPgBatchK ? sh.cand[j * kPgBatchK + lane] : kNone;
            const bool free_ = cand != kNone && !pg_set_has(sh.taken, cand);
            const uint32_t fb = __ballot_sync(0xffffffffu, free_);
            if ((uint32_t)__popc(fb) < k) { next = j; break; }             // out of spares: this seed opens the next batch
            const uint32_t rank = (uint32_t)__popc(fb & ((1u << lane) - 1u));
            if (free_ && rank < k) { sh.grp[formed * kPgBatchK + 1u + rank] = cand; pg_set_put(sh.taken, cand); }
            if (lane == 0) { sh.grp[formed * kPgBatchK] = sp; pg_set_put(sh.taken, sp); }
            __syncwarp();
            ++formed;
          }
          if (lane == 0) { sh.u[4] = formed; sh.u[5] = next < nb ? sh.seeds[next] : sh.seeds[nb - 1u] + 1u; }
        }
        __syncthreads();
        const uint32_t formed = sh.u[4];
        ploc = sh.u[5];
        // marks (after the barrier, same bits from every CTA) and, on the lead CTA, the group tables
        for (uint32_t t = tid; t < formed * mx; t += kPgThreads) {
          const uint32_t gi = t / mx, m = t % mx;
          const uint32_t q = sh.grp[gi * kPgBatchK + m];
          atomicOr(p.list + q, kTakenBit);
          if (lead) {
            const uint32_t w = pg_ld(p.list + q) & kIdxMask;
            p.members[mpos + t] = w;
            p.worker_group[w] = g + gi;
            p.worker_ask[w] = c;
            if (m == 0u) { p.group_ask[g + gi] = c; p.group_off[g + gi] = mpos + gi * mx; }
          }
        }
        __syncthreads();
        g += formed;
        mpos += formed * mx;
        remaining -= formed * mx;
        ++n_batches;
        n_batch_groups += formed;
        if (formed == 0u) break;
      }
    }

    for (;;) {
      if (remaining < mn) break;                                           // :507 / :517
      uint32_t seed_pos = n;
      bool seed_loc = false;
      if (remaining) {
        if (ploc < n) {
          seed_pos = pg_find_first(sh, p.list, n, ploc, kTakenBit | kLocBit, kLocBit);
          ploc = seed_pos;
        }
        if (seed_pos < n) {
          seed_loc = true;
        } else {                                                           // .or(compatible_nodes.first())
          seed_pos = pg_find_first(sh, p.list, n, pany, kTakenBit, 0u);
          pany = seed_pos;
        }
      }
      const bool have_seed = seed_pos < n;
      if (have_seed && !seed_loc && remaining >= mx) {
        // Nobody located is left: from here on every group is its seed (the first live candidate) plus the next max-1
        // live candidates in list order (the stable sort leaves f64::MAX distances in list order) — consecutive runs of
        // `max` live candidates.  All full groups at once by an ordered partition (two barriers) instead of a barrier
        // per group; what remains afterwards (fewer than `max`) goes through the loop below.
        const uint32_t n_full = remaining / mx;
        if ((uint64_t)g + n_full > (uint64_t)p.group_cap) { overflow = true; break; }
        const uint32_t chunk = ((n + ncta - 1) / ncta + kPgThreads - 1) / kPgThreads * kPgThreads;
        const uint32_t i_lo = min(n, cta * chunk), i_hi = min(n, i_lo + chunk);
        uint32_t mine = 0;
        for (uint32_t i = i_lo + tid; i < i_hi; i += kPgThreads) mine += (pg_ld(p.list + i) & kTakenBit) ? 0u : 1u;
        const uint32_t cnt = pg_block_sum(sh, mine);
        if (tid == 0) gp.cta_cnt[ncta + cta] = cnt;
        __threadfence();
        grid.sync();
        uint32_t before = 0, total = 0;
        pg_prefix_of_ctas(sh, gp.cta_cnt + ncta, cta, ncta, &before, &total);   // total == remaining
        for (uint32_t i0 = i_lo; i0 < i_hi; i0 += kPgThreads) {
          const uint32_t i = i0 + tid;
          const uint32_t e = i < i_hi ? pg_ld(p.list + i) : kTakenBit;
          const bool live = (e & kTakenBit) == 0u;
          uint32_t tile = 0;
          const uint32_t r = before + pg_block_rank(sh, live, &tile);
          if (live && r < n_full * mx) {
            const uint32_t w = e & kIdxMask, gi = r / mx;
            atomicOr(p.list + i, kTakenBit);
            p.members[mpos + r] = w;
            p.worker_group[w] = g + gi;
            p.worker_ask[w] = c;
            if (r % mx == 0u) { p.group_ask[g + gi] = c; p.group_off[g + gi] = mpos + r; }
          }
          before += tile;
        }
        g += n_full;
        mpos += n_full * mx;
        remaining -= n_full * mx;
        n_single += n_full;
        __threadfence();
        grid.sync();          // the marks of every CTA are in place before anybody looks at the list again
        continue;
      }
      const uint32_t size = have_seed ? min(mx, remaining) : 0u;
      if (size < mn) break;                                                // :564
      if (g >= p.group_cap) { overflow = true; break; }
      if (lead && tid == 0) { p.group_ask[g] = c; p.group_off[g] = mpos; }
      if (!have_seed) {                                                    // min_group_size == 0: the empty group
        ++g;
        break;                                                             // :606 no progress
      }
      const uint32_t seed_w = pg_ld(p.list + seed_pos) & kIdxMask;
      uint32_t need = size - 1u, written = 1u;   // members written so far (the seed is written with the first marks)
      bool seed_marked = false;
      if (need && seed_loc) {
        const double slat = __ldcg(gp.clat + seed_pos), slon = __ldcg(gp.clon + seed_pos), scos = __ldcg(gp.ccos + seed_pos);
        bool first_pass = true;
        while (need) {
          const uint32_t kk = min(need, kPgTopK);
          const bool keep = first_pass ? need > kk || kk > 1u : true;     // distances are re-read in later rounds
          parity ^= 1u;
          double* my_d = gp.part_d + ((size_t)parity * ncta + cta) * kPgTopK;
          uint32_t* my_i = gp.part_i + ((size_t)parity * ncta + cta) * kPgTopK;
          // 2. this CTA's kk best by (distance, position): round r takes the smallest pair above round r-1's
          double prev_d = -1.0;
          uint32_t prev_i = 0;
          for (uint32_t r = 0; r < kk; ++r) {
            double bd = kPgMax;
            uint32_t bi = kNone;
            for (uint32_t i = gtid; i < n; i += nthr) {
              const uint32_t e = pg_ld(p.list + i);
              if ((e & kTakenBit) || i == seed_pos) continue;
              double d;
              if (first_pass && r == 0u) {
                d = (e & kLocBit) ? haversine_km_cached(slat, slon, scos, __ldcg(gp.clat + i), __ldcg(gp.clon + i), __ldcg(gp.ccos + i)) : kPgMax;
                if (keep) p.dist[i] = d;                                  // owner-private: read back by this thread only
              } else {
                d = p.dist[i];
              }
              if (r && !(d > prev_d || (d == prev_d && i > prev_i))) continue;
              if (bi == kNone || d < bd) { bd = d; bi = i; }              // i increases: ties keep the smaller position
            }
            double wd;
            const uint32_t wi = pg_block_argmin(sh, bd, bi, &wd);
            if (tid == 0) { my_d[r] = wd; my_i[r] = wi; }
            prev_d = wd;
            prev_i = wi;
            if (wi == kNone) {                                            // this CTA has no more candidates
              if (tid == 0) for (uint32_t q = r + 1; q < kk; ++q) { my_d[q] = kPgMax; my_i[q] = kNone; }
              break;
            }
          }
          __threadfence();
          grid.sync();                                                     // 3. the group's one barrier
          // 4. every CTA merges all partial lists to the same kk winners
          const double* all_d = gp.part_d + (size_t)parity * ncta * kPgTopK;
          const uint32_t* all_i = gp.part_i + (size_t)parity * ncta * kPgTopK;
          prev_d = -1.0;
          prev_i = 0;
          for (uint32_t r = 0; r < kk; ++r) {
            double bd = kPgMax;
            uint32_t bi = kNone;
            for (uint32_t e = tid; e < ncta * kk; e += kPgThreads) {
              const uint32_t slot = (e / kk) * kPgTopK + (e % kk);
              const uint32_t i = __ldcg(all_i + slot);
              if (i == kNone) continue;
              const double d = __ldcg(all_d + slot);
              if (r && !(d > prev_d || (d == prev_d && i > prev_i))) continue;
              if (bi == kNone || d < bd || (d == bd && i < bi)) { bd = d; bi = i; }
            }
            double wd;
            const uint32_t wi = pg_block_argmin(sh, bd, bi, &wd);
            if (tid == 0) sh.picks[r] = wi;
            prev_d = wd;
            prev_i = wi;
          }
          __syncthreads();
          // marks: only now, after the barrier — every CTA sets the same bits; the lead CTA writes the tables
          if (tid < kk) {
            const uint32_t pos = sh.picks[tid];
            if (pos != kNone) {
              atomicOr(p.list + pos, kTakenBit);
              if (lead) {
                const uint32_t w = pg_ld(p.list + pos) & kIdxMask;
                p.members[mpos + written + tid] = w;
                p.worker_group[w] = g;
                p.worker_ask[w] = c;
              }
            }
          }
          if (!seed_marked && tid == kPgThreads - 1) {
            atomicOr(p.list + seed_pos, kTakenBit);
            if (lead) { p.members[mpos] = seed_w; p.worker_group[seed_w] = g; p.worker_ask[seed_w] = c; }
          }
          seed_marked = true;
          __syncthreads();
          need -= kk;
          written += kk;
          first_pass = false;
        }
      } else {
        // no distances involved: a group of one, or a seed without location (then nobody left has one): the next
        // `need` live candidates in list order, taken in chunks of the pick buffer; marks after a barrier as above
        uint32_t pos = pany;
        bool first = true;
        while (first || need) {
          uint32_t got = 0;
          const uint32_t want = min(need, (uint32_t)kPgThreads);
          while (got < want && pos < n) {
            const uint32_t i = pos + tid;
            const bool live = i < n && i != seed_pos && (pg_ld(p.list + i) & kTakenBit) == 0u;
            uint32_t tile = 0;
            const uint32_t r = pg_block_rank(sh, live, &tile);
            if (live && got + r < want) sh.picks[got + r] = i;
            if (got + tile > want) {
              // the chunk ends inside this tile: continue right after its last pick (uniform: read back below)
              __syncthreads();
              const uint32_t last = sh.picks[want - 1];
              got = want;
              pos = last + 1u;
              __syncthreads();
            } else {
              got += tile;
              pos += kPgThreads;
            }
          }
          __syncthreads();
          __threadfence();
          grid.sync();
          for (uint32_t j = tid; j < got; j += kPgThreads) {
            const uint32_t q = sh.picks[j];
            atomicOr(p.list + q, kTakenBit);
            if (lead) {
              const uint32_t w = pg_ld(p.list + q) & kIdxMask;
              p.members[mpos + written + j] = w;
              p.worker_group[w] = g;
              p.worker_ask[w] = c;
            }
          }
          if (!seed_marked && tid == kPgThreads - 1) {
            atomicOr(p.list + seed_pos, kTakenBit);
            if (lead) { p.members[mpos] = seed_w; p.worker_group[seed_w] = g; p.worker_ask[seed_w] = c; }
          }
          seed_marked = true;
          __syncthreads();
          need -= got;
          written += got;
          first = false;
          if (got == 0u) break;   // cannot happen (remaining >= size); keeps the loop finite
        }
      }
      remaining -= size;
      mpos += size;
      ++g;
      ++n_single;
    }

    // ---- leftovers move on to their next feasible configuration
    if (!overflow) {
      for (uint32_t i = gtid; i < n; i += nthr) {
        const uint32_t e = pg_ld(p.list + i);
        if ((e & kTakenBit) == 0u) p.popped[atomicAdd(gp.gctl, 1u)] = e & kIdxMask;
      }
      __threadfence();
      grid.sync();
      const uint32_t npop = pg_ld(gp.gctl);
      const uint32_t gwarp = cta * kPgWarps + warp, nwarp = ncta * kPgWarps;
      for (uint32_t i = gwarp; i < npop; i += nwarp) {
        const uint32_t w = pg_ld(p.popped + i);
        const WorkerReg wr = make_worker(p.ev.wa[w], p.ev.wb[w]);
        uint32_t found = kNone;
        for (uint32_t cbase = c + 1; cbase < T; cbase += 32) {
          const uint32_t c2 = cbase + lane;
          bool ok = false;
          if (c2 < T) ok = ask_meets(p.ev.asks[c2], p.ev.opts, wr, p.ev.bits, p.ev.words);
          const uint32_t b = __ballot_sync(0xffffffffu, ok);
          if (b) { found = cbase + (uint32_t)__ffs(b) - 1u; break; }
        }
        if (lane == 0) {
          p.cur[w] = found;
          if (found != kNone) {
            const uint32_t old = atomicExch(&p.xhead[found], w);
            p.xnext[w] = old;
            atomicAdd(&p.xcount[found], 1u);
          }
          atomicAdd(&p.out_counts[2], 1u);
        }
      }
      __threadfence();
      grid.sync();
      if (lead && tid == 0) { p.base_len[c] = 0; p.xcount[c] = 0; p.xhead[c] = kNone; gp.gctl[0] = 0; }
    }
    c_lo = c + 1;
  }
  if (lead && tid == 0) {
    p.out_counts[0] = g;
    p.out_counts[1] = mpos;
    if (overflow) p.out_counts[3] = 1u;
    gp.gctl[4] = n_batches; gp.gctl[5] = n_batch_groups; gp.gctl[6] = n_single;
    if (g < p.group_cap + 1) p.group_off[g] = mpos;
  }
}

}  // namespace pm
