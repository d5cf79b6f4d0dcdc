// pm_proximity.cuh — group formation with ProximityOptimizationPolicy{enabled:true}
// for configurations with max_group_size > 1.
//
// Reference: NodeGroupsPlugin::try_form_new_groups, proximity branch
// (crates/orchestrator/src/plugins/node_groups/mod.rs:524-552), calculate_distance
// (:218-231, haversine, f64) and sort_nodes_by_proximity (:234-255, stable sort by
// distance to the seed, missing location = f64::MAX).
//
// The loop is sequential per group by construction (each group's seed is the first
// remaining located node, each group takes the max-1 nearest remaining nodes), so
// this is one resident CTA walking configurations in priority order; all the
// per-group work — distances to the seed for every remaining compatible worker,
// k rounds of (distance, position) arg-min — is block-parallel.  The compatible set
// of a configuration comes from the T x W evaluation pass (first feasible ask per
// worker) plus the workers handed down by earlier configurations, exactly as in
// pm_sweep.
#pragma once
#include "pm_kernels.cuh"

namespace pm {

constexpr uint32_t kTakenBit = 1u << 31;
constexpr uint32_t kLocBit = 1u << 30;
constexpr uint32_t kTentBit = 1u << 29;   // merge sweep: tentatively in the current batch
constexpr uint32_t kIdxMask = 0x1FFFFFFFu;
constexpr int kProxThreads = 1024;

struct ProxParams {
  EvalParams ev;
  const double* lat;
  const double* lon;
  uint32_t* cur;              // [W] current ask of each worker (first feasible ask on entry)
  uint32_t* base_len;         // [T] members still in the sorted base segment
  const uint32_t* seg_start;  // [T+1]
  const uint32_t* order;      // [W] workers sorted by (ask, index)
  uint32_t* xhead;            // [T] workers handed down into the ask
  uint32_t* xnext;            // [W]
  uint32_t* xcount;           // [T]
  const uint32_t* amin;
  const uint32_t* amax;
  uint32_t* list;             // [W] scratch: member list of the current ask (canonical order)
  uint32_t* xs;               // [pow2 >= W] scratch: sorted handed-down workers
  double* dist;               // [W] scratch
  uint32_t* popped;           // [W] scratch
  uint32_t* worker_group;
  uint32_t* worker_ask;
  uint32_t* group_ask;
  uint32_t* group_off;
  uint32_t* members;          // selection order; pm_order_members puts them in BTreeSet order
  uint32_t* out_counts;       // [0] n_groups  [1] n_members  [2] n_bumped  [3] overflow flag
  uint32_t group_cap;
};

// calculate_distance, mod.rs:218-231.  Same operation order as the reference in
// IEEE f64; explicit _rn intrinsics keep nvcc from contracting a*b+c into FMA
// (rustc never does).  Only the ORDER of distances is observable.
__device__ __noinline__ double haversine_km(double lat1, double lon1, double lat2, double lon2) {
  const double kRadsPerDeg = 3.14159265358979323846264338327950288 / 180.0;
  const double lat1_rad = __dmul_rn(lat1, kRadsPerDeg);
  const double lat2_rad = __dmul_rn(lat2, kRadsPerDeg);
  const double delta_lat = __dmul_rn(__dsub_rn(lat2, lat1), kRadsPerDeg);
  const double delta_lon = __dmul_rn(__dsub_rn(lon2, lon1), kRadsPerDeg);
  const double s1 = sin(__dmul_rn(delta_lat, 0.5));
  const double s2 = sin(__dmul_rn(delta_lon, 0.5));
  double a = __dadd_rn(__dmul_rn(s1, s1),
                             __dmul_rn(__dmul_rn(cos(lat1_rad), cos(lat2_rad)), __dmul_rn(s2, s2)));
  if (a > 1.0) a = 1.0;   // antipodal rounding pushes a above 1 (NaN in the reference, no total order): clamped, DESIGN.md determinisation rule 7
  const double c = __dmul_rn(2.0, atan2(sqrt(a), sqrt(__dsub_rn(1.0, a))));
  return __dmul_rn(6371.0, c);
}

struct ProxShared {
  uint32_t u[8];
  uint32_t warp_cnt[32];
  double red_d[32];
  uint32_t red_i[32];
};

// smallest index >= start with pred(list[idx]) true, or n.  Block-uniform result.
template <class Pred>
__device__ __forceinline__ uint32_t block_find_first(ProxShared& sh, const uint32_t* list, uint32_t n,
                                                     uint32_t start, Pred pred) {
  for (uint32_t base = start; base < n; base += kProxThreads) {
    if (threadIdx.x == 0) sh.u[7] = kNone;
    __syncthreads();
    const uint32_t i = base + threadIdx.x;
    if (i < n && pred(list[i])) atomicMin(&sh.u[7], i);
    __syncthreads();
    const uint32_t f = sh.u[7];
    __syncthreads();
    if (f != kNone) return f;
  }
  return n;
}

// lexicographic (distance, position) arg-min over the entries with (e & skip) == 0 and, when
// need != 0, (e & need) != 0; kNone when there is none
__device__ __forceinline__ uint32_t block_argmin(ProxShared& sh, const uint32_t* list,
                                                 const double* dist, uint32_t n,
                                                 uint32_t skip = kTakenBit, uint32_t need = 0u) {
  double bd = 1.7976931348623157e308;
  uint32_t bi = kNone;
  for (uint32_t i = threadIdx.x; i < n; i += kProxThreads) {
    const uint32_t e = list[i];
    if ((e & skip) != 0u || (need != 0u && (e & need) == 0u)) continue;
    const double d = dist[i];
    if (bi == kNone || d < bd) { bd = d; bi = i; }  // i increases: ties keep the smaller position
  }
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
    bd = sh.red_d[lane];
    bi = sh.red_i[lane];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, bd, off);
      const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (oi != kNone && (bi == kNone || od < bd || (od == bd && oi < bi))) { bd = od; bi = oi; }
    }
    if (lane == 0) sh.u[6] = bi;
  }
  __syncthreads();
  const uint32_t r = sh.u[6];
  __syncthreads();
  return r;
}

__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t* a, uint32_t n, uint32_t key) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kProxThreads) pm_proximity_sweep(ProxParams p) {
  __shared__ ProxShared sh;
  const uint32_t T = p.ev.n_asks;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  uint32_t g = 0, mpos = 0, c_lo = 0;

  while (c_lo < T) {
    // ---- next configuration that has members
    if (tid == 0) sh.u[0] = kNone;
    __syncthreads();
    {
      const uint32_t c = c_lo + tid;
      if (c < T && (p.base_len[c] + p.xcount[c] != 0u || p.amin[c] == 0u)) atomicMin(&sh.u[0], c);
    }
    __syncthreads();
    const uint32_t c = sh.u[0];
    __syncthreads();
    if (c == kNone) { c_lo += kProxThreads; continue; }
    const uint32_t mn = p.amin[c], mx = p.amax[c];
    const uint32_t bl = p.base_len[c], xc = p.xcount[c];
    const uint32_t n = bl + xc;
    const uint32_t* base = p.order + p.seg_start[c];

    // ---- handed-down workers: gather, bitonic sort by index
    uint32_t P = 1;
    while (P < xc) P <<= 1;
    if (xc) {
      if (tid == 0) {
        uint32_t j = 0;
        for (uint32_t x = p.xhead[c]; x != kNone; x = p.xnext[x]) p.xs[j++] = x;
      }
      for (uint32_t j = xc + tid; j < P; j += kProxThreads) p.xs[j] = kNone;
      __syncthreads();
      for (uint32_t k = 2; k <= P; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
          for (uint32_t i = tid; i < P; i += kProxThreads) {
            const uint32_t l = i ^ j;
            if (l > i) {
              const uint32_t a = p.xs[i], b = p.xs[l];
              const bool up = (i & k) == 0;
              if ((a > b) == up) { p.xs[i] = b; p.xs[l] = a; }
            }
          }
          __syncthreads();
        }
    }
    // ---- merge into canonical (index) order, tagging located workers
    for (uint32_t i = tid; i < bl; i += kProxThreads) {
      const uint32_t w = base[i];
      const uint32_t pos = i + lower_bound_u32(p.xs, xc, w);
      p.list[pos] = w | ((p.ev.wa[w].w & PM_W_HAS_LOC) ? kLocBit : 0u);
    }
    for (uint32_t j = tid; j < xc; j += kProxThreads) {
      const uint32_t w = p.xs[j];
      const uint32_t pos = j + lower_bound_u32(base, bl, w);
      p.list[pos] = w | ((p.ev.wa[w].w & PM_W_HAS_LOC) ? kLocBit : 0u);
    }
    __syncthreads();

    // ---- the group loop of mod.rs:507-609 for this configuration
    uint32_t remaining = n, ploc = 0, pany = 0;
    for (;;) {
      if (remaining < mn) break;                                         // :507 / :517
      uint32_t seed_pos = n;
      bool seed_loc = false;
      if (remaining) {
        if (ploc < n) {
          seed_pos = block_find_first(sh, p.list, n, ploc,
                                      [](uint32_t e) { return (e & (kTakenBit | kLocBit)) == kLocBit; });
          ploc = seed_pos;
        }
        if (seed_pos < n) {
          seed_loc = true;
        } else {                                                         // .or(compatible_nodes.first())
          seed_pos = block_find_first(sh, p.list, n, pany, [](uint32_t e) { return (e & kTakenBit) == 0u; });
          pany = seed_pos;
        }
      }
      const bool have_seed = seed_pos < n;
      const uint32_t size = have_seed ? (mx ? min(mx, remaining) : 1u) : 0u;  // seed is inserted before the max check
      if (size < mn) break;                                              // :564
      if (g >= p.group_cap) { if (tid == 0) p.out_counts[3] = 1u; break; }
      if (tid == 0) { p.group_ask[g] = c; p.group_off[g] = mpos; }
      if (have_seed) {
        const uint32_t seed_w = p.list[seed_pos] & kIdxMask;
        __syncthreads();
        if (tid == 0) {
          p.list[seed_pos] |= kTakenBit;
          p.members[mpos] = seed_w;
          p.worker_group[seed_w] = g;
          p.worker_ask[seed_w] = c;
        }
        const uint32_t k = size - 1;
        if (k) {
          if (seed_loc) {
            const double slat = p.lat[seed_w], slon = p.lon[seed_w];
            for (uint32_t i = tid; i < n; i += kProxThreads) {
              const uint32_t e = p.list[i];
              if (e & kTakenBit) continue;
              p.dist[i] = (e & kLocBit) ? haversine_km(slat, slon, p.lat[e & kIdxMask], p.lon[e & kIdxMask])
                                        : 1.7976931348623157e308;
            }
            __syncthreads();
            for (uint32_t r = 0; r < k; ++r) {
              const uint32_t bi = block_argmin(sh, p.list, p.dist, n);
              if (tid == 0) {
                const uint32_t w = p.list[bi] & kIdxMask;
                p.list[bi] |= kTakenBit;
                p.members[mpos + 1 + r] = w;
                p.worker_group[w] = g;
                p.worker_ask[w] = c;
              }
              __syncthreads();
            }
          } else {
            // seed has no location => nobody left has one: the next k remaining, in order
            __syncthreads();
            uint32_t cnt = 0;
            for (uint32_t pos = pany; cnt < k && pos < n; pos += kProxThreads) {
              const uint32_t i = pos + tid;
              const bool live = i < n && (p.list[i] & kTakenBit) == 0u;
              const uint32_t b = __ballot_sync(0xffffffffu, live);
              if (lane == 0) sh.warp_cnt[warp] = (uint32_t)__popc(b);
              __syncthreads();
              uint32_t before = 0, total = 0;
              for (uint32_t q = 0; q < (uint32_t)kProxThreads / 32; ++q) {
                const uint32_t v = sh.warp_cnt[q];
                if (q < warp) before += v;
                total += v;
              }
              const uint32_t rank = cnt + before + (uint32_t)__popc(b & ((1u << lane) - 1u));
              if (live && rank < k) {
                const uint32_t w = p.list[i] & kIdxMask;
                p.list[i] |= kTakenBit;
                p.members[mpos + 1 + rank] = w;
                p.worker_group[w] = g;
                p.worker_ask[w] = c;
              }
              cnt += total;
              __syncthreads();
            }
          }
        }
      }
      __syncthreads();
      remaining -= size;
      mpos += size;
      ++g;
      if (size == 0) break;                                              // :606 no progress
    }

    // ---- leftovers move on to their next feasible configuration
    if (tid == 0) sh.u[1] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < n; i += kProxThreads) {
      const uint32_t e = p.list[i];
      if ((e & kTakenBit) == 0u) p.popped[atomicAdd(&sh.u[1], 1u)] = e & kIdxMask;
    }
    __syncthreads();
    const uint32_t npop = sh.u[1];
    for (uint32_t i = warp; i < npop; i += kProxThreads / 32) {
      const uint32_t w = p.popped[i];
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
    __syncthreads();
    if (tid == 0) { p.base_len[c] = 0; p.xcount[c] = 0; p.xhead[c] = kNone; }
    __syncthreads();
    c_lo = c + 1;
  }
  if (tid == 0) {
    p.out_counts[0] = g;
    p.out_counts[1] = mpos;
    if (g < p.group_cap + 1) p.group_off[g] = mpos;
  }
}


// ------------------------------------------------------------------------------------------
// try_merge_solo_groups with ProximityOptimizationPolicy enabled (mod.rs:631-873).  The caller
// passes ONLY the nodes of solo groups, in get_all_groups() order (sorted by group id), all
// marked as candidates; "configurations" and the hand-down of unmerged nodes work as in
// pm_proximity_sweep.  Batch selection follows attempt_group_merge (mod.rs:752-848):
//   seed = first remaining LOCATED node; then the nearest LOCATED nodes while size + 1 <= max;
//   if there was no seed, or size < max && size < min: (clear if size < min and) fill in list order;
//   the merge happens iff the batch has >= 2 groups (is_merge_beneficial, :863-873), else the
//   configuration is done (:704).  Task-switching policy is applied by the host.
__global__ void __launch_bounds__(kProxThreads) pm_merge_sweep(ProxParams p) {
  __shared__ ProxShared sh;
  const uint32_t T = p.ev.n_asks;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  uint32_t g = 0, mpos = 0, c_lo = 0;
  uint32_t* pickpos = p.xs;  // positions of the tentative batch (xs is free after the merge step)

  while (c_lo < T) {
    if (tid == 0) sh.u[0] = kNone;
    __syncthreads();
    {
      const uint32_t c = c_lo + tid;
      if (c < T && (p.base_len[c] + p.xcount[c] != 0u)) atomicMin(&sh.u[0], c);
    }
    __syncthreads();
    const uint32_t c = sh.u[0];
    __syncthreads();
    if (c == kNone) { c_lo += kProxThreads; continue; }
    const uint32_t mn = p.amin[c], mx = p.amax[c];
    const uint32_t bl = p.base_len[c], xc = p.xcount[c];
    const uint32_t n = bl + xc;
    const uint32_t* base = p.order + p.seg_start[c];

    uint32_t P = 1;
    while (P < xc) P <<= 1;
    if (xc) {
      if (tid == 0) {
        uint32_t j = 0;
        for (uint32_t x = p.xhead[c]; x != kNone; x = p.xnext[x]) p.xs[j++] = x;
      }
      for (uint32_t j = xc + tid; j < P; j += kProxThreads) p.xs[j] = kNone;
      __syncthreads();
      for (uint32_t k = 2; k <= P; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
          for (uint32_t i = tid; i < P; i += kProxThreads) {
            const uint32_t l = i ^ j;
            if (l > i) {
              const uint32_t a = p.xs[i], b = p.xs[l];
              const bool up = (i & k) == 0;
              if ((a > b) == up) { p.xs[i] = b; p.xs[l] = a; }
            }
          }
          __syncthreads();
        }
    }
    for (uint32_t i = tid; i < bl; i += kProxThreads) {
      const uint32_t w = base[i];
      p.list[i + lower_bound_u32(p.xs, xc, w)] = w | ((p.ev.wa[w].w & PM_W_HAS_LOC) ? kLocBit : 0u);
    }
    for (uint32_t j = tid; j < xc; j += kProxThreads) {
      const uint32_t w = p.xs[j];
      p.list[j + lower_bound_u32(base, bl, w)] = w | ((p.ev.wa[w].w & PM_W_HAS_LOC) ? kLocBit : 0u);
    }
    __syncthreads();

    uint32_t remaining = n, ploc = 0;
    if (remaining >= mn) {   // "if compatible_groups.len() < min_group_size return" (:687)
      for (;;) {
        if (remaining < mn || remaining == 0) break;                       // :694
        uint32_t size = 0;
        // --- proximity phase
        uint32_t seed_pos = n;
        if (ploc < n) {
          seed_pos = block_find_first(sh, p.list, n, ploc,
                                      [](uint32_t e) { return (e & (kTakenBit | kLocBit)) == kLocBit; });
          ploc = seed_pos;
        }
        if (seed_pos < n) {
          const uint32_t seed_w = p.list[seed_pos] & kIdxMask;
          __syncthreads();
          if (tid == 0) { p.list[seed_pos] |= kTentBit; pickpos[0] = seed_pos; }
          size = 1;
          if (mx > 1) {
            const double slat = p.lat[seed_w], slon = p.lon[seed_w];
            for (uint32_t i = tid; i < n; i += kProxThreads) {
              const uint32_t e = p.list[i];
              if ((e & kTakenBit) || !(e & kLocBit)) continue;
              p.dist[i] = haversine_km(slat, slon, p.lat[e & kIdxMask], p.lon[e & kIdxMask]);
            }
            __syncthreads();
            while (size < mx) {                                           // total + 1 <= max
              const uint32_t bi = block_argmin(sh, p.list, p.dist, n, kTakenBit | kTentBit, kLocBit);
              if (bi == kNone) break;
              if (tid == 0) { p.list[bi] |= kTentBit; pickpos[size] = bi; }
              ++size;
              __syncthreads();
            }
          }
          __syncthreads();
        }
        // --- fallback to list order (:823-848)
        if (size == 0 || (size < mx && size < mn)) {
          if (size < mn) {
            for (uint32_t j = tid; j < size; j += kProxThreads) p.list[pickpos[j]] &= ~kTentBit;
            __syncthreads();
            size = 0;
          }
          uint32_t pos = block_find_first(sh, p.list, n, 0, [](uint32_t e) { return (e & kTakenBit) == 0u; });
          while (size < mx && pos < n) {
            const uint32_t i = pos + tid;
            const bool live = i < n && (p.list[i] & (kTakenBit | kTentBit)) == 0u;
            const uint32_t b = __ballot_sync(0xffffffffu, live);
            if (lane == 0) sh.warp_cnt[warp] = (uint32_t)__popc(b);
            __syncthreads();
            uint32_t before = 0, total = 0;
            for (uint32_t q = 0; q < (uint32_t)kProxThreads / 32; ++q) {
              const uint32_t v = sh.warp_cnt[q];
              if (q < warp) before += v;
              total += v;
            }
            const uint32_t rank = size + before + (uint32_t)__popc(b & ((1u << lane) - 1u));
            if (live && rank < mx) { p.list[i] |= kTentBit; pickpos[rank] = i; }
            size = min(mx, size + total);
            pos += kProxThreads;
            __syncthreads();
          }
        }
        // --- is_merge_beneficial: at least two groups
        if (size < 2) {
          for (uint32_t j = tid; j < size; j += kProxThreads) p.list[pickpos[j]] &= ~kTentBit;
          __syncthreads();
          break;                                                           // :704
        }
        if (g >= p.group_cap) { if (tid == 0) p.out_counts[3] = 1u; break; }
        for (uint32_t j = tid; j < size; j += kProxThreads) {
          const uint32_t pos = pickpos[j];
          const uint32_t w = p.list[pos] & kIdxMask;
          p.list[pos] = (p.list[pos] & ~kTentBit) | kTakenBit;
          p.members[mpos + j] = w;
          p.worker_group[w] = g;
          p.worker_ask[w] = c;
        }
        if (tid == 0) { p.group_ask[g] = c; p.group_off[g] = mpos; }
        __syncthreads();
        remaining -= size;
        mpos += size;
        ++g;
      }
    }

    // unmerged nodes stay solo and are considered by the following configurations
    if (tid == 0) sh.u[1] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < n; i += kProxThreads) {
      const uint32_t e = p.list[i];
      if ((e & kTakenBit) == 0u) p.popped[atomicAdd(&sh.u[1], 1u)] = e & kIdxMask;
    }
    __syncthreads();
    const uint32_t npop = sh.u[1];
    for (uint32_t i = warp; i < npop; i += kProxThreads / 32) {
      const uint32_t w = p.popped[i];
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
      }
    }
    __syncthreads();
    if (tid == 0) { p.base_len[c] = 0; p.xcount[c] = 0; p.xhead[c] = kNone; }
    __syncthreads();
    c_lo = c + 1;
  }
  if (tid == 0) {
    p.out_counts[0] = g;
    p.out_counts[1] = mpos;
    if (g < p.group_cap + 1) p.group_off[g] = mpos;
  }
}

}  // namespace pm
