// pm_proximity_band.cuh — EXPERIMENTAL, OFF BY DEFAULT (PM_TUNE_PROX=1 selects it; not yet run on a GPU).
//
// pm_proximity_sweep (pm_proximity.cuh) restates the reference's proximity loop literally: for every group it
// computes the distance from the seed to EVERY remaining compatible worker and takes max-1 arg-mins over all of
// them — quadratic in the candidates of a configuration (0.39 s at 100k workers, 40 s at 1M workers / 125k pair
// groups, profiles/r01_host_paths.txt).  This variant forms the same groups from a latitude-ordered view:
//
//   haversine distance >= R * |delta latitude|   (a >= sin^2(dlat / 2) in calculate_distance, mod.rs:218-231)
//
// so, with a configuration's located candidates sorted by (latitude, list position), a group gathers a window of
// ranks around its seed, takes its k nearest inside the window, and is done as soon as the latitude band the window
// covers completely has a lower bound above the k-th distance found; otherwise the window grows fourfold.  Ties
// (equal distance) go to the smaller list position exactly as the stable sort of the reference does, unlocated
// workers (distance f64::MAX) follow in list order once the located ones are used up.  Taken workers stay in the
// latitude order as tombstones and are squeezed out whenever they outnumber the live ones.
//
// The same pruning is implemented and TESTED in the test suite's CPU checker (its proximity mode 2,
// tests/test_oracle_groups.py::test_latitude_pruned_proximity_equals_the_restated_loop), where it turns 7.9 s into
// 0.16 s at 100k workers x 2000 configurations.  Everything outside the located-seed branch is pm_proximity_sweep's
// code unchanged (configuration walk, canonical list, unlocated seeds, hand-down of leftovers).
#pragma once
#include "pm_proximity.cuh"

namespace pm {

struct BandParams {
  ProxParams p;
  double* lat_key;     // [pow2 >= W] latitude of the entry at each rank (padding: DBL_MAX)
  uint32_t* lat_ord;   // [pow2 >= W] list position of the entry at each rank (padding: kNone)
  uint32_t* rank_of;   // [W] list position -> rank
};

struct BandShared {
  ProxShared base;
  uint32_t red_p[32];
  uint32_t cnt;        // window fill / member slots
};

constexpr double kBandMax = 1.7976931348623157e308;

// (distance, list position) arg-min over the window slots that are not marked; returns the slot or kNone
__device__ __forceinline__ uint32_t block_argmin_window(BandShared& sh, const uint32_t* win_pos, const double* win_dist, uint32_t m) {
  double bd = kBandMax;
  uint32_t bp = kNone, bs = kNone;
  for (uint32_t s = threadIdx.x; s < m; s += kProxThreads) {
    const uint32_t e = win_pos[s];
    if (e & kTakenBit) continue;
    const double d = win_dist[s];
    if (bs == kNone || d < bd || (d == bd && e < bp)) { bd = d; bp = e; bs = s; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, bd, off);
    const uint32_t op = __shfl_xor_sync(0xffffffffu, bp, off);
    const uint32_t os = __shfl_xor_sync(0xffffffffu, bs, off);
    if (os != kNone && (bs == kNone || od < bd || (od == bd && op < bp))) { bd = od; bp = op; bs = os; }
  }
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (lane == 0) { sh.base.red_d[warp] = bd; sh.red_p[warp] = bp; sh.base.red_i[warp] = bs; }
  __syncthreads();
  if (warp == 0) {
    bd = sh.base.red_d[lane];
    bp = sh.red_p[lane];
    bs = sh.base.red_i[lane];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, bd, off);
      const uint32_t op = __shfl_xor_sync(0xffffffffu, bp, off);
      const uint32_t os = __shfl_xor_sync(0xffffffffu, bs, off);
      if (os != kNone && (bs == kNone || od < bd || (od == bd && op < bp))) { bd = od; bp = op; bs = os; }
    }
    if (lane == 0) sh.base.u[6] = bs;
  }
  __syncthreads();
  const uint32_t r = sh.base.u[6];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kProxThreads) pm_proximity_sweep_banded(BandParams bp_) {
  __shared__ BandShared shb;
  ProxShared& sh = shb.base;
  const ProxParams& p = bp_.p;
  double* const lat_key = bp_.lat_key;
  uint32_t* const lat_ord = bp_.lat_ord;
  uint32_t* const rank_of = bp_.rank_of;
  uint32_t* const win_pos = p.popped;   // free until the leftovers phase of the configuration
  double* const win_dist = p.dist;
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

    // ---- latitude order of the located entries (only configurations that can form groups of more than one)
    uint32_t nloc = 0, dead_loc = 0;
    const bool banded = mx > 1u;
    if (banded) {
      if (tid == 0) shb.cnt = 0;
      __syncthreads();
      for (uint32_t i = tid; i < n; i += kProxThreads) {
        const uint32_t e = p.list[i];
        if (e & kLocBit) {
          const uint32_t s = atomicAdd(&shb.cnt, 1u);
          lat_key[s] = p.lat[e & kIdxMask];
          lat_ord[s] = i;
        }
      }
      __syncthreads();
      nloc = shb.cnt;
      uint32_t P2 = 1;
      while (P2 < nloc) P2 <<= 1;
      for (uint32_t j = nloc + tid; j < P2; j += kProxThreads) { lat_key[j] = kBandMax; lat_ord[j] = kNone; }
      __syncthreads();
      for (uint32_t k = 2; k <= P2; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
          for (uint32_t i = tid; i < P2; i += kProxThreads) {
            const uint32_t l = i ^ j;
            if (l > i) {
              const double ka = lat_key[i], kb = lat_key[l];
              const uint32_t oa = lat_ord[i], ob = lat_ord[l];
              const bool up = (i & k) == 0;
              const bool gt = ka > kb || (ka == kb && oa > ob);
              if (gt == up) { lat_key[i] = kb; lat_key[l] = ka; lat_ord[i] = ob; lat_ord[l] = oa; }
            }
          }
          __syncthreads();
        }
      for (uint32_t j = tid; j < nloc; j += kProxThreads) rank_of[lat_ord[j]] = j;
      __syncthreads();
    }

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
        __syncthreads();
        const uint32_t k = size - 1;
        uint32_t n_sel = 0;        // located members chosen by distance
        if (k && seed_loc) {
          // (banded is true here: k >= 1 means mx > 1)
          const double slat = p.lat[seed_w], slon = p.lon[seed_w];
          const uint32_t rs = rank_of[seed_pos];
          uint32_t half = 512u;
          for (;;) {
            const uint32_t lo = rs > half ? rs - half : 0u;
            const uint32_t hi = min(nloc - 1u, rs + half);
            if (tid == 0) shb.cnt = 0;
            __syncthreads();
            for (uint32_t j = lo + tid; j <= hi; j += kProxThreads) {
              const uint32_t i = lat_ord[j];
              const uint32_t e = p.list[i];
              if ((e & kTakenBit) == 0u) {
                const uint32_t s = atomicAdd(&shb.cnt, 1u);
                win_pos[s] = i;
                win_dist[s] = haversine_km(slat, slon, p.lat[e & kIdxMask], p.lon[e & kIdxMask]);
              }
            }
            __syncthreads();
            const uint32_t m = shb.cnt;
            __syncthreads();   // everybody has read the count before the next window resets it
            const bool all_in = lo == 0u && hi == nloc - 1u;
            if (m < k && !all_in) { half = half > (1u << 28) ? half : half * 4u; continue; }
            // the latitude band that lies completely inside the window
            double band = kBandMax;
            if (lo != 0u) band = fmin(band, slat - lat_key[lo]);
            if (hi != nloc - 1u) band = fmin(band, lat_key[hi] - slat);
            const uint32_t sel = min(k, m);
            double d_last = 0.0;
            for (uint32_t r = 0; r < sel; ++r) {
              const uint32_t s = block_argmin_window(shb, win_pos, win_dist, m);
              d_last = win_dist[s];
              if (tid == 0) win_pos[s] |= kTakenBit;
              __syncthreads();
            }
            bool accept = all_in;
            if (!accept && sel == k) {
              const double lb = 6371.0 * (band * (3.14159265358979323846264338327950288 / 180.0)) * (1.0 - 1e-9) - 1e-9;
              accept = lb > d_last;
            }
            if (!accept) { half = half > (1u << 28) ? half : half * 4u; continue; }   // the marks die with the window
            // commit the marked slots
            if (tid == 0) shb.cnt = 0;
            __syncthreads();
            for (uint32_t s = tid; s < m; s += kProxThreads) {
              const uint32_t e = win_pos[s];
              if (e & kTakenBit) {
                const uint32_t i = e & kIdxMask;
                const uint32_t w = p.list[i] & kIdxMask;
                p.list[i] |= kTakenBit;
                const uint32_t slot = atomicAdd(&shb.cnt, 1u);
                p.members[mpos + 1 + slot] = w;
                p.worker_group[w] = g;
                p.worker_ask[w] = c;
              }
            }
            __syncthreads();
            n_sel = sel;
            break;
          }
          dead_loc += 1u + n_sel;
        } else if (seed_loc) {
          dead_loc += 1u;
        }
        if (k > n_sel) {
          // no located worker left (or the seed has no location): the next ones in list order
          const uint32_t kk = k - n_sel;
          __syncthreads();
          uint32_t cnt = 0;
          for (uint32_t pos = pany; cnt < kk && pos < n; pos += kProxThreads) {
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
            if (live && rank < kk) {
              const uint32_t w = p.list[i] & kIdxMask;
              p.list[i] |= kTakenBit;
              p.members[mpos + 1 + n_sel + rank] = w;
              p.worker_group[w] = g;
              p.worker_ask[w] = c;
            }
            cnt += total;
            __syncthreads();
          }
        }
      }
      __syncthreads();
      remaining -= size;
      mpos += size;
      ++g;
      if (size == 0) break;                                              // :606 no progress

      // ---- squeeze the tombstones out of the latitude order once they outnumber the live entries
      if (banded && nloc > 2048u && dead_loc * 2u > nloc) {
        uint32_t out = 0;
        for (uint32_t b0 = 0; b0 < nloc; b0 += kProxThreads) {
          const uint32_t j = b0 + tid;
          double key = 0.0;
          uint32_t ord = kNone;
          bool keep = false;
          if (j < nloc) {
            key = lat_key[j];
            ord = lat_ord[j];
            keep = (p.list[ord] & kTakenBit) == 0u;
          }
          const uint32_t bal = __ballot_sync(0xffffffffu, keep);
          if (lane == 0) sh.warp_cnt[warp] = (uint32_t)__popc(bal);
          __syncthreads();   // every read of this chunk is done; targets are <= the slots just read
          uint32_t before = 0, total = 0;
          for (uint32_t q = 0; q < (uint32_t)kProxThreads / 32; ++q) {
            const uint32_t v = sh.warp_cnt[q];
            if (q < warp) before += v;
            total += v;
          }
          if (keep) {
            const uint32_t dst = out + before + (uint32_t)__popc(bal & ((1u << lane) - 1u));
            lat_key[dst] = key;
            lat_ord[dst] = ord;
            rank_of[ord] = dst;
          }
          out += total;
          __syncthreads();
        }
        nloc = out;
        dead_loc = 0;
      }
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

}  // namespace pm
