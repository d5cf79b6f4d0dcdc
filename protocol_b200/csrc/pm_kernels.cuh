// pm_kernels.cuh — sm_100a kernels of the matching pass.
//
//   pm_build_cost   cost-matrix build: int64 cost[t][w] for a tile of asks x a
//                   stripe of workers, HBM-write-bound (8 B per evaluation)
//   pm_argmin       one read of the tile (8 B per evaluation): per-ask packed
//                   (min_cost, argmin) + feasible count, per-worker first ask
//   pm_fused_eval   the same outputs without materialising the matrix
//   pm_hist / pm_check_tails / pm_sweep / pm_emit_* : resolution sweep that turns
//                   "first feasible ask per worker" into the reference's groups
//
// Reference semantics being reproduced: NodeGroupsPlugin::try_form_new_groups,
// crates/orchestrator/src/plugins/node_groups/mod.rs:478-628.
#pragma once
#include "pm_device.cuh"

namespace pm {

struct EvalParams {
  const uint4* __restrict__ wa;       // plane A [n_workers]
  const uint4* __restrict__ wb;       // plane B [n_workers]
  const DevAsk* __restrict__ asks;    // [n_asks] priority order
  const DevOpt* __restrict__ opts;    // CSR, ask order
  const uint32_t* __restrict__ bits;  // [(n_patterns+1) * words], row 0 all-ones
  uint32_t words;
  uint32_t n_workers;                 // global
  uint32_t n_asks;
  uint32_t n_opts;
};

// ------------------------------------------------------------------ build
constexpr int kBuildThreads = 256;
constexpr int kBuildWPT = 2;                          // workers per thread -> one 16 B store per row
constexpr int kBuildCols = kBuildThreads * kBuildWPT; // 512 workers per CTA stripe
constexpr int kBuildRows = 128;                       // asks staged in smem per CTA
constexpr int kBuildOptCap = 256;                     // options staged per CTA

struct AskStage {
  DevAsk ask[kBuildRows];
  DevOpt opt[kBuildOptCap];
  uint64_t bar;
};

// Stage rows [r0, r1) of the ask table (and their option range) into shared
// memory with two 1-D TMA bulk copies signalled on one mbarrier.  Returns the
// option base to subtract from DevAsk::opt_off, or kNone when the option range
// does not fit (then options are read from global memory through L1).
__device__ __forceinline__ uint32_t stage_asks(AskStage& s, const EvalParams& p, uint32_t r0,
                                               uint32_t r1) {
  const uint32_t o0 = p.asks[r0].opt_off;
  const DevAsk last = p.asks[r1 - 1];
  const uint32_t o1 = last.opt_off + last.n_opts;
  const bool fits = (o1 - o0) <= (uint32_t)kBuildOptCap;
  if (threadIdx.x == 0) {
    mbar_init(&s.bar, 1);
    uint32_t bytes = (r1 - r0) * (uint32_t)sizeof(DevAsk);
    uint32_t obytes = (fits && o1 > o0) ? (o1 - o0) * (uint32_t)sizeof(DevOpt) : 0u;
    mbar_expect_tx(&s.bar, bytes + obytes);
    bulk_g2s(s.ask, p.asks + r0, bytes, &s.bar);
    if (obytes) bulk_g2s(s.opt, p.opts + o0, obytes, &s.bar);
  }
  __syncthreads();          // barrier init visible to all waiters
  mbar_wait(&s.bar, 0);
  return fits ? o0 : kNone;
}

// grid = (ceil(ld / 512), ceil(nt / 128)); each thread owns 2 adjacent workers in
// registers and walks the staged asks, emitting one 128-bit streaming store per
// row: a warp writes 512 contiguous bytes, the CTA 4 KB contiguous per row.
__global__ void __launch_bounds__(kBuildThreads)
pm_build_cost(EvalParams p, uint32_t t0, uint32_t nt, uint32_t w0, uint32_t nw,
              long long* __restrict__ cost, size_t ld) {
  __shared__ __align__(128) AskStage s;
  const uint32_t r0 = blockIdx.y * kBuildRows;
  const uint32_t r1 = min(nt, r0 + (uint32_t)kBuildRows);
  const uint32_t obase = stage_asks(s, p, t0 + r0, t0 + r1);
  const DevOpt* optp = (obase == kNone) ? p.opts : (s.opt - obase);

  const uint32_t c = blockIdx.x * kBuildCols + threadIdx.x * kBuildWPT;  // column in tile
  if (c >= ld) return;
  WorkerReg w[kBuildWPT];
  long long feas_cost[kBuildWPT];
#pragma unroll
  for (int k = 0; k < kBuildWPT; ++k) {
    const uint32_t col = c + k;
    if (col < nw) {
      const uint32_t gw = w0 + col;
      w[k] = make_worker(__ldg(p.wa + gw), __ldg(p.wb + gw));
      feas_cost[k] = ((long long)w[k].price << 32) | (long long)gw;
    } else {
      w[k] = null_worker();
      feas_cost[k] = kInf;
    }
  }
  longlong2* out = reinterpret_cast<longlong2*>(cost + (size_t)r0 * ld + c);
  const size_t ld2 = ld / 2;
#pragma unroll 4
  for (uint32_t r = 0; r < r1 - r0; ++r) {
    const DevAsk a = s.ask[r];
    longlong2 v;
    v.x = (w[0].candidate && ask_meets(a, optp, w[0], p.bits, p.words)) ? feas_cost[0] : kInf;
    v.y = (w[1].candidate && ask_meets(a, optp, w[1], p.bits, p.words)) ? feas_cost[1] : kInf;
    __stcs(out + (size_t)r * ld2, v);
  }
}

// ------------------------------------------------------------------ argmin
constexpr int kArgThreads = 256;
constexpr int kArgStripes = 8;                   // 16 B loads in flight per lane per row
constexpr int kArgCols = kArgStripes * 64;       // 512 columns per CTA
constexpr int kArgRows = 64;                     // rows per CTA (8 per warp)

// grid = (ceil(ld / 512), ceil(nt / 64)).  Warp j reduces rows j, j+8, ...; for a
// row it issues 8 independent 128-bit streaming loads (4 KB per warp in flight),
// folds them lane-locally, then one shuffle reduction per row.
__global__ void __launch_bounds__(kArgThreads)
pm_argmin(const long long* __restrict__ cost, size_t ld, uint32_t nt, uint32_t t0, uint32_t w0,
          uint32_t nw, uint32_t* __restrict__ first_ask, long long* __restrict__ ask_best,
          uint32_t* __restrict__ ask_count) {
  __shared__ uint32_t s_cm[kArgCols];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t c0 = blockIdx.x * kArgCols;
  const uint32_t r0 = blockIdx.y * kArgRows;
  for (uint32_t i = threadIdx.x; i < (uint32_t)kArgCols; i += kArgThreads) s_cm[i] = kNone;
  __syncthreads();

  uint32_t cm[kArgStripes][2];
#pragma unroll
  for (int s = 0; s < kArgStripes; ++s) cm[s][0] = cm[s][1] = kNone;

  for (uint32_t rr = warp; rr < (uint32_t)kArgRows; rr += kArgThreads / 32) {
    const uint32_t r = r0 + rr;
    if (r >= nt) break;
    const longlong2* rowp = reinterpret_cast<const longlong2*>(cost + (size_t)r * ld);
    longlong2 v[kArgStripes];
#pragma unroll
    for (int s = 0; s < kArgStripes; ++s) {
      const uint32_t col = c0 + s * 64 + lane * 2;
      if (col < ld) v[s] = __ldcs(rowp + (col >> 1));
      else v[s] = make_longlong2(kInf, kInf);
    }
    const uint32_t t = t0 + r;
    long long best = kInf;
    uint32_t cnt = 0;
#pragma unroll
    for (int s = 0; s < kArgStripes; ++s) {
      best = min(best, min(v[s].x, v[s].y));
      const bool fx = v[s].x != kInf, fy = v[s].y != kInf;
      cnt += (uint32_t)fx + (uint32_t)fy;
      cm[s][0] = fx ? min(cm[s][0], t) : cm[s][0];
      cm[s][1] = fy ? min(cm[s][1], t) : cm[s][1];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      best = min(best, __shfl_xor_sync(0xffffffffu, best, off));
      cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    }
    if (lane == 0 && cnt) {
      atomicMin(ask_best + t, best);
      atomicAdd(ask_count + t, cnt);
    }
  }
#pragma unroll
  for (int s = 0; s < kArgStripes; ++s) {
    const uint32_t i = s * 64 + lane * 2;
    if (cm[s][0] != kNone) atomicMin(&s_cm[i], cm[s][0]);
    if (cm[s][1] != kNone) atomicMin(&s_cm[i + 1], cm[s][1]);
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < (uint32_t)kArgCols; i += kArgThreads) {
    const uint32_t col = c0 + i;
    if (col < nw && s_cm[i] != kNone) atomicMin(first_ask + w0 + col, s_cm[i]);
  }
}

// ------------------------------------------------------------------ fused
// Same stripe/row decomposition as pm_build_cost, but the int64 cost never
// leaves the SM: per-row results go through a ballot (first-fit cost is the
// worker index, so the row minimum is the lowest set bit), per-worker results
// stay in registers.  Integer-issue-bound, not HBM-bound.
constexpr int kFusedRows = 128;
__global__ void __launch_bounds__(kBuildThreads)
pm_fused_eval(EvalParams p, uint32_t t0, uint32_t nt, uint32_t w0, uint32_t nw,
              uint32_t* __restrict__ first_ask, long long* __restrict__ ask_best,
              uint32_t* __restrict__ ask_count) {
  __shared__ __align__(128) AskStage s;
  __shared__ uint32_t s_cnt[kFusedRows];
  __shared__ unsigned long long s_best[kFusedRows];
  const uint32_t r0 = blockIdx.y * kFusedRows;
  const uint32_t r1 = min(nt, r0 + (uint32_t)kFusedRows);
  const uint32_t obase = stage_asks(s, p, t0 + r0, t0 + r1);
  const DevOpt* optp = (obase == kNone) ? p.opts : (s.opt - obase);
  for (uint32_t i = threadIdx.x; i < (uint32_t)kFusedRows; i += kBuildThreads) {
    s_cnt[i] = 0;
    s_best[i] = (unsigned long long)kInf;
  }
  __syncthreads();

  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t c = blockIdx.x * kBuildCols + threadIdx.x * kBuildWPT;
  WorkerReg w[kBuildWPT];
#pragma unroll
  for (int k = 0; k < kBuildWPT; ++k) {
    const uint32_t col = c + k;
    if (col < nw) {
      const uint32_t gw = w0 + col;
      w[k] = make_worker(__ldg(p.wa + gw), __ldg(p.wb + gw));
    } else {
      w[k] = null_worker();
    }
  }
  uint32_t first[kBuildWPT] = {kNone, kNone};
  const bool uniform_price = true;  // reference modes: cost = worker index
  (void)uniform_price;
#pragma unroll 2
  for (uint32_t r = 0; r < r1 - r0; ++r) {
    const DevAsk a = s.ask[r];
    const bool f0 = w[0].candidate && ask_meets(a, optp, w[0], p.bits, p.words);
    const bool f1 = w[1].candidate && ask_meets(a, optp, w[1], p.bits, p.words);
    const uint32_t t = t0 + r0 + r;
    first[0] = (f0 && first[0] == kNone) ? t : first[0];
    first[1] = (f1 && first[1] == kNone) ? t : first[1];
    const uint32_t b0 = __ballot_sync(0xffffffffu, f0), b1 = __ballot_sync(0xffffffffu, f1);
    if ((b0 | b1) != 0u && lane == 0) {
      // lowest worker index among the warp's 64 columns
      const uint32_t l0 = b0 ? (uint32_t)__ffs(b0) - 1u : 64u, l1 = b1 ? (uint32_t)__ffs(b1) - 1u : 64u;
      const uint32_t rel = min(l0 * 2u, l1 * 2u + 1u);
      const uint32_t gw = w0 + (c - lane * kBuildWPT) + rel;
      atomicAdd(&s_cnt[r], (uint32_t)__popc(b0) + (uint32_t)__popc(b1));
      atomicMin(&s_best[r], (unsigned long long)gw);
    }
  }
#pragma unroll
  for (int k = 0; k < kBuildWPT; ++k)
    if (first[k] != kNone) atomicMin(first_ask + w0 + c + k, first[k]);
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < r1 - r0; i += kBuildThreads) {
    if (s_cnt[i]) {
      atomicAdd(ask_count + t0 + r0 + i, s_cnt[i]);
      atomicMin(ask_best + t0 + r0 + i, (long long)s_best[i]);
    }
  }
}

// ------------------------------------------------------------------ small utilities
__global__ void pm_fill_u32(uint32_t* p, uint32_t v, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void pm_fill_i64(long long* p, long long v, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void pm_iota_u32(uint32_t* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = (uint32_t)i;
}
__global__ void pm_scatter_flags(uint4* wa, const uint32_t* idx, const uint32_t* flags, uint32_t n, uint32_t n_workers) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && idx[i] < n_workers) wa[idx[i]].w = flags[i];
}

// ------------------------------------------------------------------ resolution
// Sort key of a worker: its (first feasible) ask; in proximity mode with solo
// groups the key also orders located workers first (seed rule, mod.rs:526-530).
__global__ void pm_make_keys(const uint32_t* __restrict__ cur, const uint4* __restrict__ wa,
                             uint32_t n_workers, uint32_t shift, uint32_t* __restrict__ keys,
                             uint32_t* __restrict__ hist) {
  uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_workers) return;
  uint32_t c = cur[w], k = kNone;
  if (c != kNone) {
    k = c << shift;
    if (shift) k |= (wa[w].w & PM_W_HAS_LOC) ? 0u : 1u;
    atomicAdd(hist + k, 1u);
  }
  keys[w] = k;
}

// any configuration whose member count leaves an under-filled tail
// (0 < K mod max < min: the tail group is not formed, mod.rs:517,564)?
__global__ void pm_check_tails(const uint32_t* __restrict__ hist, const uint32_t* __restrict__ amin,
                               const uint32_t* __restrict__ amax, uint32_t n_asks,
                               uint32_t* __restrict__ any_bad) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_asks) return;
  const uint32_t mx = amax[c];
  if (mx == 0) return;
  const uint32_t r = hist[c] % mx;
  if (r != 0 && r < amin[c]) *any_bad = 1u;
}

struct SweepParams {
  EvalParams ev;
  uint32_t* cur;             // [W] current ask of each worker
  uint32_t* base_len;        // [T] members still in the sorted base segment
  const uint32_t* seg_start; // [T]
  const uint32_t* order;     // [W] workers sorted by (ask, index)
  uint32_t* xhead;           // [T] list of workers bumped into the ask
  uint32_t* xnext;           // [W]
  uint32_t* xcount;          // [T]
  const uint32_t* amin;
  const uint32_t* amax;
  uint32_t* popped;          // [W] scratch
  uint32_t* n_bumped;
};

// Sequential-in-priority sweep (single CTA): configurations are final in
// increasing priority index; the first one with an under-filled tail hands its
// last `K mod max` members (largest canonical index — groups are cut from the
// front, mod.rs:554-561) to their next feasible configuration.
__global__ void __launch_bounds__(1024) pm_sweep(SweepParams p) {
  __shared__ uint32_t s_first_bad, s_npop;
  const uint32_t T = p.ev.n_asks;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  uint32_t c_lo = 0;
  while (c_lo < T) {
    if (tid == 0) s_first_bad = kNone;
    __syncthreads();
    const uint32_t c = c_lo + tid;
    if (c < T) {
      const uint32_t mx = p.amax[c];
      if (mx) {
        const uint32_t r = (p.base_len[c] + p.xcount[c]) % mx;
        if (r && r < p.amin[c]) atomicMin(&s_first_bad, c);
      }
    }
    __syncthreads();
    const uint32_t cb = s_first_bad;
    if (cb == kNone) {
      c_lo += 1024;
      continue;
    }
    if (tid == 0) {
      uint32_t bl = p.base_len[cb], xc = p.xcount[cb];
      const uint32_t r = (bl + xc) % p.amax[cb];
      for (uint32_t i = 0; i < r; ++i) {
        uint32_t bx = kNone, bprev = kNone, prev = kNone;  // largest index among bumped-in
        for (uint32_t x = p.xhead[cb]; x != kNone; prev = x, x = p.xnext[x])
          if (bx == kNone || x > bx) { bx = x; bprev = prev; }
        const uint32_t bw = bl ? p.order[p.seg_start[cb] + bl - 1] : kNone;
        uint32_t take;
        if (bx != kNone && (bw == kNone || bx > bw)) {
          take = bx;
          if (bprev == kNone) p.xhead[cb] = p.xnext[bx];
          else p.xnext[bprev] = p.xnext[bx];
          --xc;
        } else {
          take = bw;
          --bl;
        }
        p.popped[i] = take;
      }
      p.base_len[cb] = bl;
      p.xcount[cb] = xc;
      s_npop = r;
    }
    __syncthreads();
    const uint32_t npop = s_npop;
    for (uint32_t i = warp; i < npop; i += 32) {
      const uint32_t w = p.popped[i];
      const WorkerReg wr = make_worker(p.ev.wa[w], p.ev.wb[w]);
      uint32_t found = kNone;
      for (uint32_t cbase = cb + 1; cbase < T; cbase += 32) {
        const uint32_t c2 = cbase + lane;
        bool ok = false;
        if (c2 < T) ok = ask_meets(p.ev.asks[c2], p.ev.opts, wr, p.ev.bits, p.ev.words);
        const uint32_t b = __ballot_sync(0xffffffffu, ok);
        if (b) {
          found = cbase + (uint32_t)__ffs(b) - 1u;
          break;
        }
      }
      if (lane == 0) {
        p.cur[w] = found;
        if (found != kNone) {
          const uint32_t old = atomicExch(&p.xhead[found], w);
          p.xnext[w] = old;
          atomicAdd(&p.xcount[found], 1u);
        }
        atomicAdd(p.n_bumped, 1u);
      }
    }
    __syncthreads();
    c_lo = cb + 1;
  }
}

// groups per bin (bin = ask << shift | located-first bit): full chunks of `max`
// plus the tail when it reaches `min` (mod.rs:545-566), plus the empty group a
// min_group_size == 0 configuration produces on its last iteration (:564,:606).
__global__ void pm_count_groups(const uint32_t* __restrict__ hist, const uint32_t* __restrict__ amin,
                                const uint32_t* __restrict__ amax, uint32_t n_bins, uint32_t shift,
                                uint32_t* __restrict__ ngroups) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_bins) return;
  const uint32_t c = b >> shift, mx = amax[c], mn = amin[c], K = hist[b];
  uint32_t g = 0;
  if (mx) {
    g = K / mx;
    const uint32_t r = K % mx;
    if (r && r >= mn) ++g;
  }
  if (mn == 0 && (!shift || (b & 1u))) ++g;
  ngroups[b] = g;
}

// one thread per sorted position: worker -> (group, ask); group heads write the
// group table.  Positions past the formed groups of a bin (an unformed tail can
// only exist if the sweep was skipped by the caller) stay unassigned.
__global__ void pm_emit_workers(const uint32_t* __restrict__ keys_sorted,
                                const uint32_t* __restrict__ order, uint32_t n_assigned,
                                const uint32_t* __restrict__ hist,
                                const uint32_t* __restrict__ seg_start,
                                const uint32_t* __restrict__ group_base,
                                const uint32_t* __restrict__ amin, const uint32_t* __restrict__ amax,
                                uint32_t shift, uint32_t* __restrict__ worker_group,
                                uint32_t* __restrict__ worker_ask, uint32_t* __restrict__ group_ask,
                                uint32_t* __restrict__ group_off) {
  uint32_t pidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (pidx >= n_assigned) return;
  const uint32_t b = keys_sorted[pidx], c = b >> shift, w = order[pidx];
  const uint32_t mx = amax[c], K = hist[b], r = pidx - seg_start[b];
  if (mx == 0) return;
  const uint32_t full = K / mx, rem = K % mx;
  if (r >= full * mx && rem < amin[c]) return;  // unformed tail
  const uint32_t g = group_base[b] + r / mx;
  worker_group[w] = g;
  worker_ask[w] = c;
  if (r % mx == 0) {
    group_ask[g] = c;
    group_off[g] = pidx;
  }
}

// the trailing empty group of min_group_size == 0 configurations
__global__ void pm_emit_empty_groups(const uint32_t* __restrict__ hist,
                                     const uint32_t* __restrict__ seg_start,
                                     const uint32_t* __restrict__ group_base,
                                     const uint32_t* __restrict__ ngroups,
                                     const uint32_t* __restrict__ amin, uint32_t n_bins,
                                     uint32_t shift, uint32_t* __restrict__ group_ask,
                                     uint32_t* __restrict__ group_off) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_bins) return;
  const uint32_t c = b >> shift;
  if (amin[c] != 0 || (shift && !(b & 1u))) return;
  const uint32_t g = group_base[b] + ngroups[b] - 1;
  group_ask[g] = c;
  group_off[g] = seg_start[b] + hist[b];
}

// members of a group in BTreeSet<String> order (mod.rs:63-69): position = number
// of members whose address rank is smaller.
__global__ void pm_order_members(const uint32_t* __restrict__ order, uint32_t n_assigned,
                                 const uint32_t* __restrict__ worker_group,
                                 const uint32_t* __restrict__ group_off,
                                 const uint32_t* __restrict__ addr_rank,
                                 uint32_t* __restrict__ members) {
  uint32_t pidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (pidx >= n_assigned) return;
  const uint32_t w = order[pidx];
  const uint32_t g = worker_group[w];
  if (g == kNone) { members[pidx] = w; return; }
  const uint32_t gs = group_off[g], ge = group_off[g + 1];
  uint32_t pos = 0;
  if (addr_rank) {
    const uint32_t mine = addr_rank[w];
    for (uint32_t q = gs; q < ge; ++q) pos += (addr_rank[order[q]] < mine) ? 1u : 0u;
  } else {
    pos = pidx - gs;
  }
  members[gs + pos] = w;
}

}  // namespace pm
