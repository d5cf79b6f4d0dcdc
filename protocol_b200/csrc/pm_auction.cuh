// pm_auction.cuh — NORTH-STAR EXTENSION: price-capped forward auction (one worker per ask).
//
// The reference orchestrator has no prices, caps, reputation or auction:
// nothing here restates reference code, and parity for this mode is against the
// builder's own sequential restatement (the auction checker of the test oracle) —
// "self-oracle, parity unpinned by the reference".  The mode is inert unless
// PM_MODE_AUCTION is requested; the reference modes never read ext_ask_price.
//
// Synchronous (Jacobi) Bertsekas auction:
//   feasible(t,w) = candidate(w) && compatible(t,w) && ask_price[w] <= price_cap[t]
//   value(t,w)    = -(ask_price[w] * S) - price[w]      outside(t) = -((price_cap[t] + 1) * S)
//   bid(t)        = price[w1] + (best - max(second, outside)) + eps    on the best worker w1
//   a worker takes the highest bid (ties: lowest ask index), releasing its previous owner.
//
// Rounds are cheap because every ask keeps a cache of its 32 best feasible workers (at the
// prices of its last full scan) and the value theta that bounds every worker outside the cache:
// prices only rise, so while the cached best beats theta in (value, lowest index) order and the cached
// second-best value is >= theta's, the cached top-2 IS the global top-2, ties included
// (pm_auction_bid_cached: 32 gathers per ask); otherwise the ask rescans all
// workers at the current prices in the same round (pm_auction_bid) — bids are bit-identical to a
// full scan every round, which is what the sequential checker does.
//
// pm_auction_bid: one warp per unassigned ask; the CTA's 8 warps share stripes of the
// worker table (planes A, B) and of the per-worker price vector, staged into shared
// memory with 1-D TMA bulk copies; collisions are resolved by atomicMax on the bid and
// atomicMin on the bidder (the claim), applied by the single winner of each worker.
#pragma once
#include "pm_kernels.cuh"

namespace pm {

constexpr int kAucThreads = 256;
constexpr int kAucWarps = kAucThreads / 32;
constexpr int kAucStripe = 1024;  // workers staged per step: 16 KB + 16 KB + 8 KB
constexpr long long kAucNeg = (long long)(0x8000000000000000ull) / 4;

struct AuctionParams {
  EvalParams ev;
  const uint32_t* price_cap;   // [T]
  long long* price;            // [W] dual price of each worker
  uint32_t* owner;             // [W] ask currently holding the worker
  uint32_t* assigned;          // [T] worker held by the ask
  uint32_t* withdrawn;         // [T]
  const uint32_t* active;      // [n_active] unassigned, not withdrawn asks
  uint32_t n_active;
  uint32_t* bid_w;             // [T]
  long long* bid_p;            // [T]
  long long* bid_max;          // [W] highest bid of the round (reset by the winner)
  uint32_t* winner;            // [W]
  uint32_t* cand;              // [T * 32] cached best feasible workers of each ask
  long long* theta;            // [T] (with theta_w) the best (value, lowest index) any NON-cached worker had at
  uint32_t* theta_w;           //     scan time — an upper bound forever, prices only rise (INT64_MIN: cache complete)
  uint32_t* rescan;            // [T] asks whose cache could not decide this round
  uint32_t* n_rescan;
  const uint32_t* scan_list;   // asks pm_auction_bid scans (= rescan)
  uint32_t n_scan;
  long long scale, eps;
};
constexpr int kAucCache = 32;
constexpr long long kThetaComplete = (long long)0x8000000000000000ull;
constexpr long long kThetaInvalid = 0x7FFFFFFFFFFFFFFFll;

struct __align__(128) AuctionStage {
  uint4 a[kAucStripe];
  uint4 b[kAucStripe];
  long long price[kAucStripe];
  uint64_t bar;
};

struct Top2 {
  long long b1, b2;
  uint32_t w1;
};
__device__ __forceinline__ Top2 top2_merge(const Top2& x, const Top2& y) {
  Top2 r;
  r.b1 = max(x.b1, y.b1);
  r.w1 = (x.b1 > y.b1) ? x.w1 : (y.b1 > x.b1) ? y.w1 : min(x.w1, y.w1);
  r.b2 = max(min(x.b1, y.b1), max(x.b2, y.b2));
  return r;
}

__device__ __forceinline__ void auction_place_bid(const AuctionParams& p, uint32_t t, uint32_t cap,
                                                  long long b1, uint32_t w1, long long b2) {
  const long long outside = -(((long long)cap + 1) * p.scale);
  if (w1 == kNone || b1 < outside) {
    p.withdrawn[t] = 1u;       // prices only rise: it can never come back
    p.bid_w[t] = kNone;
  } else {
    const long long second = max(b2, outside);
    const long long bid = p.price[w1] + (b1 - second) + p.eps;
    p.bid_w[t] = w1;
    p.bid_p[t] = bid;
    atomicMax(p.bid_max + w1, bid);
  }
}

// (value desc, worker asc) arg-max across the warp; returns the winning lane
__device__ __forceinline__ uint32_t warp_argbest(long long v, uint32_t w, long long* bv, uint32_t* bw) {
  uint32_t lane_id = threadIdx.x & 31u;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const long long ov = __shfl_xor_sync(0xffffffffu, v, off);
    const uint32_t ow = __shfl_xor_sync(0xffffffffu, w, off);
    const uint32_t ol = __shfl_xor_sync(0xffffffffu, lane_id, off);
    if (ov > v || (ov == v && ow < w)) { v = ov; w = ow; lane_id = ol; }
  }
  *bv = v;
  *bw = w;
  return lane_id;
}

// Full scan of every worker for the asks in scan_list: exact top-2 -> bid, and the ask's cache.
// TPC asks per CTA share each staged worker stripe; an ask is scanned by 8/TPC warps whose
// candidate lists are merged through shared memory (TPC = 8 for the big first round, TPC = 1 so a
// handful of rescans is not one slow warp each).  Stripes are double-buffered: the bulk copies
// of stripe k+1 are in flight while stripe k is evaluated.
struct AuctionMerge {
  long long v[kAucWarps][32];
  uint32_t w[kAucWarps][32];
  long long bound_v[kAucWarps];
  uint32_t bound_w[kAucWarps];
};

template <int TPC>
__global__ void __launch_bounds__(kAucThreads) pm_auction_bid(AuctionParams p) {
  constexpr int kWpt = kAucWarps / TPC;           // warps per ask
  extern __shared__ __align__(128) unsigned char smem_raw[];
  AuctionStage* stage = reinterpret_cast<AuctionStage*>(smem_raw);            // [2]
  AuctionMerge& mg = *reinterpret_cast<AuctionMerge*>(smem_raw + 2 * sizeof(AuctionStage));
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t slot = blockIdx.x * TPC + warp / kWpt;
  const uint32_t sub = warp % kWpt;
  const bool live = slot < p.n_scan;
  const uint32_t t = live ? p.scan_list[slot] : 0u;
  const DevAsk ask = p.ev.asks[t];
  const uint32_t cap = p.price_cap[t];
  // per lane: its 4 best (value, worker) in order, and the best (value, worker) it did not keep
  long long cv[4] = {kAucNeg, kAucNeg, kAucNeg, kAucNeg};
  uint32_t cw[4] = {kNone, kNone, kNone, kNone};
  long long dropped = kAucNeg;
  uint32_t dropped_w = kNone;

  if (threadIdx.x == 0) { mbar_init(&stage[0].bar, 1); mbar_init(&stage[1].bar, 1); }
  __syncthreads();
  const uint32_t W = p.ev.n_workers;
  const uint32_t n_stripes = (W + kAucStripe - 1) / kAucStripe;
  auto issue = [&](uint32_t k) {
    AuctionStage& s = stage[k & 1u];
    const uint32_t w0 = k * kAucStripe;
    const uint32_t n = min((uint32_t)kAucStripe, W - w0);
    const uint32_t np = (n + 1u) & ~1u;   // bulk copies move multiples of 16 B; price[] is padded
    mbar_expect_tx(&s.bar, n * 32u + np * 8u);
    bulk_g2s(s.a, p.ev.wa + w0, n * 16u, &s.bar);
    bulk_g2s(s.b, p.ev.wb + w0, n * 16u, &s.bar);
    bulk_g2s(s.price, p.price + w0, np * 8u, &s.bar);
  };
  if (threadIdx.x == 0 && n_stripes) issue(0);
  for (uint32_t k = 0; k < n_stripes; ++k) {
    if (threadIdx.x == 0 && k + 1 < n_stripes) issue(k + 1);   // buffer (k+1)&1 was released by the sync below
    AuctionStage& s = stage[k & 1u];
    mbar_wait(&s.bar, (k >> 1) & 1u);
    const uint32_t w0 = k * kAucStripe;
    const uint32_t n = min((uint32_t)kAucStripe, W - w0);
    if (live) {
      for (uint32_t i = sub * 32 + lane; i < n; i += kWpt * 32) {
        const WorkerReg wr = make_worker(s.a[i], s.b[i]);
        if (wr.price <= cap && ask_meets(ask, p.ev.opts, wr, p.ev.bits, p.ev.words)) {
          const long long v = -((long long)wr.price * p.scale) - s.price[i];
          const uint32_t w = w0 + i;                     // ascends per lane: strict '>' keeps ties in index order
          if (v > cv[3]) {
            if (cv[3] > dropped || (cv[3] == dropped && cw[3] < dropped_w)) { dropped = cv[3]; dropped_w = cw[3]; }
            cv[3] = v; cw[3] = w;
            if (cv[3] > cv[2]) { const long long tv = cv[2]; const uint32_t tw = cw[2]; cv[2] = cv[3]; cw[2] = cw[3]; cv[3] = tv; cw[3] = tw; }
            if (cv[2] > cv[1]) { const long long tv = cv[1]; const uint32_t tw = cw[1]; cv[1] = cv[2]; cw[1] = cw[2]; cv[2] = tv; cw[2] = tw; }
            if (cv[1] > cv[0]) { const long long tv = cv[0]; const uint32_t tw = cw[0]; cv[0] = cv[1]; cw[0] = cw[1]; cv[1] = tv; cw[1] = tw; }
          } else if (v > dropped || (v == dropped && w < dropped_w)) {
            dropped = v; dropped_w = w;
          }
        }
      }
    }
    __syncthreads();  // stripe k fully consumed: its buffer may be refilled by issue(k + 2)
  }
  // 32 selection rounds over the warp's 128 candidates: lane r keeps the r-th best
  long long b1 = kAucNeg, b2 = kAucNeg, mine_v = kAucNeg;
  uint32_t w1 = kNone, mine = kNone;
#pragma unroll 1
  for (int r = 0; r < kAucCache; ++r) {
    long long bv;
    uint32_t bw;
    const uint32_t win = warp_argbest(cv[0], cw[0], &bv, &bw);
    if (r == 0) { b1 = bv; w1 = bw; }
    if (r == 1) b2 = bv;
    if ((int)lane == r) { mine = (bv > kAucNeg) ? bw : kNone; mine_v = bv; }
    if (lane == win) { cv[0] = cv[1]; cw[0] = cw[1]; cv[1] = cv[2]; cw[1] = cw[2]; cv[2] = cv[3]; cw[2] = cw[3]; cv[3] = kAucNeg; cw[3] = kNone; }
  }
  long long next_v, drop_v;
  uint32_t next_w, drop_w;
  warp_argbest(cv[0], cw[0], &next_v, &next_w);       // the warp's 33rd best candidate
  warp_argbest(dropped, dropped_w, &drop_v, &drop_w); // the best worker no lane kept
  const bool use_next = next_v > drop_v || (next_v == drop_v && next_w < drop_w);
  long long bound_v = use_next ? next_v : drop_v;
  uint32_t bound_w = use_next ? next_w : drop_w;

  if (kWpt > 1) {
    // merge the kWpt warps of the ask: warp `sub == 0` re-selects from kWpt sorted lists
    mg.v[warp][lane] = mine_v;
    mg.w[warp][lane] = mine;
    if (lane == 0) { mg.bound_v[warp] = bound_v; mg.bound_w[warp] = bound_w; }
    __syncthreads();
    if (sub != 0) return;
    long long lv[kWpt];
    uint32_t lw[kWpt];
#pragma unroll
    for (int q = 0; q < kWpt; ++q) { lv[q] = mg.v[warp + q][lane]; lw[q] = mg.w[warp + q][lane]; }
#pragma unroll
    for (int i = 1; i < kWpt; ++i)      // sort the lane's kWpt entries by (value desc, worker asc)
#pragma unroll
      for (int j = i; j > 0; --j)
        if (lv[j] > lv[j - 1] || (lv[j] == lv[j - 1] && lw[j] < lw[j - 1])) {
          const long long tv = lv[j]; const uint32_t tw = lw[j];
          lv[j] = lv[j - 1]; lw[j] = lw[j - 1]; lv[j - 1] = tv; lw[j - 1] = tw;
        }
    b1 = kAucNeg; b2 = kAucNeg; w1 = kNone; mine = kNone;
#pragma unroll 1
    for (int r = 0; r < kAucCache; ++r) {
      long long bv;
      uint32_t bw;
      const uint32_t win = warp_argbest(lv[0], lw[0], &bv, &bw);
      if (r == 0) { b1 = bv; w1 = bw; }
      if (r == 1) b2 = bv;
      if ((int)lane == r) mine = (bv > kAucNeg) ? bw : kNone;
      if (lane == win) {
#pragma unroll
        for (int q = 0; q + 1 < kWpt; ++q) { lv[q] = lv[q + 1]; lw[q] = lw[q + 1]; }
        lv[kWpt - 1] = kAucNeg; lw[kWpt - 1] = kNone;
      }
    }
    warp_argbest(lv[0], lw[0], &bound_v, &bound_w);   // 33rd of the merged lists ...
#pragma unroll
    for (int q = 0; q < kWpt; ++q) {                  // ... against everything the warps left out
      const long long qv = mg.bound_v[warp + q];
      const uint32_t qw = mg.bound_w[warp + q];
      if (qv > bound_v || (qv == bound_v && qw < bound_w)) { bound_v = qv; bound_w = qw; }
    }
  }
  if (!live) return;
  p.cand[(size_t)t * kAucCache + lane] = mine;
  if (lane == 0) {
    p.theta[t] = (bound_v == kAucNeg) ? kThetaComplete : bound_v;
    p.theta_w[t] = bound_w;
    auction_place_bid(p, t, cap, b1, w1, b2);
  }
}

// One round for an ask from its cache: 32 gathers instead of a scan of every worker.
__global__ void __launch_bounds__(kAucThreads) pm_auction_bid_cached(AuctionParams p) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t slot = blockIdx.x * kAucWarps + warp;
  if (slot >= p.n_active) return;
  const uint32_t t = p.active[slot];
  const long long theta = p.theta[t];
  const uint32_t w = p.cand[(size_t)t * kAucCache + lane];
  Top2 best{kAucNeg, kAucNeg, kNone};
  if (theta != kThetaInvalid && w != kNone) {
    best.b1 = -((long long)p.ev.wb[w].w * p.scale) - p.price[w];
    best.w1 = w;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    Top2 o;
    o.b1 = __shfl_xor_sync(0xffffffffu, best.b1, off);
    o.b2 = __shfl_xor_sync(0xffffffffu, best.b2, off);
    o.w1 = __shfl_xor_sync(0xffffffffu, best.w1, off);
    best = top2_merge(best, o);
  }
  if (lane == 0) {
    // every worker outside the cache ranks at or below (theta, theta_w) — its rank at scan time; prices only rise
    const uint32_t theta_w = p.theta_w[t];
    const bool decided = theta == kThetaComplete ||
                         (theta != kThetaInvalid && best.b2 >= theta &&
                          (best.b1 > theta || (best.b1 == theta && best.w1 < theta_w)));
    if (decided) auction_place_bid(p, t, p.price_cap[t], best.b1, best.w1, best.b2);
    else { p.bid_w[t] = kNone; p.rescan[atomicAdd(p.n_rescan, 1u)] = t; }
  }
}

// the claim: among the highest bidders of a worker the lowest ask index wins
__global__ void pm_auction_claim(AuctionParams p) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n_active) return;
  const uint32_t t = p.active[i], w = p.bid_w[t];
  if (w != kNone && p.bid_p[t] == p.bid_max[w]) atomicMin(p.winner + w, t);
}

__global__ void pm_auction_apply(AuctionParams p) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n_active) return;
  const uint32_t t = p.active[i], w = p.bid_w[t];
  if (w == kNone || p.winner[w] != t) return;
  const uint32_t prev = p.owner[w];
  if (prev != kNone) p.assigned[prev] = kNone;   // prev holds a worker, so it did not bid this round
  p.owner[w] = t;
  p.assigned[t] = w;
  p.price[w] = p.bid_p[t];
  p.bid_max[w] = kAucNeg;
  p.winner[w] = kNone;
}

__global__ void pm_auction_compact(const uint32_t* __restrict__ assigned, const uint32_t* __restrict__ withdrawn,
                                   uint32_t n_asks, uint32_t* __restrict__ active, uint32_t* __restrict__ n_active) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_asks) return;
  if (assigned[t] == kNone && !withdrawn[t]) active[atomicAdd(n_active, 1u)] = t;
}

// result in the engine's group form: one solo group per assigned ask, in ask order
__global__ void pm_auction_flags(const uint32_t* __restrict__ assigned, uint32_t n_asks, uint32_t* __restrict__ flag) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_asks) flag[t] = assigned[t] != kNone ? 1u : 0u;
  if (t == n_asks) flag[t] = 0u;
}
__global__ void pm_auction_emit(const uint32_t* __restrict__ assigned, const uint32_t* __restrict__ gidx,
                                const uint4* __restrict__ wb, uint32_t n_asks, uint32_t* __restrict__ worker_group,
                                uint32_t* __restrict__ worker_ask, uint32_t* __restrict__ group_ask,
                                uint32_t* __restrict__ group_off, uint32_t* __restrict__ members,
                                long long* __restrict__ ask_best, uint32_t* __restrict__ ask_count) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_asks) return;
  const uint32_t w = assigned[t];
  if (w == kNone) { ask_best[t] = kInf; ask_count[t] = 0; return; }
  const uint32_t g = gidx[t];
  worker_group[w] = g;
  worker_ask[w] = t;
  group_ask[g] = t;
  group_off[g] = g;
  members[g] = w;
  ask_best[t] = ((long long)wb[w].w << 32) | (long long)w;
  ask_count[t] = 1;
}

}  // namespace pm
