// Batched fixed-base MSM for many commitments at once (the throughput path of the prover: B proofs x columns MSMs per call).
//
// Replaces halo2_proofs `Params::{commit, commit_lagrange}` -> `arithmetic::best_multiexp` (EXT, called under
// taiga_halo2/src/proof.rs:33-40) for K MSMs that share the SRS basis.  The bases were
// premultiplied by 2^(c*w) at SRS load (srs.cuh), so the W windows of one MSM share ONE set of NB = 2^(c-1) buckets.
//
//   1. msm_sort_kernel      one CTA per MSM: signed c-bit digits of every scalar, bucket histogram and counting sort
//                           entirely in shared memory (no global atomics, no scan launches).  The scalars are staged
//                           through shared memory by TMA (cp.async.bulk.tensor, two-stage mbarrier pipeline).
//   2. msm_ba_round_kernel  R rounds of pairwise reduction inside every bucket with BATCH-AFFINE additions: a CTA takes
//                           2048 pairs, multiplies their denominators together (per-thread prefix products, then a
//                           product tree across the 256 threads in shared memory), inverts ONCE, and walks back.  An
//                           addition costs 6 field multiplications + ~1/2048 of an inversion instead of the 10 of the
//                           XYZZ mixed addition: the path is bound by the integer pipe (tools/modmul_bench.cu), so
//                           multiplications are what counts.  Round r halves every bucket: ceil(c/2^r) items are left,
//                           so the offsets of every round follow from the round-0 counts and are re-derived per CTA
//                           by a 4096-element shared-memory scan -- no per-round bookkeeping in global memory.
//   3. msm_ba_finish_kernel the (normally single) item left in every bucket becomes the XYZZ bucket sum that the
//                           two-level weighted bucket reduction of msm.cu consumes.
//
// Exceptional pairs (equal points, opposite points, the identity) are handled inside the batch: their denominator is
// replaced (2y for a doubling) or left out, so any input -- including the structured SRS of the tests -- is exact.
//
// Algorithmic bytes: 64*N*W (table) + 32*N*K.  The rounds deliberately spend HBM bytes (each round writes its items) to
// save integer instructions; DESIGN.md section 4 has the accounting.
#define TB_NOINLINE_MUL 0   // every loop of this file is rolled (small code): inline multiplies keep live values in registers instead of spilling them around calls
#include <cuda.h>
#include <algorithm>
#include <memory>
#include <type_traits>
#include "common.cuh"
#include "kernels.cuh"

namespace tb {

constexpr int BA_THREADS = 256;
constexpr int BA_M = 32;                        // default pairs per thread (TB_MSM_BA_M)
constexpr int BA_PAIRS = BA_THREADS * BA_M;     // output items per CTA
constexpr int BA_MAX_NB = 4096;
constexpr int SORT_THREADS = 1024;
constexpr int SORT_TILE = 1024;                 // scalars per pipeline stage (32 KB), one per thread
constexpr int SORT_BOX = 256;                   // scalars per TMA box (a box dimension is limited to 256)

// ---------------------------------------------------------------- small shared-memory helpers
// exclusive scan of v[0..n) in place (n <= 16 * blockDim.x), v[n] = total.  blockDim.x threads, tmp: 32 words.
__device__ __forceinline__ void block_scan_excl(uint32_t* v, int n, uint32_t* tmp) {
  const int T = blockDim.x, t = threadIdx.x, per = (n + T - 1) / T;
  const int b0 = t * per, b1 = min(n, b0 + per);
  uint32_t local = 0;
  for (int i = b0; i < b1; ++i) local += v[i];
  uint32_t incl = local;
  const int lane = t & 31, warp = t >> 5, nw = (T + 31) >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
  if (lane == 31) tmp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < nw ? tmp[lane] : 0u, wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += o; }
    if (lane < nw) tmp[lane] = wi - w;
    if (lane == nw - 1) v[n] = wi;
  }
  __syncthreads();
  uint32_t run = tmp[warp] + incl - local;
  for (int i = b0; i < b1; ++i) { uint32_t x = v[i]; v[i] = run; run += x; }
  __syncthreads();
}

// ---------------------------------------------------------------- TMA / mbarrier primitives (sm_100a PTX)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a TMA that never lands must not hang the GPU box -- trap instead
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
// box (8 words, SORT_BOX scalars, 1 MSM) of the 3-D scalar tensor [K][N][8 x u32] -> shared memory
__device__ __forceinline__ void tma_load_tile(void* dst, const CUtensorMap* map, uint64_t* bar, int scalar0, int item) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(0), "r"(scalar0), "r"(item) : "memory");
}

// ---------------------------------------------------------------- 1. digits + counting sort, one CTA per MSM
// entries of MSM k: entries[k * cap0 + pos] = (w * table_stride + i) | sign << 31, grouped by bucket; counts[k * NB + b]
// CT > 0: window width known at compile time (the digit loop unrolls and the limbs stay in registers); CT = 0: generic
template <class S, int CT>
__global__ void __launch_bounds__(SORT_THREADS) msm_sort_kernel(const __grid_constant__ CUtensorMap smap, const S* __restrict__ extras, int N, int n_extra, int c_rt, int W_rt,
                                                                 int NB, int table_stride, uint32_t* __restrict__ counts, uint32_t* __restrict__ entries, long long cap0,
                                                                 unsigned long long* __restrict__ total_entries) {
  const int c = CT ? CT : c_rt, W = CT ? (256 + CT - 1) / CT : W_rt;
  extern __shared__ __align__(128) uint8_t sort_smem[];
  S* tile = reinterpret_cast<S*>(sort_smem);                                   // [2][SORT_TILE]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sort_smem + 2 * SORT_TILE * sizeof(S));   // [2], 8-byte aligned
  uint32_t* hist = reinterpret_cast<uint32_t*>(bars + 2);                       // [NB + 1]
  uint32_t* tmp = hist + NB + 1;                                                // [32]
  const int k = blockIdx.x, t = threadIdx.x;
  const uint32_t half = 1u << (c - 1);
  const int ntiles = (N + SORT_TILE - 1) / SORT_TILE;
  if (t == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int b = t; b <= NB; b += SORT_THREADS) hist[b] = 0;
  __syncthreads();
  uint32_t* ent = entries + (long long)k * cap0;

  // `scatter` is a compile-time constant of each pass: the histogram pass issues its shared-memory atomics without waiting for a
  // return value (no scoreboard stall), only the scatter pass needs the old counter
  auto digits = [&](const S& sm, uint32_t idx, auto scatter_c) {
    constexpr bool scatter = decltype(scatter_c)::value;
    S s = sm.from_mont();
    if (s.is_zero()) return;
    uint32_t carry = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const int bit = w * c, limb = bit >> 5, off = bit & 31;
      uint64_t v64 = s.l[limb];
      if (limb + 1 < 8) v64 |= (uint64_t)s.l[limb + 1] << 32;
      uint32_t v = ((uint32_t)(v64 >> off) & ((1u << c) - 1)) + carry, neg = 0;
      if (v > half) { v = (1u << c) - v; neg = 1; carry = 1; } else carry = 0;
      if (v) {
        if (scatter) { const uint32_t pos = atomicAdd(&hist[v - 1], 1u); ent[pos] = (uint32_t)(w * table_stride + idx) | (neg << 31); }
        else atomicAdd(&hist[v - 1], 1u);
      }
    }
  };
  // two passes over the scalars (histogram, then scatter with the scanned histogram as cursors); tiles arrive by TMA
  auto issue = [&](int stage, int it) {   // thread 0: fill `stage` with scalars [it * SORT_TILE, (it + 1) * SORT_TILE) of MSM k
    const int s0 = it * SORT_TILE;
    int nbox = (N - s0 + SORT_BOX - 1) / SORT_BOX; if (nbox > SORT_TILE / SORT_BOX) nbox = SORT_TILE / SORT_BOX;
    mbar_expect_tx(&bars[stage], (uint32_t)(nbox * SORT_BOX * sizeof(S)));
    for (int j = 0; j < nbox; ++j) tma_load_tile(tile + stage * SORT_TILE + j * SORT_BOX, &smap, &bars[stage], s0 + j * SORT_BOX, k);
  };
  uint32_t phase[2] = {0, 0};
  for (int pass = 0; pass < 2; ++pass) {
    if (t == 0) issue(0, 0);
    for (int it = 0; it < ntiles; ++it) {
      const int st = it & 1;
      if (t == 0 && it + 1 < ntiles) issue(st ^ 1, it + 1);   // the other stage: its readers finished before the last barrier
      mbar_wait(&bars[st], phase[st]); phase[st] ^= 1;
      const int i = it * SORT_TILE + t;
      if (i < N) { if (pass) digits(tile[st * SORT_TILE + t], (uint32_t)i, std::true_type()); else digits(tile[st * SORT_TILE + t], (uint32_t)i, std::false_type()); }
      __syncthreads();   // everybody is done with stage st before it is refilled
    }
    if (t < n_extra) { const S ex = ldg_fe(extras + (long long)k * n_extra + t); if (pass) digits(ex, (uint32_t)(N + t), std::true_type()); else digits(ex, (uint32_t)(N + t), std::false_type()); }
    __syncthreads();
    if (pass == 0) {
      for (int b = t; b < NB; b += SORT_THREADS) counts[(long long)k * NB + b] = hist[b];
      __syncthreads();
      block_scan_excl(hist, NB, tmp);
      if (t == 0) atomicAdd(total_entries, (unsigned long long)hist[NB]);
    }
  }
}

// ---------------------------------------------------------------- 2. batch-affine pairwise reduction round
template <class F> __device__ __forceinline__ void sts_fe(uint4* lo, uint4* hi, int idx, const F& v) {
  lo[idx] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]); hi[idx] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
template <class F> __device__ __forceinline__ F lds_fe(const uint4* lo, const uint4* hi, int idx) {
  uint4 x = lo[idx], y = hi[idx]; F v;
  v.l[0] = x.x; v.l[1] = x.y; v.l[2] = x.z; v.l[3] = x.w; v.l[4] = y.x; v.l[5] = y.y; v.l[6] = y.z; v.l[7] = y.w; return v;
}

enum { PK_NONE = 0, PK_COPY1, PK_COPY2, PK_ADD, PK_DBL, PK_INF };

// what a round needs to know about one MSM: the bucket offsets of its input (round r) and output (round r + 1) items
struct RoundOffsets {
  uint32_t* off_cur; uint32_t* off_nxt; uint32_t n_next;
  __device__ __forceinline__ void build(uint8_t* smem, const uint32_t* __restrict__ counts0, int NB, int round) {
    off_cur = reinterpret_cast<uint32_t*>(smem); off_nxt = off_cur + (NB + 4);
    uint32_t* tmp = off_nxt + (NB + 4);
    for (int b = threadIdx.x; b < NB; b += blockDim.x) {
      const uint32_t c0 = counts0[b], cr = (c0 + ((1u << round) - 1)) >> round;
      off_cur[b] = cr; off_nxt[b] = (cr + 1) >> 1;
    }
    __syncthreads();
    block_scan_excl(off_cur, NB, tmp);
    block_scan_excl(off_nxt, NB, tmp);
    n_next = off_nxt[NB];
  }
  // output item q -> position of its first input, and whether a second input exists.  `hint`: a bucket at or before the one of q
  // (the previous item of the same thread): a few steps forward usually find it, a binary search takes over otherwise.
  __device__ __forceinline__ void locate(uint32_t q, int NB, uint32_t& in0, bool& two, int& hint) const {
    int lo = hint;   // invariant: off_nxt[lo] <= q
#pragma unroll 1
    for (int s = 0; s < 6 && off_nxt[lo + 1] <= q; ++s) ++lo;
    if (off_nxt[lo + 1] <= q) {
      int hi = NB;   // largest b with off_nxt[b] <= q
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off_nxt[mid] <= q) lo = mid; else hi = mid; }
    }
    hint = lo;
    in0 = off_cur[lo] + 2 * (q - off_nxt[lo]);
    two = in0 + 1 < off_cur[lo + 1];
  }
};
constexpr size_t ba_off_bytes(int NB) { return (size_t)(2 * (NB + 4) + 32) * 4; }

template <class B, bool FIRST> struct PairLoader {
  const uint32_t* entries; const Aff<B>* table; const Aff<B>* items;
  __device__ __forceinline__ Aff<B> point(uint32_t pos) const {
    if (FIRST) {
      const uint32_t e = __ldg(entries + pos);
      Aff<B> p = ldg_aff(table + (e & 0x7fffffffu));
      if (e >> 31) p.y = p.y.neg();
      return p;
    }
    return ldg_aff(items + pos);
  }
  __device__ __forceinline__ B x(uint32_t pos) const {   // the identity is (0, 0): two identities give den = 0 and take the slow path
    if (FIRST) return ldg_fe(&table[__ldg(entries + pos) & 0x7fffffffu].x);
    return ldg_fe(&items[pos].x);
  }
};
// classification of a pair and its denominator; identical in the forward and the backward kernel
template <class B> __device__ __forceinline__ int classify_pair(const Aff<B>& p1, const Aff<B>& p2, bool two, B& den) {
  if (!two || p2.is_inf()) return PK_COPY1;
  if (p1.is_inf()) return PK_COPY2;
  den = p2.x - p1.x;
  if (!den.is_zero()) return PK_ADD;
  if (p1.y == p2.y && !p1.y.is_zero()) { den = p1.y.dbl(); return PK_DBL; }
  return PK_INF;
}

// ---- 2.0 items per MSM after every round (n_items[k * (R + 1) + r] = sum_b ceil(c_b / 2^r)): lets the CTAs of a round that have
// nothing to do leave before they build any offsets (sparse witness columns leave most of the worst-case grid idle)
__global__ void __launch_bounds__(BA_THREADS) msm_ba_count_kernel(const uint32_t* __restrict__ counts0, int NB, int R, uint32_t* __restrict__ n_items) {
  __shared__ uint32_t red[BA_THREADS / 32];
  const int k = blockIdx.x, t = threadIdx.x;
  for (int r = 0; r <= R; ++r) {
    uint32_t s = 0;
    for (int b = t; b < NB; b += BA_THREADS) s += (counts0[(long long)k * NB + b] + ((1u << r) - 1)) >> r;
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
    if ((t & 31) == 0) red[t >> 5] = s;
    __syncthreads();
    if (t == 0) { uint32_t tot = 0; for (int w = 0; w < BA_THREADS / 32; ++w) tot += red[w]; n_items[(long long)k * (R + 1) + r] = tot; }
    __syncthreads();
  }
}

// ---- 2a. forward: prefix products of the denominators (to global memory), product tree of the CTA (to global memory)
// A CTA owns BA_PAIRS consecutive output items of one MSM; thread t owns items Q0 + i * BA_THREADS + t.
template <class B, bool FIRST, int M>
__global__ void __launch_bounds__(BA_THREADS) msm_ba_fwd_kernel(const uint32_t* __restrict__ counts0, int NB, int round, const uint32_t* __restrict__ entries,
                                                                 const Aff<B>* __restrict__ table, const Aff<B>* __restrict__ items_in, long long cap_in, long long cap_out,
                                                                 B* __restrict__ pre, B* __restrict__ tree, uint32_t* __restrict__ meta, const uint32_t* __restrict__ n_items, int R) {
  extern __shared__ __align__(16) uint8_t ba_smem[];
  const int k = blockIdx.y, t = threadIdx.x;
  const uint32_t Q0 = blockIdx.x * (M * BA_THREADS);
  if (Q0 >= n_items[(long long)k * (R + 1) + round + 1]) return;   // nothing of this MSM left for this CTA (uniform over the CTA)
  RoundOffsets ro; ro.build(ba_smem, counts0 + (long long)k * NB, NB, round);
  PairLoader<B, FIRST> ld{entries + (FIRST ? (long long)k * cap_in : 0), table, items_in + (FIRST ? 0 : (long long)k * cap_in)};
  B* pre_k = pre + (long long)k * cap_out;
  uint32_t* meta_k = meta + (long long)k * cap_out;
  B acc = B::one();
  int hint = 0;
#pragma unroll 1
  for (int i = 0; i < M; ++i) {
    const uint32_t q = Q0 + (uint32_t)i * BA_THREADS + t;
    if (q >= ro.n_next) break;
    uint32_t in0; bool two; ro.locate(q, NB, in0, two, hint);
    meta_k[q] = in0 | (two ? 0x80000000u : 0u);   // the backward kernel does not search again
    st_fe(pre_k + q, acc);
    if (!two) continue;
    const B x1 = ld.x(in0), x2 = ld.x(in0 + 1);    // the common case needs the x coordinates only
    B den = x2 - x1;
    if (den.is_zero() || x1.is_zero() || x2.is_zero()) {   // equal x (doubling / cancellation) or possibly an identity (0, 0): classify on the full points
      const Aff<B> p1 = ld.point(in0), p2 = ld.point(in0 + 1);
      const int kind = classify_pair(p1, p2, two, den);
      if (kind != PK_ADD && kind != PK_DBL) continue;
    }
    acc = acc * den;
  }
  __syncthreads();   // the offset arrays are dead: the product tree (heap layout, nd[1] = root, leaves nd[BA_THREADS + t]) takes their place
  B* nd = reinterpret_cast<B*>(ba_smem);
  nd[BA_THREADS + t] = acc;
  __syncthreads();
  for (int w = BA_THREADS >> 1; w >= 1; w >>= 1) {
    if (t < w) nd[w + t] = nd[2 * (w + t)] * nd[2 * (w + t) + 1];
    __syncthreads();
  }
  B* tr = tree + ((long long)k * gridDim.x + blockIdx.x) * (2 * BA_THREADS);
  st_fe(tr + t, t ? nd[t] : B::one()); st_fe(tr + BA_THREADS + t, nd[BA_THREADS + t]);
}

// ---- 2b. the roots of all CTAs are inverted together, one per thread (every lane busy: nobody waits for a lone inversion)
template <class B>
__global__ void msm_ba_inv_kernel(B* __restrict__ tree, uint32_t n_trees) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_trees) return;
  B* root = tree + (size_t)i * (2 * BA_THREADS) + 1;
  st_fe(root, ld_fe(root).inv());
}

// ---- 2c. backward: push the inverted root down the tree, then walk every thread's pairs back and write the sums
template <class B, bool FIRST, int M, int MINB>
__global__ void __launch_bounds__(BA_THREADS, MINB) msm_ba_bwd_kernel(const uint32_t* __restrict__ counts0, int NB, int round, const uint32_t* __restrict__ entries,
                                                                 const Aff<B>* __restrict__ table, const Aff<B>* __restrict__ items_in, long long cap_in,
                                                                 Aff<B>* __restrict__ items_out, long long cap_out, const B* __restrict__ pre, const B* __restrict__ tree,
                                                                 const uint32_t* __restrict__ meta, const uint32_t* __restrict__ n_items, int R) {
  extern __shared__ __align__(16) uint8_t ba_smem[];
  const int k = blockIdx.y, t = threadIdx.x;
  const uint32_t Q0 = blockIdx.x * (M * BA_THREADS);
  const uint32_t n_next = n_items[(long long)k * (R + 1) + round + 1];
  if (Q0 >= n_next) return;
  B* nd = reinterpret_cast<B*>(ba_smem);   // [2 * BA_THREADS]
  { const B* tr = tree + ((long long)k * gridDim.x + blockIdx.x) * (2 * BA_THREADS);
    nd[t] = ld_fe(tr + t); nd[BA_THREADS + t] = ld_fe(tr + BA_THREADS + t); }
  __syncthreads();
  for (int w = 1; w < BA_THREADS; w <<= 1) {   // node i in [w, 2w) holds the inverse of its subtree product: inverse(child) = inverse(parent) * product(sibling)
    B child_inv;
    const int node = w + (t >> 1);
    if (t < 2 * w) child_inv = nd[node] * nd[2 * node + ((t & 1) ^ 1)];
    __syncthreads();
    if (t < 2 * w) nd[2 * node + (t & 1)] = child_inv;
    __syncthreads();
  }
  B inv_run = nd[BA_THREADS + t];   // 1 / (product of this thread's denominators)
  PairLoader<B, FIRST> ld{entries + (FIRST ? (long long)k * cap_in : 0), table, items_in + (FIRST ? 0 : (long long)k * cap_in)};
  const B* pre_k = pre + (long long)k * cap_out;
  const uint32_t* meta_k = meta + (long long)k * cap_out;
  Aff<B>* out = items_out + (long long)k * cap_out;
  int last = M - 1;
  while (last >= 0 && Q0 + (uint32_t)last * BA_THREADS + t >= n_next) --last;
#pragma unroll 1
  for (int i = last; i >= 0; --i) {
    const uint32_t q = Q0 + (uint32_t)i * BA_THREADS + t;
    const uint32_t mt = __ldg(meta_k + q);
    const uint32_t in0 = mt & 0x7fffffffu; const bool two = mt >> 31;
    const Aff<B> p1 = ld.point(in0);
    Aff<B> p2 = p1;
    if (two) p2 = ld.point(in0 + 1);
    B den;
    const int kind = classify_pair(p1, p2, two, den);
    Aff<B> r;
    if (kind == PK_COPY1) r = p1;
    else if (kind == PK_COPY2) r = p2;
    else if (kind == PK_INF) r = Aff<B>::inf();
    else {
      const B dinv = inv_run * ldg_fe(pre_k + q);   // 1 / den
      inv_run = inv_run * den;
      B num;
      if (kind == PK_ADD) num = p2.y - p1.y;
      else { const B x2 = p1.x.sqr(); num = x2.dbl() + x2; }
      const B lam = num * dinv;
      r.x = lam.sqr() - p1.x - p2.x;
      r.y = lam * (p1.x - r.x) - p1.y;
    }
    st_fe(&out[q].x, r.x); st_fe(&out[q].y, r.y);
  }
}

// ---------------------------------------------------------------- 3. what is left in every bucket -> XYZZ bucket sums
template <class B>
__global__ void __launch_bounds__(BA_THREADS) msm_ba_finish_kernel(const uint32_t* __restrict__ counts0, int NB, int round, const Aff<B>* __restrict__ items, long long cap,
                                                                    Xyzz<B>* __restrict__ buckets) {
  extern __shared__ __align__(16) uint8_t fin_smem[];
  uint32_t* off = reinterpret_cast<uint32_t*>(fin_smem);   // [NB + 1]
  uint32_t* tmp = off + NB + 4;
  const int k = blockIdx.y, t = threadIdx.x;
  for (int b = t; b < NB; b += BA_THREADS) { uint32_t c0 = counts0[(long long)k * NB + b]; off[b] = (c0 + ((1u << round) - 1)) >> round; }
  __syncthreads();
  block_scan_excl(off, NB, tmp);
  const int b = blockIdx.x * BA_THREADS + t;
  if (b >= NB) return;
  Xyzz<B> acc = Xyzz<B>::inf();
  const Aff<B>* it = items + (long long)k * cap;
  for (uint32_t p = off[b]; p < off[b + 1]; ++p) acc.add_affine(ldg_aff(it + p));
  buckets[(long long)k * NB + b] = acc;
}

// ---------------------------------------------------------------- host driver
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr; cudaDriverEntryPointQueryResult qr;
    TB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr));
    if (qr != cudaDriverEntryPointSuccess || !p) throw CudaError("cuTensorMapEncodeTiled is not available from this driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// one reduction round = forward, root inversion, backward; (pairs per thread, resident CTAs per SM) are tuning parameters
template <class B, bool FIRST, int M, int MINB>
static void launch_round_t(Ctx* ctx, dim3 grid, size_t fwd_smem, size_t bwd_smem, const uint32_t* counts, int NB, int r, const uint32_t* entries, const Aff<B>* table, const Aff<B>* in,
                           long long cap_in, Aff<B>* out, long long cap_out, B* pre, B* tree, uint32_t* meta, const uint32_t* n_items, int R, uint32_t n_trees) {
  cudaStream_t st = ctx->stream;
  ctx->opt_in_smem(msm_ba_fwd_kernel<B, FIRST, M>, fwd_smem);
  ctx->opt_in_smem(msm_ba_bwd_kernel<B, FIRST, M, MINB>, bwd_smem);
  msm_ba_fwd_kernel<B, FIRST, M><<<grid, BA_THREADS, fwd_smem, st>>>(counts, NB, r, entries, table, in, cap_in, cap_out, pre, tree, meta, n_items, R);
  msm_ba_inv_kernel<B><<<(n_trees + 63) / 64, 64, 0, st>>>(tree, n_trees);
  msm_ba_bwd_kernel<B, FIRST, M, MINB><<<grid, BA_THREADS, bwd_smem, st>>>(counts, NB, r, entries, table, in, cap_in, out, cap_out, pre, tree, meta, n_items, R);
}
template <class B, typename... A> static void launch_round(Ctx* ctx, int M, int minb, bool first, A... a) {
#define TB_LR(MM, NN) (first ? launch_round_t<B, true, MM, NN>(ctx, a...) : launch_round_t<B, false, MM, NN>(ctx, a...))
  if (M == 8) { if (minb == 2) TB_LR(8, 2); else if (minb == 3) TB_LR(8, 3); else TB_LR(8, 4); }
  else if (M == 32) { if (minb == 2) TB_LR(32, 2); else if (minb == 3) TB_LR(32, 3); else TB_LR(32, 4); }
  else { if (minb == 2) TB_LR(16, 2); else if (minb == 3) TB_LR(16, 3); else TB_LR(16, 4); }
#undef TB_LR
}

bool msm_batch_applicable(int N, int K, const MsmConfig& cfg, int c) {
  if (cfg.table_windows <= 0 || (1 << (c - 1)) > BA_MAX_NB || c < 6) return false;
  const long long min_terms = tb_tune("TB_MSM_BA_MIN_TERMS", 1 << 21);
  return (long long)N * K >= min_terms;
}

// bucket sums of K fixed-base MSMs (all windows of an MSM share NB buckets): buckets[k * NB + b], XYZZ
template <class B, class S>
void msm_batch_buckets(Ctx* ctx, const S* scalars, long long sstride, const Aff<B>* table, int N, int K, int c, int W, int table_stride, const S* extras, int n_extra,
                       Xyzz<B>* buckets) {
  const int NB = 1 << (c - 1);
  cudaStream_t st = ctx->stream;
  TB_REQUIRE(((uintptr_t)scalars & 15) == 0 && (sstride * (long long)sizeof(S)) % 16 == 0, "scalar vectors must be 16-byte aligned for TMA");
  const long long cap0 = (long long)(N + n_extra) * W;
  std::vector<long long> cap(1, cap0);
  const int R = tb_tune("TB_MSM_BA_ROUNDS", 10);
  TB_REQUIRE(R >= 1 && R <= 20, "TB_MSM_BA_ROUNDS out of range");
  for (int r = 0; r < R; ++r) cap.push_back((cap.back() + NB + 1) / 2);
  const int Kc_max = tb_tune("TB_MSM_BA_CHUNK", 512);
  const int Kc = K < Kc_max ? K : Kc_max;
  DevBuf<uint32_t> counts(ctx, (size_t)Kc * NB), entries(ctx, (size_t)Kc * cap0);
  DevBuf<Aff<B>> itA(ctx, (size_t)Kc * cap[1]), itB(ctx, (size_t)Kc * (R > 1 ? cap[2] : 1));
  const size_t sort_smem = 2 * SORT_TILE * sizeof(S) + 16 + (size_t)(NB + 1 + 32) * 4 + 16;
  const size_t fwd_smem = std::max(ba_off_bytes(NB), (size_t)2 * BA_THREADS * sizeof(B));
  const size_t bwd_smem = (size_t)2 * BA_THREADS * sizeof(B);
  const size_t fin_smem = (size_t)(NB + 4 + 32) * 4;
  ctx->opt_in_smem(msm_sort_kernel<S, 13>, sort_smem);
  ctx->opt_in_smem(msm_sort_kernel<S, 0>, sort_smem);
  ctx->opt_in_smem(msm_ba_finish_kernel<B>, fin_smem);
  const int Mv = tb_tune("TB_MSM_BA_M", 32) >= 32 ? 32 : tb_tune("TB_MSM_BA_M", 32) <= 8 ? 8 : 16, minb = tb_tune("TB_MSM_BA_MINB", 3) <= 2 ? 2 : tb_tune("TB_MSM_BA_MINB", 3) >= 4 ? 4 : 3;
  const long long pairs_per_cta = (long long)Mv * BA_THREADS;
  const unsigned ctas1 = (unsigned)((cap[1] + pairs_per_cta - 1) / pairs_per_cta);
  DevBuf<B> pre(ctx, (size_t)Kc * cap[1]), tree(ctx, (size_t)Kc * ctas1 * 2 * BA_THREADS);
  DevBuf<uint32_t> meta(ctx, (size_t)Kc * cap[1]), n_items(ctx, (size_t)Kc * (R + 1));
  for (int k0 = 0; k0 < K; k0 += Kc) {
    const int kc = K - k0 < Kc ? K - k0 : Kc;
    { ProfScope ps(ctx, PC_MSM_SORT);
      // 3-D tensor over the scalar vectors of this chunk: [kc][N][8 x u32], box = 8 x SORT_BOX x 1
      CUtensorMap smap;
      const cuuint64_t dims[3] = {8, (cuuint64_t)N, (cuuint64_t)kc};
      const cuuint64_t strides[2] = {sizeof(S), (cuuint64_t)sstride * sizeof(S)};
      const cuuint32_t box[3] = {8, SORT_BOX, 1}, estr[3] = {1, 1, 1};
      CUresult cr = encode_tiled_fn()(&smap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<S*>(scalars + (long long)k0 * sstride), dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr != CUDA_SUCCESS) throw CudaError("cuTensorMapEncodeTiled failed for the scalar tensor (" + std::to_string((int)cr) + ")");
      const S* ex = extras ? extras + (long long)k0 * n_extra : nullptr;
      if (c == 13) msm_sort_kernel<S, 13><<<kc, SORT_THREADS, sort_smem, st>>>(smap, ex, N, extras ? n_extra : 0, c, W, NB, table_stride, counts.get(), entries.get(), cap0, ctx->d_msm_adds);
      else msm_sort_kernel<S, 0><<<kc, SORT_THREADS, sort_smem, st>>>(smap, ex, N, extras ? n_extra : 0, c, W, NB, table_stride, counts.get(), entries.get(), cap0, ctx->d_msm_adds);
      TB_LAUNCH_CHECK(); ctx->launches++; }
    { ProfScope ps(ctx, PC_MSM_ACCUM);
      msm_ba_count_kernel<<<kc, BA_THREADS, 0, st>>>(counts.get(), NB, R, n_items.get());
      for (int r = 0; r < R; ++r) {
        const Aff<B>* in = (r & 1) ? itA.get() : itB.get();   // round r reads what round r-1 wrote (r = 0 reads the entries)
        Aff<B>* out = (r & 1) ? itB.get() : itA.get();
        const unsigned gx = (unsigned)((cap[r + 1] + pairs_per_cta - 1) / pairs_per_cta);
        dim3 grid(gx, kc);
        const uint32_t n_trees = gx * (uint32_t)kc;
        launch_round<B>(ctx, Mv, minb, r == 0, grid, fwd_smem, bwd_smem, counts.get(), NB, r, entries.get(), table, in, r == 0 ? cap0 : cap[r], out, cap[r + 1], pre.get(), tree.get(),
                        meta.get(), n_items.get(), R, n_trees);
        TB_LAUNCH_CHECK(); ctx->launches += 3;
      }
      const Aff<B>* last = (R & 1) ? itA.get() : itB.get();
      msm_ba_finish_kernel<B><<<dim3((NB + BA_THREADS - 1) / BA_THREADS, kc), BA_THREADS, fin_smem, st>>>(counts.get(), NB, R, last, cap[R], buckets + (size_t)k0 * NB);
      TB_LAUNCH_CHECK(); ctx->launches++; }
  }
}

template void msm_batch_buckets<Fq, Fp>(Ctx*, const Fp*, long long, const Aff<Fq>*, int, int, int, int, int, const Fp*, int, Xyzz<Fq>*);
template void msm_batch_buckets<Fp, Fq>(Ctx*, const Fq*, long long, const Aff<Fp>*, int, int, int, int, int, const Fq*, int, Xyzz<Fp>*);

}  // namespace tb
