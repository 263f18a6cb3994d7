// Shared host-side plumbing for libtaiga_b200: context, error handling, stream-ordered device memory.
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <set>
#include <vector>
#include "curve.cuh"

namespace tb {

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
// the witness does not satisfy the circuit (halo2 Error::ConstraintSystemFailure class)
struct ConstraintError : std::runtime_error { using std::runtime_error::runtime_error; };

#define TB_CUDA(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess)                                                                              \
      throw tb::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)
#define TB_LAUNCH_CHECK() TB_CUDA(cudaGetLastError())
#define TB_REQUIRE(cond, msg) do { if (!(cond)) throw std::invalid_argument(std::string(msg) + " (" #cond ")"); } while (0)

// NTT twiddle tables: powers of the 2^24-th root of unity, two-level (tables are 2 x 128 KiB per
// field and direction, L2 resident).  w_S^e = hi[e >> 12] * lo[e & 4095].
constexpr int TW_LOG = 24;
constexpr int TW_HALF = 12;
// w_S^e, S = 2^24: two-level tables (e = hi * 2^12 + lo) plus, for the circuit field, a flat table of the 2^full_log-th roots --
// every twiddle the k <= 19 prover NTTs and the quotient kernels ask for is then one 32-byte load instead of a load pair and a multiply
template <class F> struct TwiddleTables { F* lo = nullptr; F* hi = nullptr; F* full = nullptr; int full_log = 0; };
constexpr int TW_FULL_LOG = 19;

template <class F> struct FieldTables {
  TwiddleTables<F> fwd, inv;
};

// kernel categories for the built-in CUDA-event profiler (bench.py's roofline / share-of-step numbers)
enum ProfCat { PC_NTT = 0, PC_MSM_SORT, PC_MSM_ACCUM, PC_MSM_REDUCE, PC_QUOT_GATES, PC_QUOT_FINISH, PC_IPA_FOLD, PC_TRANSCRIPT, PC_LOOKUP_SORT,
               PC_POLY, PC_COUNT };
struct ProfRec { int cat; cudaEvent_t a, b; };

// experiment knob: integer from the environment (read on every call; only used on host set-up paths)
inline int tb_tune(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string last_error;
  FieldTables<Fp> tw_fp;
  FieldTables<Fq> tw_fq;
  int sm_count = 148;
  uint64_t launches = 0;  // kernels launched through this context (bench's gpu_launches)
  bool prof = false; std::vector<ProfRec> prof_recs; std::vector<cudaEvent_t> event_pool;
  // executed 255-bit Montgomery multiplications per kernel category (host-side accounting at every launch, for the integer-pipe
  // roofline of bench.py); the bucket additions of the batched MSM are counted on the device (they depend on the scalars)
  double work[PC_COUNT] = {0};
  unsigned long long* d_msm_adds = nullptr;
  cudaEvent_t get_event() {
    cudaEvent_t e;
    if (!event_pool.empty()) { e = event_pool.back(); event_pool.pop_back(); return e; }
    TB_CUDA(cudaEventCreate(&e)); return e;
  }

  template <class T> T* alloc(size_t count) {
    void* p = nullptr;
    if (count == 0) count = 1;
    TB_CUDA(cudaMallocAsync(&p, count * sizeof(T), stream));
    return reinterpret_cast<T*>(p);
  }
  void free(void* p) { if (p) cudaFreeAsync(p, stream); }
  // Opt a kernel into more than 48 KB of dynamic shared memory.  The attribute belongs to the (function, device) pair, so the
  // "already done" set lives in the context (= one device), not in a process-wide static.
  std::set<const void*> smem_opted;
  template <class K> void opt_in_smem(K kernel, size_t bytes) {
    if (smem_opted.insert(reinterpret_cast<const void*>(kernel)).second)
      TB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  }
  void sync() { TB_CUDA(cudaStreamSynchronize(stream)); }
};

// records a pair of CUDA events on the context's stream around a group of launches (only when profiling is on)
struct ProfScope {
  Ctx* c; int idx = -1;
  ProfScope(Ctx* ctx, int cat) : c(ctx) {
    if (!c->prof) return;
    ProfRec r; r.cat = cat; r.a = c->get_event(); r.b = c->get_event();
    cudaEventRecord(r.a, c->stream);
    idx = (int)c->prof_recs.size(); c->prof_recs.push_back(r);
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(c->prof_recs[idx].b, c->stream); }
};

// RAII stream-ordered device buffer
template <class T> struct DevBuf {
  Ctx* ctx = nullptr; T* p = nullptr; size_t n = 0;
  DevBuf() {}
  DevBuf(Ctx* c, size_t count) : ctx(c), n(count) { p = c->alloc<T>(count); }
  DevBuf(const DevBuf&) = delete; DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : ctx(o.ctx), p(o.p), n(o.n) { o.p = nullptr; }
  DevBuf& operator=(DevBuf&& o) noexcept { if (this != &o) { release(); ctx = o.ctx; p = o.p; n = o.n; o.p = nullptr; } return *this; }
  ~DevBuf() { release(); }
  void release() { if (p && ctx) ctx->free(p); p = nullptr; }
  T* get() const { return p; }
  void zero() { TB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), ctx->stream)); }
  void upload(const void* host, size_t count) { TB_CUDA(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream)); }
  void download(void* host, size_t count) const { TB_CUDA(cudaMemcpyAsync(host, p, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream)); }
};

template <class F> inline FieldTables<F>& field_tables(Ctx* c);
template <> inline FieldTables<Fp>& field_tables<Fp>(Ctx* c) { return c->tw_fp; }
template <> inline FieldTables<Fq>& field_tables<Fq>(Ctx* c) { return c->tw_fq; }

// 2^32-th roots of unity (pasta_curves ROOT_OF_UNITY), canonical 32-bit LE limbs
template <class F> TB_HD F root_of_unity_2_32();
template <> TB_HD Fp root_of_unity_2_32<Fp>() {
  Fp r; const uint32_t v[8] = {0xd87ea32fu, 0xbdad6fabu, 0xb7bb7584u, 0xea322bf2u, 0x0561f81au, 0x36212083u, 0xac30ebdau, 0x2bce74deu};
  for (int i = 0; i < 8; ++i) r.l[i] = v[i]; return r.to_mont();
}
template <> TB_HD Fq root_of_unity_2_32<Fq>() {
  Fq r; const uint32_t v[8] = {0x02b6d05fu, 0xa70e2c11u, 0xc106f049u, 0x9bb97ea3u, 0x492ae26eu, 0x9e5c4dfdu, 0x746d3f58u, 0x2de6a9b8u};
  for (int i = 0; i < 8; ++i) r.l[i] = v[i]; return r.to_mont();
}
template <class F> TB_HD F zeta_const();  // pasta_curves ZETA (cube root of unity; halo2 coset shift)
template <> TB_HD Fp zeta_const<Fp>() {
  Fp r; const uint32_t v[8] = {0xfdfe4ab9u, 0x1dad5ebdu, 0x37ad3149u, 0x1d1f8bd2u, 0x57aab1b0u, 0x2caad5dcu, 0x4acdba71u, 0x12ccca83u};
  for (int i = 0; i < 8; ++i) r.l[i] = v[i]; return r.to_mont();
}
template <> TB_HD Fq zeta_const<Fq>() {
  Fq r; const uint32_t v[8] = {0x50aa0e4fu, 0x2aa9d2e0u, 0x47c033afu, 0x0fed467du, 0x1cf70f5au, 0x511db4d8u, 0x283e528eu, 0x06819a58u};
  for (int i = 0; i < 8; ++i) r.l[i] = v[i]; return r.to_mont();
}
template <class F> inline F omega_k(int k) { F w = root_of_unity_2_32<F>(); for (int i = k; i < 32; ++i) w = w.sqr(); return w; }
template <class F> inline F delta_const() { return F::from_u32(5).pow_u64(1ull << 32); }  // pasta DELTA = 5^(2^32)

#ifdef __CUDACC__
// 128-bit vectorised global access of a field element (two LDG.128 / STG.128)
template <class F> __device__ __forceinline__ F ldg_fe(const F* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  F r; r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w; r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
  return r;
}
template <class F> __device__ __forceinline__ F ld_fe(const F* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  F r; r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w; r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
  return r;
}
template <class F> __device__ __forceinline__ void st_fe(F* p, const F& v) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
  q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
template <class F> __device__ __forceinline__ Aff<F> ldg_aff(const Aff<F>* p) {
  Aff<F> a; a.x = ldg_fe(&p->x); a.y = ldg_fe(&p->y); return a;
}
// w_S^e from the two-level table
template <class F> __device__ __forceinline__ F tw_pow2(const TwiddleTables<F>& t, uint32_t e) {
  F h = ldg_fe(t.hi + (e >> TW_HALF));
  uint32_t lo = e & ((1u << TW_HALF) - 1);
  if (lo) h = h * ldg_fe(t.lo + lo);
  return h;
}
template <class F> __device__ __forceinline__ F tw_pow(const TwiddleTables<F>& t, uint32_t e) {
  if (t.full && (e & ((1u << (TW_LOG - TW_FULL_LOG)) - 1)) == 0) return ldg_fe(t.full + (e >> (TW_LOG - TW_FULL_LOG)));
  return tw_pow2(t, e);
}
#endif

}  // namespace tb
