// ORACLE (test infrastructure, never shipped): C ABI over the CPU restatement of the primitives.
// Field id: 0 = Fp (circuit field, Vesta scalar), 1 = Fq.  Curve id: 0 = Vesta (base Fq, scalars Fp),
// 1 = Pallas (base Fp, scalars Fq).  All field elements cross this ABI as 32-byte canonical LE,
// points as 64-byte affine (x||y), identity = zeros.
#include "prims.hpp"

using namespace orc;

template <class Sc> static int field_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
  Sc x = Sc::from_bytes(a), y = b ? Sc::from_bytes(b) : Sc::zero(), r;
  switch (op) {
    case 0: r = x + y; break;
    case 1: r = x - y; break;
    case 2: r = x * y; break;
    case 3: r = x.inv(); break;
    case 4: if (!x.sqrt(r)) return 1; break;
    case 5: r = x.neg(); break;
    default: return -1;
  }
  r.to_bytes(out); return 0;
}

template <class Sc> static void field_consts(int k, uint8_t* out) {
  Sc::omega(k).to_bytes(out); Sc::delta().to_bytes(out + 32); Sc::zeta().to_bytes(out + 64);
  memcpy(out + 96, Sc::C().r, 32); memcpy(out + 128, Sc::C().r2, 32);
  uint64_t inv = Sc::C().inv; memset(out + 160, 0, 32); memcpy(out + 160, &inv, 8);
}

template <class Sc> static std::vector<Sc> load_vec(const uint8_t* d, size_t n) {
  std::vector<Sc> v(n);
  parallel_for(n, [&](size_t s, size_t e) { for (size_t i = s; i < e; ++i) v[i] = Sc::from_bytes(d + 32 * i); });
  return v;
}
template <class Sc> static void store_vec(const std::vector<Sc>& v, uint8_t* d) {
  parallel_for(v.size(), [&](size_t s, size_t e) { for (size_t i = s; i < e; ++i) v[i].to_bytes(d + 32 * i); });
}

template <class Sc> static int ntt_api(int logn, int inverse, uint8_t* data) {
  size_t n = size_t(1) << logn;
  auto v = load_vec<Sc>(data, n);
  Sc w = Sc::omega(logn);
  if (inverse) { fft(v.data(), logn, w.inv()); Sc ni = Sc::from_u64(n).inv(); for (auto& x : v) x = x * ni; }
  else fft(v.data(), logn, w);
  store_vec(v, data); return 0;
}

template <class F, class Sc> static int msm_api(size_t n, const uint8_t* scalars, const uint8_t* points, uint8_t* out) {
  std::vector<u64[4]> can(n);
  std::vector<Affine<F>> bases(n);
  parallel_for(n, [&](size_t s, size_t e) {
    for (size_t i = s; i < e; ++i) { memcpy(can[i], scalars + 32 * i, 32); bases[i] = affine_from_bytes<F>(points + 64 * i); }
  });
  affine_to_bytes(best_multiexp<F>(can.data(), bases.data(), n).to_affine(), out);
  return 0;
}

template <class F> static int decompress_api(size_t n, const uint8_t* in, uint8_t* out) {
  int bad = 0;
  parallel_for(n, [&](size_t s, size_t e) {
    for (size_t i = s; i < e; ++i) { Affine<F> a; if (!decompress<F>(in + 32 * i, a)) { bad = 1; continue; } affine_to_bytes(a, out + 64 * i); }
  });
  return bad;
}
template <class F> static int compress_api(size_t n, const uint8_t* in, uint8_t* out) {
  for (size_t i = 0; i < n; ++i) compress(affine_from_bytes<F>(in + 64 * i), out + 32 * i);
  return 0;
}
template <class F> static int point_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
  Affine<F> pa = affine_from_bytes<F>(a);
  Jac<F> r;
  if (op == 0) r = Jac<F>::from_affine(pa).add_affine(affine_from_bytes<F>(b));
  else if (op == 1) { u64 k[4]; memcpy(k, b, 32); r = Jac<F>::from_affine(pa).mul(k); }
  else if (op == 2) r = Jac<F>::from_affine(pa).dbl();
  else return -1;
  affine_to_bytes(r.to_affine(), out); return 0;
}

extern "C" {
int orc_set_threads(int n) { int old = num_threads(); if (n > 0) num_threads() = n; return old; }
int orc_field_op(int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
  return field == 0 ? field_op<Fp>(op, a, b, out) : field_op<Fq>(op, a, b, out);
}
int orc_field_consts(int field, int k, uint8_t* out) { if (field == 0) field_consts<Fp>(k, out); else field_consts<Fq>(k, out); return 0; }
int orc_from_uniform(int field, const uint8_t* in64, uint8_t* out) {
  if (field == 0) Fp::from_uniform(in64).to_bytes(out); else Fq::from_uniform(in64).to_bytes(out); return 0;
}
int orc_ntt(int field, int logn, int inverse, uint8_t* data) { return field == 0 ? ntt_api<Fp>(logn, inverse, data) : ntt_api<Fq>(logn, inverse, data); }
int orc_coeff_to_extended(int k, int cs_degree, const uint8_t* coeffs, uint8_t* out) {
  Domain<Fp> d(cs_degree, k);
  store_vec(d.coeff_to_extended(load_vec<Fp>(coeffs, d.n)), out); return d.ext_k;
}
int orc_extended_to_coeff(int k, int cs_degree, const uint8_t* evals, uint8_t* out) {
  Domain<Fp> d(cs_degree, k);
  store_vec(d.extended_to_coeff(load_vec<Fp>(evals, d.ext_n)), out); return (int)d.quotient_poly_degree;
}
int orc_msm(int curve, size_t n, const uint8_t* scalars, const uint8_t* points, uint8_t* out64) {
  return curve == 0 ? msm_api<Fq, Fp>(n, scalars, points, out64) : msm_api<Fp, Fq>(n, scalars, points, out64);
}
int orc_decompress(int curve, size_t n, const uint8_t* in32, uint8_t* out64) { return curve == 0 ? decompress_api<Fq>(n, in32, out64) : decompress_api<Fp>(n, in32, out64); }
int orc_compress(int curve, size_t n, const uint8_t* in64, uint8_t* out32) { return curve == 0 ? compress_api<Fq>(n, in64, out32) : compress_api<Fp>(n, in64, out32); }
int orc_point_op(int curve, int op, const uint8_t* a, const uint8_t* b, uint8_t* out) { return curve == 0 ? point_op<Fq>(op, a, b, out) : point_op<Fp>(op, a, b, out); }
}
