// ORACLE (test infrastructure, never shipped or linked into the product library).
// CPU restatement of the Pasta field / curve arithmetic that Taiga's prover path uses through the
// un-vendored crate pasta_curves 0.5.1 (heliaxdev fork, branch `taiga`;
// /root/reference/taiga_halo2/Cargo.toml:10, /root/reference/Cargo.toml:10).  4x64-bit Montgomery limbs,
// the representation Rust's `Fp([u64;4])` holds.  Constants are derived at start-up from the moduli
//  and cross-checked against oracle/pasta.py in tests/test_oracle.py.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {
typedef unsigned __int128 u128;
typedef uint64_t u64;

struct FpTag { static constexpr u64 M[4] = {0x992d30ed00000001ULL, 0x224698fc094cf91bULL, 0, 0x4000000000000000ULL}; };
struct FqTag { static constexpr u64 M[4] = {0x8c46eb2100000001ULL, 0x224698fc0994a8ddULL, 0, 0x4000000000000000ULL}; };

template <class T>
struct Fe {
  u64 l[4];  // Montgomery form, always fully reduced (< modulus)

  static inline bool geq_mod(const u64* a) {
    for (int i = 3; i >= 0; --i) {
      if (a[i] > T::M[i]) return true;
      if (a[i] < T::M[i]) return false;
    }
    return true;
  }
  static inline void sub_mod(u64* a) {
    u128 br = 0;
    for (int i = 0; i < 4; ++i) {
      u128 d = (u128)a[i] - T::M[i] - (u64)br;
      a[i] = (u64)d;
      br = (d >> 64) & 1;
    }
  }
  struct Consts {
    u64 inv;       // -m^-1 mod 2^64
    u64 r[4];      // 2^256 mod m
    u64 r2[4];     // 2^512 mod m
    Consts() {
      u64 x = 1;
      for (int i = 0; i < 6; ++i) x *= 2 - T::M[0] * x;  // Newton: m^-1 mod 2^64
      inv = (u64)0 - x;
      u64 v[4] = {1, 0, 0, 0};
      for (int i = 0; i < 512; ++i) {
        // v = 2v mod m  (v < m < 2^255 so no overflow)
        u64 c = 0;
        for (int j = 0; j < 4; ++j) { u64 n = (v[j] << 1) | c; c = v[j] >> 63; v[j] = n; }
        if (geq_mod(v)) sub_mod(v);
        if (i == 255) memcpy(r, v, 32);
      }
      memcpy(r2, v, 32);
    }
  };
  static const Consts& C() { static Consts c; return c; }

  static Fe zero() { Fe z; memset(z.l, 0, 32); return z; }
  static Fe one() { Fe o; memcpy(o.l, C().r, 32); return o; }
  static Fe from_u64(u64 v) { u64 c[4] = {v, 0, 0, 0}; return from_canonical(c); }
  static Fe from_canonical(const u64* c) {  // c must be < m
    Fe a; memcpy(a.l, c, 32);
    Fe r2; memcpy(r2.l, C().r2, 32);
    return a * r2;
  }
  static bool canonical_ok(const uint8_t* b) { u64 c[4]; memcpy(c, b, 32); return !geq_mod(c); }
  static Fe from_bytes(const uint8_t* b) { u64 c[4]; memcpy(c, b, 32); return from_canonical(c); }
  // 64 little-endian bytes reduced mod m (pasta `from_uniform_bytes`, used by halo2 Challenge255)
  static Fe from_uniform(const uint8_t* b) {
    // Horner over 8-bit digits from the top: slow but obviously right
    Fe acc = zero();
    Fe base = from_u64(256);
    for (int i = 63; i >= 0; --i) acc = acc * base + from_u64(b[i]);
    return acc;
  }
  void to_canonical(u64* out) const {
    Fe one_raw; one_raw.l[0] = 1; one_raw.l[1] = one_raw.l[2] = one_raw.l[3] = 0;
    Fe t = (*this) * one_raw;
    memcpy(out, t.l, 32);
  }
  void to_bytes(uint8_t* b) const { u64 c[4]; to_canonical(c); memcpy(b, c, 32); }

  bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
  bool operator==(const Fe& o) const { return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3]; }
  bool operator!=(const Fe& o) const { return !(*this == o); }

  Fe operator+(const Fe& o) const {
    Fe r; u128 c = 0;
    for (int i = 0; i < 4; ++i) { c += (u128)l[i] + o.l[i]; r.l[i] = (u64)c; c >>= 64; }
    if (geq_mod(r.l)) sub_mod(r.l);
    return r;
  }
  Fe operator-(const Fe& o) const {
    Fe r; u128 br = 0;
    for (int i = 0; i < 4; ++i) { u128 d = (u128)l[i] - o.l[i] - (u64)br; r.l[i] = (u64)d; br = (d >> 64) & 1; }
    if (br) { u128 c = 0; for (int i = 0; i < 4; ++i) { c += (u128)r.l[i] + T::M[i]; r.l[i] = (u64)c; c >>= 64; } }
    return r;
  }
  Fe neg() const { return zero() - *this; }
  Fe dbl() const { return *this + *this; }
  // CIOS Montgomery multiplication
  Fe operator*(const Fe& o) const {
    u64 t[6] = {0, 0, 0, 0, 0, 0};
    const u64 inv = C().inv;
    for (int i = 0; i < 4; ++i) {
      u128 c = 0;
      for (int j = 0; j < 4; ++j) { c += (u128)l[j] * o.l[i] + t[j]; t[j] = (u64)c; c >>= 64; }
      c += t[4]; t[4] = (u64)c; t[5] = (u64)(c >> 64);
      u64 m = t[0] * inv;
      c = (u128)m * T::M[0] + t[0]; c >>= 64;
      for (int j = 1; j < 4; ++j) { c += (u128)m * T::M[j] + t[j]; t[j - 1] = (u64)c; c >>= 64; }
      c += t[4]; t[3] = (u64)c; t[4] = t[5] + (u64)(c >> 64);
    }
    Fe r; memcpy(r.l, t, 32);
    if (t[4] || geq_mod(r.l)) sub_mod(r.l);
    return r;
  }
  Fe sqr() const { return (*this) * (*this); }
  Fe pow(const u64* e, int nlimbs) const {
    Fe acc = one();
    for (int i = nlimbs - 1; i >= 0; --i)
      for (int b = 63; b >= 0; --b) { acc = acc.sqr(); if ((e[i] >> b) & 1) acc = acc * (*this); }
    return acc;
  }
  Fe pow_u64(u64 e) const { return pow(&e, 1); }
  Fe inv() const {  // 0 -> 0 (halo2 batch_invert skips zeros)
    u64 e[4]; memcpy(e, T::M, 32); e[0] -= 2;
    return pow(e, 4);
  }
  // canonical-integer ordering (pasta `impl Ord for Fp`: compares to_repr from the top byte)
  static int cmp(const Fe& a, const Fe& b) {
    u64 x[4], y[4]; a.to_canonical(x); b.to_canonical(y);
    for (int i = 3; i >= 0; --i) { if (x[i] < y[i]) return -1; if (x[i] > y[i]) return 1; }
    return 0;
  }
  // field constants (pasta_curves fields/fp.rs, fq.rs: GENERATOR=5, S=32, ROOT_OF_UNITY, DELTA, ZETA)
  static Fe root_of_unity() {  // 5^((m-1)/2^32)
    u64 e[4]; memcpy(e, T::M, 32);
    // (m-1) >> 32
    e[0] -= 1;
    for (int i = 0; i < 4; ++i) e[i] = (e[i] >> 32) | (i < 3 ? (e[i + 1] << 32) : 0);
    return from_u64(5).pow(e, 4);
  }
  static Fe delta() { u64 e[1] = {1ULL << 32}; return from_u64(5).pow(e, 1); }
  static Fe zeta();
  static Fe omega(int k) { Fe w = root_of_unity(); for (int i = k; i < 32; ++i) w = w.sqr(); return w; }
  // Tonelli-Shanks (2-adicity 32); returns false if non-residue
  bool sqrt(Fe& out) const {
    if (is_zero()) { out = zero(); return true; }
    u64 q[4]; memcpy(q, T::M, 32); q[0] -= 1;
    for (int i = 0; i < 4; ++i) q[i] = (q[i] >> 32) | (i < 3 ? (q[i + 1] << 32) : 0);  // odd part
    u64 qp1h[4];  // (q+1)/2
    { u128 c = 1; for (int i = 0; i < 4; ++i) { c += q[i]; qp1h[i] = (u64)c; c >>= 64; }
      for (int i = 0; i < 4; ++i) qp1h[i] = (qp1h[i] >> 1) | (i < 3 ? (qp1h[i + 1] << 63) : 0); }
    Fe c = root_of_unity();
    Fe t = pow(q, 4), r = pow(qp1h, 4);
    int m = 32;
    while (t != one()) {
      int i = 0; Fe t2 = t;
      while (t2 != one()) { t2 = t2.sqr(); ++i; if (i == m) return false; }
      Fe b = c; for (int j = 0; j < m - i - 1; ++j) b = b.sqr();
      m = i; c = b.sqr(); t = t * c; r = r * b;
    }
    out = r; return true;
  }
};

typedef Fe<FpTag> Fp;
typedef Fe<FqTag> Fq;

template <> inline Fp Fp::zeta() {
  static const u64 z[4] = {0x1dad5ebdfdfe4ab9ULL, 0x1d1f8bd237ad3149ULL, 0x2caad5dc57aab1b0ULL, 0x12ccca834acdba71ULL};
  return from_canonical(z);
}
template <> inline Fq Fq::zeta() {
  static const u64 z[4] = {0x2aa9d2e050aa0e4fULL, 0x0fed467d47c033afULL, 0x511db4d81cf70f5aULL, 0x06819a58283e528eULL};
  return from_canonical(z);
}

// Montgomery-trick batch inversion; zeros stay zero (halo2 `BatchInvert` semantics)
template <class F>
inline void batch_invert(F* v, size_t n) {
  std::vector<F> pre(n);
  F acc = F::one();
  for (size_t i = 0; i < n; ++i) { pre[i] = acc; if (!v[i].is_zero()) acc = acc * v[i]; }
  acc = acc.inv();
  for (size_t i = n; i-- > 0;) {
    if (v[i].is_zero()) continue;
    F t = acc * pre[i]; acc = acc * v[i]; v[i] = t;
  }
}

// ------------------------------------------------------------------ curves y^2 = x^3 + 5 (a = 0)
template <class F> struct Affine { F x, y; bool inf; };

template <class F>
struct Jac {
  F X, Y, Z;  // identity: Z == 0
  static Jac identity() { Jac j; j.X = F::zero(); j.Y = F::one(); j.Z = F::zero(); return j; }
  static Jac from_affine(const Affine<F>& a) { if (a.inf) return identity(); Jac j; j.X = a.x; j.Y = a.y; j.Z = F::one(); return j; }
  bool is_identity() const { return Z.is_zero(); }
  Jac dbl() const {
    if (is_identity()) return *this;
    F A = X.sqr(), B = Y.sqr(), Cc = B.sqr();
    F D = ((X + B).sqr() - A - Cc).dbl();
    F E = A.dbl() + A, Fv = E.sqr();
    Jac r;
    r.X = Fv - D.dbl();
    r.Y = E * (D - r.X) - Cc.dbl().dbl().dbl();
    r.Z = (Y * Z).dbl();
    return r;
  }
  Jac add_affine(const Affine<F>& b) const {
    if (b.inf) return *this;
    if (is_identity()) return from_affine(b);
    F Z2 = Z.sqr(), U2 = b.x * Z2, S2 = b.y * Z * Z2;
    F H = U2 - X, R = S2 - Y;
    if (H.is_zero()) { if (R.is_zero()) return dbl(); return identity(); }
    F H2 = H.sqr(), H3 = H * H2, V = X * H2;
    Jac r;
    r.X = R.sqr() - H3 - V.dbl();
    r.Y = R * (V - r.X) - Y * H3;
    r.Z = Z * H;
    return r;
  }
  Jac add(const Jac& b) const {
    if (b.is_identity()) return *this;
    if (is_identity()) return b;
    F Z1Z1 = Z.sqr(), Z2Z2 = b.Z.sqr();
    F U1 = X * Z2Z2, U2 = b.X * Z1Z1, S1 = Y * b.Z * Z2Z2, S2 = b.Y * Z * Z1Z1;
    F H = U2 - U1, R = S2 - S1;
    if (H.is_zero()) { if (R.is_zero()) return dbl(); return identity(); }
    F H2 = H.sqr(), H3 = H * H2, V = U1 * H2;
    Jac r;
    r.X = R.sqr() - H3 - V.dbl();
    r.Y = R * (V - r.X) - S1 * H3;
    r.Z = Z * b.Z * H;
    return r;
  }
  Jac neg() const { Jac r = *this; r.Y = Y.neg(); return r; }
  Affine<F> to_affine() const {
    Affine<F> a;
    if (is_identity()) { a.x = F::zero(); a.y = F::zero(); a.inf = true; return a; }
    F zi = Z.inv(), zi2 = zi.sqr();
    a.x = X * zi2; a.y = Y * zi2 * zi; a.inf = false;
    return a;
  }
  // scalar given as canonical 4x64 limbs
  Jac mul(const u64* k) const {
    Jac acc = identity();
    for (int i = 3; i >= 0; --i)
      for (int b = 63; b >= 0; --b) { acc = acc.dbl(); if ((k[i] >> b) & 1) acc = acc.add(*this); }
    return acc;
  }
};

// affine <-> 64 byte (x||y canonical LE, identity = all zero)  [vesta::Affine identity is (0,0), EXT]
template <class F> inline Affine<F> affine_from_bytes(const uint8_t* b) {
  Affine<F> a; a.x = F::from_bytes(b); a.y = F::from_bytes(b + 32); a.inf = a.x.is_zero() && a.y.is_zero(); return a;
}
template <class F> inline void affine_to_bytes(const Affine<F>& a, uint8_t* b) {
  if (a.inf) { memset(b, 0, 64); return; }
  a.x.to_bytes(b); a.y.to_bytes(b + 32);
}
// pasta compressed encoding: x LE with bit 255 = parity of y; identity = zeros
template <class F> inline bool decompress(const uint8_t* b, Affine<F>& out) {
  uint8_t t[32]; memcpy(t, b, 32);
  int sign = t[31] >> 7; t[31] &= 0x7f;
  bool allz = true; for (int i = 0; i < 32; ++i) allz &= (t[i] == 0);
  if (allz && !sign) { out.inf = true; out.x = F::zero(); out.y = F::zero(); return true; }
  if (!F::canonical_ok(t)) return false;
  F x = F::from_bytes(t);
  F rhs = x.sqr() * x + F::from_u64(5), y;
  if (!rhs.sqrt(y)) return false;
  uint8_t yb[32]; y.to_bytes(yb);
  if ((yb[0] & 1) != sign) y = y.neg();
  out.x = x; out.y = y; out.inf = false; return true;
}
template <class F> inline void compress(const Affine<F>& a, uint8_t* b) {
  if (a.inf) { memset(b, 0, 32); return; }
  uint8_t yb[32]; a.x.to_bytes(b); a.y.to_bytes(yb);
  b[31] |= (yb[0] & 1) << 7;
}
}  // namespace orc
