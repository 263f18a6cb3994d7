// Host instantiation of the product's __host__ __device__ field/curve arithmetic (taiga_b200/csrc/*.cuh),
// so the same source the kernels compile can be checked on a CPU-only box against the oracle.
#include <cstring>
#include "../taiga_b200/csrc/curve.cuh"
using namespace tb;
template <class F> static void fop(int op, const uint8_t* a, const uint8_t* b, uint8_t* o) {
  F x, y, r; memcpy(&x, a, 32); memcpy(&y, b, 32);
  x = x.to_mont(); y = y.to_mont();
  switch (op) { case 0: r = x + y; break; case 1: r = x - y; break; case 2: r = x * y; break; case 3: r = x.inv(); break; case 4: r = x.neg(); break; default: r = x.sqr(); }
  r = r.from_mont(); memcpy(o, &r, 32);
}
template <class F> static Aff<F> ld(const uint8_t* p) { Aff<F> a; memcpy(&a, p, 64); a.x = a.x.to_mont(); a.y = a.y.to_mont(); return a; }
template <class F> static void st(const Aff<F>& a, uint8_t* p) { Aff<F> b; b.x = a.x.from_mont(); b.y = a.y.from_mont(); memcpy(p, &b, 64); }
template <class F> static void pop(int op, const uint8_t* a, const uint8_t* b, uint8_t* o) {
  Aff<F> pa = ld<F>(a); Xyzz<F> r;
  if (op == 0) { r = Xyzz<F>::from_affine(pa); r.add_affine(ld<F>(b)); }
  else if (op == 1) { uint32_t k[8]; memcpy(k, b, 32); r = scalar_mul(pa, k); }
  else if (op == 2) { r = Xyzz<F>::from_affine(pa).dbl(); }
  else if (op == 3) { r = Xyzz<F>::from_affine(pa).dbl(); Xyzz<F> s = Xyzz<F>::from_affine(ld<F>(b)).dbl(); s.add_affine(ld<F>(b)); r.add(s); }  // 2a + 3b
  else { r = Xyzz<F>::dbl_affine(pa); }
  st(r.to_affine(), o);
}
extern "C" {
void hs_field(int f, int op, const uint8_t* a, const uint8_t* b, uint8_t* o) { if (f == 0) fop<Fp>(op, a, b, o); else fop<Fq>(op, a, b, o); }
void hs_point(int c, int op, const uint8_t* a, const uint8_t* b, uint8_t* o) { if (c == 0) pop<Fq>(op, a, b, o); else pop<Fp>(op, a, b, o); }
}
