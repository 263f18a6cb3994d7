// Minimal C++ caller of the host-side mirror (include/taiga_b200.hpp): a one-gate PLONKish circuit
//     q * (a * b - c) = 0,  public c            (k = 5, 3 advice + 1 instance + 1 selector, no lookups,
// a, b, c, instance in the permutation with the identity wiring except c[0] <-> instance[0]), proved and verified with
// Proof::create / Proof::verify as a Taiga caller would (proof.rs:25-54).  Field arithmetic for the key material comes
// from the library's own header in host mode.  Without a B200 it must fail loudly: there is no CPU fallback (that path and
// the build are what tests/test_abi.py checks; the proving path itself is covered through the same C ABI by tests/test_gpu_*.py).
//
//   g++ -std=c++17 -I include examples/prove_cpp.cpp -L taiga_b200 -ltaiga_b200 -Wl,-rpath,$PWD/taiga_b200 -o /tmp/prove_cpp
#include <cstdio>
#include <vector>

#define TB_PORTABLE_FIELD 1
#include "../taiga_b200/csrc/field.cuh"
#include "taiga_b200.hpp"

using namespace taiga_b200;
using tb::Fp;

static FieldBytes to_bytes(const Fp& v) {
  Fp c = v.from_mont();
  FieldBytes b;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) b[4 * i + j] = (uint8_t)(c.l[i] >> (8 * j));
  return b;
}
static Fp small(uint32_t v) { return Fp::from_u32(v); }

int main() {
  try {
    const uint32_t k = 5, n = 1u << k;
    Context ctx(0);   // throws without a usable sm_100 device

    // ---- a throw-away SRS: g_i = [s_i] G is not computable on the host without curve code, so this example uses the
    // generator for every basis point.  (Binding is irrelevant for a smoke run; real callers pass params_15.)
    std::vector<uint8_t> g(64 * n), gl(64 * n);
    PointBytes G{};   // Vesta generator (-1, 2): x = q - 1, y = 2
    { const uint32_t qm1[8] = {0x00000000u, 0x8c46eb21u, 0x0994a8ddu, 0x224698fcu, 0, 0, 0, 0x40000000u};
      for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) G[4 * i + j] = (uint8_t)(qm1[i] >> (8 * j));
      G[32] = 2; }
    for (uint32_t i = 0; i < n; ++i) { std::memcpy(&g[64 * i], G.data(), 64); std::memcpy(&gl[64 * i], G.data(), 64); }
    Params params(ctx, k, g.data(), gl.data(), G, G);

    // ---- constraint system
    ConstraintSystem cs;
    cs.k = k; cs.num_advice = 3; cs.num_fixed = 1; cs.num_instance = 1;
    cs.cs_degree = 3; cs.blinding_factors = 5;   // max(3, 1 query per column) + 2
    for (uint32_t c = 0; c < 3; ++c) cs.advice_queries.push_back(tb_query{c, 0});
    cs.fixed_queries.push_back(tb_query{0, 0});
    cs.instance_queries.push_back(tb_query{0, 0});
    for (uint32_t c = 0; c < 3; ++c) cs.perm_columns.push_back(tb_column{TB_COL_ADVICE, c});
    cs.perm_columns.push_back(tb_column{TB_COL_INSTANCE, 0});
    uint32_t a = cs.add_node(TB_EX_ADVICE, 0), b = cs.add_node(TB_EX_ADVICE, 1), c = cs.add_node(TB_EX_ADVICE, 2), q = cs.add_node(TB_EX_FIXED, 0);
    uint32_t ab = cs.add_node(TB_EX_MUL, a, b), nc = cs.add_node(TB_EX_NEG, c), d = cs.add_node(TB_EX_ADD, ab, nc);
    cs.constraint_roots.push_back(cs.add_node(TB_EX_MUL, q, d));

    // ---- key material: selector on row 0; sigma = identity (delta^col * omega^row) with c[0] <-> instance[0] swapped
    std::vector<uint8_t> fixed(32 * n, 0), sigma(32 * n * 4);
    { FieldBytes one = to_bytes(Fp::one()); std::memcpy(&fixed[0], one.data(), 32); }
    Fp omega = Fp::one();
    { // omega_k = ROOT_OF_UNITY^(2^(32-k)), ROOT_OF_UNITY = 5^((p-1)/2^32);  delta = 5^(2^32)
      Fp root = Fp::one();
      // (p - 1) / 2^32 as 7 limbs: p = 2^254 + 0x224698fc094cf91b992d30ed00000001
      const uint32_t t[8] = {0x992d30edu, 0x094cf91bu, 0x224698fcu, 0, 0, 0, 0x40000000u, 0};
      Fp base = small(5);
      for (int limb = 7; limb >= 0; --limb) for (int bit = 31; bit >= 0; --bit) { root = root.sqr(); if ((t[limb] >> bit) & 1) root = root * base; }
      omega = root;
      for (uint32_t i = 0; i < 32 - k; ++i) omega = omega.sqr();
    }
    Fp delta = small(5);
    for (int i = 0; i < 32; ++i) delta = delta.sqr();
    std::vector<Fp> sig(4 * n);
    Fp dc = Fp::one();
    for (uint32_t col = 0; col < 4; ++col) { Fp w = dc; for (uint32_t r = 0; r < n; ++r) { sig[col * n + r] = w; w = w * omega; } dc = dc * delta; }
    std::swap(sig[2 * n + 0], sig[3 * n + 0]);
    for (size_t i = 0; i < sig.size(); ++i) { FieldBytes v = to_bytes(sig[i]); std::memcpy(&sigma[32 * i], v.data(), 32); }
    ProvingKey pk(params, cs, fixed.data(), sigma.data());

    // ---- witness: 3 * 7 = 21 on row 0
    std::vector<uint8_t> advice(32 * n * 3, 0);
    FieldBytes three = to_bytes(small(3)), seven = to_bytes(small(7)), twentyone = to_bytes(small(21));
    std::memcpy(&advice[0], three.data(), 32); std::memcpy(&advice[32 * n], seven.data(), 32); std::memcpy(&advice[64 * n], twentyone.data(), 32);
    std::vector<std::vector<FieldBytes>> instance = {{twentyone}};
    std::array<uint8_t, 32> seed{};   // a real caller draws this from its RNG (proof.rs:30)

    Proof proof = Proof::create(pk, params, AdviceTable{advice.data()}, instance, seed);
    proof.verify(pk, params, instance);
    std::printf("proof of %zu bytes created and verified; kernels launched: %llu\n", proof.inner().size(), (unsigned long long)ctx.launch_count());
    instance[0][0] = seven;
    try { proof.verify(pk, params, instance); std::printf("ERROR: wrong instance accepted\n"); return 2; }
    catch (const Error& e) { std::printf("wrong instance rejected: %s (%s)\n", e.what(), e.kind()); }
    return 0;
  } catch (const Error& e) {
    std::fprintf(stderr, "taiga_b200 error [%s, status %d]: %s\n", e.kind(), e.status(), e.what());
    return 1;
  }
}
