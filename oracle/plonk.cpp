// ORACLE (test infrastructure, never shipped): CPU restatement of halo2_proofs `plonk::{keygen, create_proof,
// verify_proof}`, `poly::multiopen` and `poly::commitment` (IPA) for one circuit instance over Vesta, as called by
// taiga_halo2/src/proof.rs:25-54.  halo2_proofs is an un-vendored git dependency (heliaxdev/halo2, branch `taiga`,
// taiga_halo2/Cargo.toml:14-15): the algorithm is restated from the published zcash/halo2 0.3 lineage
// (SURVEY.md App. A).  PARITY UNPINNED against the real Rust verifier (no cargo in this image; the reference holds
// no golden proof bytes because every test uses OsRng): the pins are the SRS fixture identities, the in-container
// prover<->verifier round trip incl. tamper rejection, and proof-size accounting.
//
// Blinding randomness: the reference draws from the caller's RNG; here every random scalar is derived from a
// 32-byte seed with a BLAKE2b PRF (rnd()) so that the CUDA prover can be compared byte for byte.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <set>
#include "../include/taiga_b200.h"
#include "blake2b.hpp"
#include "prims.hpp"

namespace orc {

typedef Affine<Fq> Pt;
typedef Jac<Fq> JPt;

enum RndTag { R_ADVICE_ROWS = 1, R_ADVICE_BLIND, R_LK_IN_ROWS, R_LK_TAB_ROWS, R_LK_IN_BLIND, R_LK_TAB_BLIND, R_PERM_ROWS, R_PERM_BLIND,
              R_LKZ_ROWS, R_LKZ_BLIND, R_RANDOM_POLY, R_RANDOM_BLIND, R_H_BLIND, R_QPRIME_BLIND, R_S_POLY, R_S_BLIND, R_IPA_L, R_IPA_R };

static Fp rnd(const uint8_t* seed, uint32_t proof, uint32_t tag, uint32_t idx) {
  Blake2b b(64, "TaigaB200-Blind\0");
  uint8_t msg[48]; memcpy(msg, seed, 32);
  uint32_t w[4] = {proof, tag, idx, 0}; memcpy(msg + 32, w, 16);
  b.update(msg, 48);
  uint8_t out[64]; b.finalize(out);
  return Fp::from_uniform(out);
}

struct Desc {
  uint32_t k, na, nf, ni, degree, bf;
  std::vector<tb_query> aq, fq, iq;
  std::vector<tb_column> perm;
  std::vector<Fp> consts;
  std::vector<tb_expr_node> nodes;
  std::vector<uint32_t> roots;
  struct Lk { std::vector<uint32_t> in, tab; };
  std::vector<Lk> lookups;
  Fp vk_repr;
  explicit Desc(const tb_cs_desc* c) {
    k = c->k; na = c->num_advice; nf = c->num_fixed; ni = c->num_instance; degree = c->cs_degree; bf = c->blinding_factors;
    aq.assign(c->advice_queries, c->advice_queries + c->num_advice_queries);
    fq.assign(c->fixed_queries, c->fixed_queries + c->num_fixed_queries);
    iq.assign(c->instance_queries, c->instance_queries + c->num_instance_queries);
    perm.assign(c->perm_columns, c->perm_columns + c->num_perm_columns);
    for (uint32_t i = 0; i < c->num_constants; ++i) consts.push_back(Fp::from_bytes(c->constants + 32 * i));
    nodes.assign(c->nodes, c->nodes + c->num_nodes);
    roots.assign(c->constraint_roots, c->constraint_roots + c->num_constraints);
    for (uint32_t i = 0; i < c->num_lookups; ++i) {
      Lk l; l.in.assign(c->lookups[i].input_roots, c->lookups[i].input_roots + c->lookups[i].num_exprs);
      l.tab.assign(c->lookups[i].table_roots, c->lookups[i].table_roots + c->lookups[i].num_exprs);
      lookups.push_back(l);
    }
    vk_repr = Fp::from_bytes(c->vk_transcript_repr);
  }
  int query_index(const std::vector<tb_query>& qs, uint32_t col, int rot) const {
    for (size_t i = 0; i < qs.size(); ++i) if (qs[i].column == col && qs[i].rotation == rot) return (int)i;
    return -1;
  }
  // halo2 Expression::evaluate over per-query values
  template <class GA, class GF, class GI>
  void eval_nodes(std::vector<Fp>& v, GA ga, GF gf, GI gi) const {
    v.resize(nodes.size());
    for (size_t i = 0; i < nodes.size(); ++i) {
      const tb_expr_node& nd = nodes[i];
      switch (nd.op) {
        case TB_EX_CONST: v[i] = consts[nd.a]; break;
        case TB_EX_ADVICE: v[i] = ga(nd.a); break;
        case TB_EX_FIXED: v[i] = gf(nd.a); break;
        case TB_EX_INSTANCE: v[i] = gi(nd.a); break;
        case TB_EX_NEG: v[i] = v[nd.a].neg(); break;
        case TB_EX_ADD: v[i] = v[nd.a] + v[nd.b]; break;
        case TB_EX_MUL: v[i] = v[nd.a] * v[nd.b]; break;
        case TB_EX_SCALE: v[i] = v[nd.a] * consts[nd.b]; break;
        default: v[i] = Fp::zero();
      }
    }
  }
};

// ---- transcript: Blake2bWrite / Blake2bRead with Challenge255
struct Transcript {
  Blake2b st; std::vector<uint8_t> proof; const uint8_t* rd = nullptr; size_t rd_len = 0, rd_pos = 0; bool bad = false;
  Transcript() : st(64, "Halo2-Transcript") {}
  void common_point(const Pt& p) {
    if (p.inf) { bad = true; return; }  // "cannot write points at infinity to the transcript"
    uint8_t b[65]; b[0] = 1; p.x.to_bytes(b + 1); p.y.to_bytes(b + 33); st.update(b, 65);
  }
  void common_scalar(const Fp& s) { uint8_t b[33]; b[0] = 2; s.to_bytes(b + 1); st.update(b, 33); }
  void write_point(const Pt& p) { common_point(p); uint8_t b[32]; compress(p, b); proof.insert(proof.end(), b, b + 32); }
  void write_scalar(const Fp& s) { common_scalar(s); uint8_t b[32]; s.to_bytes(b); proof.insert(proof.end(), b, b + 32); }
  Fp squeeze() { uint8_t z = 0; st.update(&z, 1); uint8_t out[64]; st.finalize(out); return Fp::from_uniform(out); }
  bool read_point(Pt& p) {
    if (rd_pos + 32 > rd_len) { bad = true; return false; }
    if (!decompress<Fq>(rd + rd_pos, p)) { bad = true; return false; }
    rd_pos += 32; common_point(p); return !bad;
  }
  bool read_scalar(Fp& s) {
    if (rd_pos + 32 > rd_len || !Fp::canonical_ok(rd + rd_pos)) { bad = true; return false; }
    s = Fp::from_bytes(rd + rd_pos); rd_pos += 32; common_scalar(s); return true;
  }
};

// optional phase timers (ORC_TIMING=1): where a CPU proof spends its time; test infrastructure only
struct PhaseTimer {
  static std::map<std::string, double>& table() { static std::map<std::string, double> t; return t; }
  static bool on() { static bool f = getenv("ORC_TIMING") != nullptr; return f; }
  const char* name; std::chrono::steady_clock::time_point t0;
  explicit PhaseTimer(const char* n) : name(n), t0(std::chrono::steady_clock::now()) {}
  ~PhaseTimer() { if (on()) table()[name] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
  static void report() {
    if (!on()) return;
    double tot = 0; for (auto& kv : table()) tot += kv.second;
    for (auto& kv : table()) fprintf(stderr, "[orc] %-22s %8.3f s %5.1f %%\n", kv.first.c_str(), kv.second, 100 * kv.second / tot);
    table().clear();
  }
};
struct PhaseSeq {   // consecutive phases of prove(): next("x") closes the previous one
  PhaseTimer* cur = nullptr;
  void next(const char* n) { delete cur; cur = n ? new PhaseTimer(n) : nullptr; }
  ~PhaseSeq() { delete cur; }
};

struct Key {
  Desc d; Domain<Fp> dom; size_t n; int chunk_len; size_t nsets;
  std::vector<Pt> g, gl; Pt w, u;
  std::vector<std::vector<Fp>> fixed_vals, fixed_polys, fixed_cosets, sig_vals, sig_polys, sig_cosets;
  std::vector<Fp> l0, l_last, l_blind;  // extended cosets
  std::vector<Pt> fixed_comms, sig_comms;
  Key(const tb_cs_desc* c) : d(c), dom(c->cs_degree, c->k), n(size_t(1) << c->k) {}

  JPt commit(const std::vector<Pt>& bases, const std::vector<Fp>& v, const Fp& blind) const {
    PhaseTimer pt("commit (n-term MSM)");
    JPt r = msm<Fq, Fp>(v.data(), bases.data(), v.size());
    u64 b[4]; blind.to_canonical(b);
    return r.add(JPt::from_affine(w).mul(b));
  }
};

static std::vector<Fp> lagrange_basis_coset(const Key& key, const std::vector<size_t>& rows) {
  std::vector<Fp> v(key.n, Fp::zero());
  for (size_t r : rows) v[r] = Fp::one();
  key.dom.lagrange_to_coeff(v);
  return key.dom.coeff_to_extended(v);
}

static Key* keygen(const tb_cs_desc* c, const uint8_t* g, const uint8_t* gl, const uint8_t* w, const uint8_t* u, const uint8_t* fixed,
                   const uint8_t* sigma) {
  Key* key = new Key(c);
  size_t n = key->n;
  key->g.resize(n); key->gl.resize(n);
  for (size_t i = 0; i < n; ++i) { key->g[i] = affine_from_bytes<Fq>(g + 64 * i); key->gl[i] = affine_from_bytes<Fq>(gl + 64 * i); }
  key->w = affine_from_bytes<Fq>(w); key->u = affine_from_bytes<Fq>(u);
  key->chunk_len = (int)key->d.degree - 2;
  key->nsets = key->d.perm.empty() ? 0 : (key->d.perm.size() + key->chunk_len - 1) / key->chunk_len;
  auto load_cols = [&](const uint8_t* src, size_t ncols, std::vector<std::vector<Fp>>& vals, std::vector<std::vector<Fp>>& polys,
                       std::vector<std::vector<Fp>>& cosets, std::vector<Pt>& comms) {
    for (size_t cidx = 0; cidx < ncols; ++cidx) {
      std::vector<Fp> v(n);
      for (size_t i = 0; i < n; ++i) v[i] = Fp::from_bytes(src + 32 * (cidx * n + i));
      comms.push_back(key->commit(key->gl, v, Fp::one()).to_affine());  // keygen_vk: commit_lagrange(col, Blind::default())
      vals.push_back(v);
      key->dom.lagrange_to_coeff(v);
      polys.push_back(v);
      cosets.push_back(key->dom.coeff_to_extended(v));
    }
  };
  load_cols(fixed, key->d.nf, key->fixed_vals, key->fixed_polys, key->fixed_cosets, key->fixed_comms);
  load_cols(sigma, key->d.perm.size(), key->sig_vals, key->sig_polys, key->sig_cosets, key->sig_comms);
  size_t bf = key->d.bf;
  key->l0 = lagrange_basis_coset(*key, {0});
  key->l_last = lagrange_basis_coset(*key, {n - bf - 1});
  std::vector<size_t> blind_rows; for (size_t r = n - bf; r < n; ++r) blind_rows.push_back(r);
  key->l_blind = lagrange_basis_coset(*key, blind_rows);
  return key;
}

// halo2 lookup::prover::permute_expression_pair
static bool permute_pair(const Key& key, const std::vector<Fp>& input, const std::vector<Fp>& table, std::vector<Fp>& pin, std::vector<Fp>& ptab) {
  size_t usable = key.n - (key.d.bf + 1);
  pin.assign(input.begin(), input.begin() + usable);
  std::sort(pin.begin(), pin.end(), [](const Fp& a, const Fp& b) { return Fp::cmp(a, b) < 0; });
  struct Less { bool operator()(const Fp& a, const Fp& b) const { return Fp::cmp(a, b) < 0; } };
  std::map<Fp, uint32_t, Less> leftover;
  for (size_t i = 0; i < usable; ++i) leftover[table[i]]++;
  ptab.assign(usable, Fp::zero());
  std::vector<size_t> repeated;
  for (size_t row = 0; row < usable; ++row) {
    if (row == 0 || pin[row] != pin[row - 1]) {
      ptab[row] = pin[row];
      auto it = leftover.find(pin[row]);
      if (it == leftover.end() || it->second == 0) return false;  // Error::ConstraintSystemFailure
      it->second--;
    } else repeated.push_back(row);
  }
  for (auto& kv : leftover)
    for (uint32_t c = 0; c < kv.second; ++c) { ptab[repeated.back()] = kv.first; repeated.pop_back(); }
  return repeated.empty();
}

struct PolyRef { const std::vector<Fp>* poly; Fp blind; int rot; };  // ProverQuery (point = x * omega^rot)

// poly/commitment/prover.rs create_proof (inner product argument)
static void ipa_prove(const Key& key, Transcript& tr, const uint8_t* seed, uint32_t pidx, const std::vector<Fp>& p_poly, const Fp& p_blind, const Fp& x3) {
  size_t n = key.n; uint32_t k = key.d.k;
  std::vector<Fp> s_poly(n);
  for (size_t i = 0; i < n; ++i) s_poly[i] = rnd(seed, pidx, R_S_POLY, (uint32_t)i);
  Fp s_at = eval_polynomial(s_poly.data(), n, x3);
  s_poly[0] = s_poly[0] - s_at;
  Fp s_blind = rnd(seed, pidx, R_S_BLIND, 0);
  tr.write_point(key.commit(key.g, s_poly, s_blind).to_affine());
  Fp xi = tr.squeeze(), z = tr.squeeze();
  std::vector<Fp> p(n);
  for (size_t i = 0; i < n; ++i) p[i] = s_poly[i] * xi + p_poly[i];
  Fp v = eval_polynomial(p.data(), n, x3);
  p[0] = p[0] - v;
  Fp f = s_blind * xi + p_blind;
  std::vector<Fp> b(n); { Fp cur = Fp::one(); for (size_t i = 0; i < n; ++i) { b[i] = cur; cur = cur * x3; } }
  std::vector<Pt> gp = key.g;
  for (uint32_t j = 0; j < k; ++j) {
    size_t half = size_t(1) << (k - j - 1);
    JPt l = msm<Fq, Fp>(p.data() + half, gp.data(), half), r = msm<Fq, Fp>(p.data(), gp.data() + half, half);
    Fp vl = Fp::zero(), vr = Fp::zero();
    for (size_t i = 0; i < half; ++i) { vl = vl + p[half + i] * b[i]; vr = vr + p[i] * b[half + i]; }
    Fp lr = rnd(seed, pidx, R_IPA_L, j), rr = rnd(seed, pidx, R_IPA_R, j);
    u64 c[4];
    (vl * z).to_canonical(c); l = l.add(JPt::from_affine(key.u).mul(c)); lr.to_canonical(c); l = l.add(JPt::from_affine(key.w).mul(c));
    (vr * z).to_canonical(c); r = r.add(JPt::from_affine(key.u).mul(c)); rr.to_canonical(c); r = r.add(JPt::from_affine(key.w).mul(c));
    tr.write_point(l.to_affine()); tr.write_point(r.to_affine());
    Fp uj = tr.squeeze(), uj_inv = uj.inv();
    u64 uc[4]; uj.to_canonical(uc);
    parallel_for(half, [&](size_t s, size_t e) {
      for (size_t i = s; i < e; ++i) gp[i] = JPt::from_affine(gp[i]).add(JPt::from_affine(gp[i + half]).mul(uc)).to_affine();
    });
    for (size_t i = 0; i < half; ++i) { p[i] = p[i] + p[i + half] * uj_inv; b[i] = b[i] + b[i + half] * uj; }
    f = f + lr * uj_inv + rr * uj;
  }
  tr.write_scalar(p[0]);
  tr.write_scalar(f);
}

// groups queries by polynomial / point set (halo2 multiopen::construct_intermediate_sets); points identified by rotation
struct Sets { std::vector<std::vector<int>> point_sets; std::vector<int> poly_set; std::vector<int> uniq; };
template <class Q> static Sets intermediate_sets(const std::vector<Q>& qs, const std::vector<const void*>& ids) {
  Sets s;
  std::vector<const void*> uniq_ids; std::vector<std::set<int>> rots; std::map<int, int> point_index;
  for (size_t i = 0; i < qs.size(); ++i) {
    if (!point_index.count(qs[i].rot)) { int idx = (int)point_index.size(); point_index[qs[i].rot] = idx; }
    size_t pos = 0; for (; pos < uniq_ids.size(); ++pos) if (uniq_ids[pos] == ids[i]) break;
    if (pos == uniq_ids.size()) { uniq_ids.push_back(ids[i]); rots.emplace_back(); s.uniq.push_back((int)i); }
    rots[pos].insert(point_index[qs[i].rot]);
  }
  std::map<int, int> inv_point; for (auto& kv : point_index) inv_point[kv.second] = kv.first;
  std::map<std::set<int>, int> set_index;
  for (size_t c = 0; c < uniq_ids.size(); ++c) {
    if (!set_index.count(rots[c])) { int idx = (int)set_index.size(); set_index[rots[c]] = idx; }
    s.poly_set.push_back(set_index[rots[c]]);
  }
  s.point_sets.resize(set_index.size());
  for (auto& kv : set_index) for (int pi : kv.first) s.point_sets[kv.second].push_back(inv_point[pi]);  // rotations, ordered by point index
  return s;
}

static int prove(const Key& key, const uint8_t* advice_bytes, const uint8_t* instance_bytes, const uint32_t* instance_len, const uint8_t* seed,
                 uint32_t pidx, std::vector<uint8_t>& out) {
  const Desc& d = key.d; const Domain<Fp>& dom = key.dom;
  size_t n = key.n, bf = d.bf, usable_start = n - (bf + 1), ext_n = dom.ext_n;
  int ext_shift = dom.ext_k - dom.k; size_t rot_scale = size_t(1) << ext_shift;
  Transcript tr;
  tr.common_scalar(d.vk_repr);
  PhaseSeq ph;
  ph.next("1 instance+advice (incl. commits, ffts)");
  // ---- instance columns
  std::vector<std::vector<Fp>> inst_vals(d.ni), inst_polys(d.ni), inst_cosets(d.ni);
  { size_t off = 0;
    for (uint32_t c = 0; c < d.ni; ++c) {
      if (instance_len[c] > usable_start) return 2;  // Error::InstanceTooLarge
      inst_vals[c].assign(n, Fp::zero());
      for (uint32_t i = 0; i < instance_len[c]; ++i) inst_vals[c][i] = Fp::from_bytes(instance_bytes + 32 * (off + i));
      off += instance_len[c];
      tr.common_point(key.commit(key.gl, inst_vals[c], Fp::one()).to_affine());
      inst_polys[c] = inst_vals[c]; dom.lagrange_to_coeff(inst_polys[c]);
      inst_cosets[c] = dom.coeff_to_extended(inst_polys[c]);
    } }
  // ---- advice columns
  std::vector<std::vector<Fp>> adv_vals(d.na), adv_polys(d.na), adv_cosets(d.na); std::vector<Fp> adv_blinds(d.na);
  for (uint32_t c = 0; c < d.na; ++c) {
    adv_vals[c].resize(n);
    for (size_t i = 0; i < n; ++i) adv_vals[c][i] = Fp::from_bytes(advice_bytes + 32 * (c * n + i));
    for (size_t r = 0; r <= bf; ++r) adv_vals[c][usable_start + r] = rnd(seed, pidx, R_ADVICE_ROWS, (uint32_t)(c * (bf + 1) + r));
  }
  for (uint32_t c = 0; c < d.na; ++c) adv_blinds[c] = rnd(seed, pidx, R_ADVICE_BLIND, c);
  for (uint32_t c = 0; c < d.na; ++c) tr.write_point(key.commit(key.gl, adv_vals[c], adv_blinds[c]).to_affine());
  for (uint32_t c = 0; c < d.na; ++c) { adv_polys[c] = adv_vals[c]; dom.lagrange_to_coeff(adv_polys[c]); adv_cosets[c] = dom.coeff_to_extended(adv_polys[c]); }
  Fp theta = tr.squeeze();
  ph.next("2 lookups permute (incl.)");
  // ---- lookups: compress, permute, commit A', S'
  size_t nl = d.lookups.size();
  std::vector<std::vector<Fp>> lk_in(nl), lk_tab(nl), lk_pin(nl), lk_ptab(nl), lk_pin_poly(nl), lk_ptab_poly(nl), lk_pin_coset(nl), lk_ptab_coset(nl);
  std::vector<Fp> lk_pin_blind(nl), lk_ptab_blind(nl);
  for (size_t l = 0; l < nl; ++l) {
    lk_in[l].assign(n, Fp::zero()); lk_tab[l].assign(n, Fp::zero());
    parallel_for(n, [&](size_t s, size_t e) {
      std::vector<Fp> v;
      for (size_t i = s; i < e; ++i) {
        auto rowq = [&](const std::vector<std::vector<Fp>>& cols, const tb_query& q) { return cols[q.column][(i + n + (int64_t)q.rotation) % n]; };
        d.eval_nodes(v, [&](uint32_t q) { return rowq(adv_vals, d.aq[q]); }, [&](uint32_t q) { return rowq(key.fixed_vals, d.fq[q]); },
                     [&](uint32_t q) { return rowq(inst_vals, d.iq[q]); });
        Fp a = Fp::zero(), t = Fp::zero();
        for (uint32_t r : d.lookups[l].in) a = a * theta + v[r];
        for (uint32_t r : d.lookups[l].tab) t = t * theta + v[r];
        lk_in[l][i] = a; lk_tab[l][i] = t;
      }
    });
    if (!permute_pair(key, lk_in[l], lk_tab[l], lk_pin[l], lk_ptab[l])) return 3;  // ConstraintSystemFailure
    for (size_t r = 0; r <= bf; ++r) lk_pin[l].push_back(rnd(seed, pidx, R_LK_IN_ROWS, (uint32_t)(l * (bf + 1) + r)));
    for (size_t r = 0; r <= bf; ++r) lk_ptab[l].push_back(rnd(seed, pidx, R_LK_TAB_ROWS, (uint32_t)(l * (bf + 1) + r)));
    lk_pin_blind[l] = rnd(seed, pidx, R_LK_IN_BLIND, (uint32_t)l); lk_ptab_blind[l] = rnd(seed, pidx, R_LK_TAB_BLIND, (uint32_t)l);
    tr.write_point(key.commit(key.gl, lk_pin[l], lk_pin_blind[l]).to_affine());
    tr.write_point(key.commit(key.gl, lk_ptab[l], lk_ptab_blind[l]).to_affine());
    lk_pin_poly[l] = lk_pin[l]; dom.lagrange_to_coeff(lk_pin_poly[l]); lk_pin_coset[l] = dom.coeff_to_extended(lk_pin_poly[l]);
    lk_ptab_poly[l] = lk_ptab[l]; dom.lagrange_to_coeff(lk_ptab_poly[l]); lk_ptab_coset[l] = dom.coeff_to_extended(lk_ptab_poly[l]);
  }
  Fp beta = tr.squeeze(), gamma = tr.squeeze();
  ph.next("3 perm products (incl.)");
  // ---- permutation argument: grand products
  auto col_vals = [&](const tb_column& c) -> const std::vector<Fp>& {
    return c.kind == TB_COL_ADVICE ? adv_vals[c.index] : c.kind == TB_COL_FIXED ? key.fixed_vals[c.index] : inst_vals[c.index]; };
  auto col_coset = [&](const tb_column& c) -> const std::vector<Fp>& {
    return c.kind == TB_COL_ADVICE ? adv_cosets[c.index] : c.kind == TB_COL_FIXED ? key.fixed_cosets[c.index] : inst_cosets[c.index]; };
  size_t nsets = key.nsets, chunk = key.chunk_len, P = d.perm.size();
  std::vector<std::vector<Fp>> pz_poly(nsets), pz_coset(nsets); std::vector<Fp> pz_blind(nsets);
  { Fp deltaomega = Fp::one(), last_z = Fp::one(), delta = Fp::delta();
    for (size_t s = 0; s < nsets; ++s) {
      size_t c0 = s * chunk, c1 = std::min(P, c0 + chunk);
      std::vector<Fp> mv(n, Fp::one());
      for (size_t c = c0; c < c1; ++c) { const auto& vals = col_vals(d.perm[c]); for (size_t i = 0; i < n; ++i) mv[i] = mv[i] * (beta * key.sig_vals[c][i] + gamma + vals[i]); }
      batch_invert(mv.data(), n);
      for (size_t c = c0; c < c1; ++c) {
        const auto& vals = col_vals(d.perm[c]); Fp dw = deltaomega;
        for (size_t i = 0; i < n; ++i) { mv[i] = mv[i] * (dw * beta + gamma + vals[i]); dw = dw * dom.omega; }
        deltaomega = deltaomega * delta;
      }
      std::vector<Fp> z(n); z[0] = last_z;
      for (size_t i = 1; i < n; ++i) z[i] = z[i - 1] * mv[i - 1];
      for (size_t r = 0; r < bf; ++r) z[n - bf + r] = rnd(seed, pidx, R_PERM_ROWS, (uint32_t)(s * bf + r));
      last_z = z[n - (bf + 1)];
      pz_blind[s] = rnd(seed, pidx, R_PERM_BLIND, (uint32_t)s);
      tr.write_point(key.commit(key.gl, z, pz_blind[s]).to_affine());
      dom.lagrange_to_coeff(z); pz_poly[s] = z; pz_coset[s] = dom.coeff_to_extended(z);
    } }
  ph.next("4 lookup products (incl.)");
  // ---- lookup grand products
  std::vector<std::vector<Fp>> lz_poly(nl), lz_coset(nl); std::vector<Fp> lz_blind(nl);
  for (size_t l = 0; l < nl; ++l) {
    std::vector<Fp> lp(n);
    for (size_t i = 0; i < n; ++i) lp[i] = (beta + lk_pin[l][i]) * (gamma + lk_ptab[l][i]);
    batch_invert(lp.data(), n);
    for (size_t i = 0; i < n; ++i) lp[i] = lp[i] * (lk_in[l][i] + beta) * (lk_tab[l][i] + gamma);
    std::vector<Fp> z(n); Fp st = Fp::one(); z[0] = st;
    for (size_t i = 1; i < n - bf; ++i) { st = st * lp[i - 1]; z[i] = st; }
    for (size_t r = 0; r < bf; ++r) z[n - bf + r] = rnd(seed, pidx, R_LKZ_ROWS, (uint32_t)(l * bf + r));
    lz_blind[l] = rnd(seed, pidx, R_LKZ_BLIND, (uint32_t)l);
    tr.write_point(key.commit(key.gl, z, lz_blind[l]).to_affine());
    dom.lagrange_to_coeff(z); lz_poly[l] = z; lz_coset[l] = dom.coeff_to_extended(z);
  }
  ph.next("5 random poly (incl.)");
  // ---- vanishing argument: random polynomial
  std::vector<Fp> random_poly(n);
  for (size_t i = 0; i < n; ++i) random_poly[i] = rnd(seed, pidx, R_RANDOM_POLY, (uint32_t)i);
  Fp random_blind = rnd(seed, pidx, R_RANDOM_BLIND, 0);
  tr.write_point(key.commit(key.g, random_poly, random_blind).to_affine());
  Fp y = tr.squeeze();
  ph.next("6 quotient eval");
  // ---- quotient h(X) on the extended domain
  std::vector<Fp> h(ext_n);
  { Fp delta = Fp::delta(); int last_rot = -(int)(bf + 1);
    std::vector<Fp> delta_start(nsets); for (size_t s = 0; s < nsets; ++s) delta_start[s] = beta * delta.pow_u64(s * chunk);
    parallel_for(ext_n, [&](size_t s0, size_t e0) {
      std::vector<Fp> v; Fp one = Fp::one();
      Fp xcur = dom.zeta * dom.ext_omega.pow_u64(s0);
      for (size_t i = s0; i < e0; ++i, xcur = xcur * dom.ext_omega) {
        auto at = [&](const std::vector<Fp>& col, int rot) -> const Fp& { return col[(i + ext_n + (int64_t)rot * (int64_t)rot_scale) % ext_n]; };
        d.eval_nodes(v, [&](uint32_t q) { return at(adv_cosets[d.aq[q].column], d.aq[q].rotation); },
                     [&](uint32_t q) { return at(key.fixed_cosets[d.fq[q].column], d.fq[q].rotation); },
                     [&](uint32_t q) { return at(inst_cosets[d.iq[q].column], d.iq[q].rotation); });
        Fp acc = Fp::zero();
        for (uint32_t r : d.roots) acc = acc * y + v[r];
        Fp l0 = key.l0[i], ll = key.l_last[i], active = one - (ll + key.l_blind[i]);
        if (nsets) {
          acc = acc * y + l0 * (one - pz_coset[0][i]);
          const Fp& zl = pz_coset[nsets - 1][i];
          acc = acc * y + ll * (zl * zl - zl);
          for (size_t s = 1; s < nsets; ++s) acc = acc * y + l0 * (pz_coset[s][i] - at(pz_coset[s - 1], last_rot));
          for (size_t s = 0; s < nsets; ++s) {
            size_t c0 = s * chunk, c1 = std::min(P, c0 + chunk);
            Fp left = at(pz_coset[s], 1), right = pz_coset[s][i], cd = delta_start[s] * xcur;
            for (size_t c = c0; c < c1; ++c) {
              const Fp& val = col_coset(d.perm[c])[i];
              left = left * (val + beta * key.sig_cosets[c][i] + gamma);
              right = right * (val + cd + gamma);
              cd = cd * delta;
            }
            acc = acc * y + (left - right) * active;
          }
        }
        for (size_t l = 0; l < nl; ++l) {
          Fp a = Fp::zero(), t = Fp::zero();
          for (uint32_t r : d.lookups[l].in) a = a * theta + v[r];
          for (uint32_t r : d.lookups[l].tab) t = t * theta + v[r];
          const Fp& z = lz_coset[l][i]; const Fp& ap = lk_pin_coset[l][i]; const Fp& sp = lk_ptab_coset[l][i];
          acc = acc * y + l0 * (one - z);
          acc = acc * y + ll * (z * z - z);
          acc = acc * y + (at(lz_coset[l], 1) * (ap + beta) * (sp + gamma) - z * (a + beta) * (t + gamma)) * active;
          acc = acc * y + l0 * (ap - sp);
          acc = acc * y + (ap - sp) * (ap - at(lk_pin_coset[l], -1)) * active;
        }
        h[i] = acc;
      }
    }); }
  ph.next("7 h to coeff + commit (incl.)");
  dom.divide_by_vanishing_poly(h);
  std::vector<Fp> hc = dom.extended_to_coeff(h);
  size_t npieces = dom.quotient_poly_degree;
  std::vector<std::vector<Fp>> h_pieces(npieces); std::vector<Fp> h_blinds(npieces);
  for (size_t p = 0; p < npieces; ++p) { h_pieces[p].assign(hc.begin() + p * n, hc.begin() + (p + 1) * n); h_blinds[p] = rnd(seed, pidx, R_H_BLIND, (uint32_t)p); }
  for (size_t p = 0; p < npieces; ++p) tr.write_point(key.commit(key.g, h_pieces[p], h_blinds[p]).to_affine());
  Fp x = tr.squeeze();
  Fp xn = x.pow_u64(n);
  ph.next("8 evaluations");
  // ---- evaluations
  auto ev = [&](const std::vector<Fp>& poly, int rot) { return eval_polynomial(poly.data(), poly.size(), dom.rotate_omega(x, rot)); };
  for (auto& q : d.iq) tr.write_scalar(ev(inst_polys[q.column], q.rotation));
  for (auto& q : d.aq) tr.write_scalar(ev(adv_polys[q.column], q.rotation));
  for (auto& q : d.fq) tr.write_scalar(ev(key.fixed_polys[q.column], q.rotation));
  std::vector<Fp> h_poly(n, Fp::zero()); Fp h_blind = Fp::zero();
  for (size_t p = npieces; p-- > 0;) { for (size_t i = 0; i < n; ++i) h_poly[i] = h_poly[i] * xn + h_pieces[p][i]; h_blind = h_blind * xn + h_blinds[p]; }
  tr.write_scalar(ev(random_poly, 0));
  for (size_t c = 0; c < P; ++c) tr.write_scalar(ev(key.sig_polys[c], 0));
  int last_rot = -(int)(bf + 1);
  for (size_t s = 0; s < nsets; ++s) {
    tr.write_scalar(ev(pz_poly[s], 0)); tr.write_scalar(ev(pz_poly[s], 1));
    if (s + 1 < nsets) tr.write_scalar(ev(pz_poly[s], last_rot));
  }
  for (size_t l = 0; l < nl; ++l) {
    tr.write_scalar(ev(lz_poly[l], 0)); tr.write_scalar(ev(lz_poly[l], 1)); tr.write_scalar(ev(lk_pin_poly[l], 0));
    tr.write_scalar(ev(lk_pin_poly[l], -1)); tr.write_scalar(ev(lk_ptab_poly[l], 0));
  }
  // ---- multiopen queries (order of plonk/prover.rs)
  std::vector<PolyRef> qs; Fp one = Fp::one();
  for (auto& q : d.iq) qs.push_back({&inst_polys[q.column], one, q.rotation});
  for (auto& q : d.aq) qs.push_back({&adv_polys[q.column], adv_blinds[q.column], q.rotation});
  for (size_t s = 0; s < nsets; ++s) { qs.push_back({&pz_poly[s], pz_blind[s], 0}); qs.push_back({&pz_poly[s], pz_blind[s], 1}); }
  for (size_t s = nsets; s-- > 0;) if (s + 1 < nsets) qs.push_back({&pz_poly[s], pz_blind[s], last_rot});
  for (size_t l = 0; l < nl; ++l) {
    qs.push_back({&lz_poly[l], lz_blind[l], 0}); qs.push_back({&lk_pin_poly[l], lk_pin_blind[l], 0}); qs.push_back({&lk_ptab_poly[l], lk_ptab_blind[l], 0});
    qs.push_back({&lk_pin_poly[l], lk_pin_blind[l], -1}); qs.push_back({&lz_poly[l], lz_blind[l], 1});
  }
  for (auto& q : d.fq) qs.push_back({&key.fixed_polys[q.column], one, q.rotation});
  for (size_t c = 0; c < P; ++c) qs.push_back({&key.sig_polys[c], one, 0});
  qs.push_back({&h_poly, h_blind, 0});
  qs.push_back({&random_poly, random_blind, 0});
  ph.next("9 multiopen (incl.)");
  // ---- multiopen::create_proof
  Fp x1 = tr.squeeze(), x2 = tr.squeeze();
  std::vector<const void*> ids; for (auto& q : qs) ids.push_back(q.poly);
  Sets sets = intermediate_sets(qs, ids);
  size_t ns = sets.point_sets.size();
  std::vector<std::vector<Fp>> q_polys(ns); std::vector<Fp> q_blinds(ns, Fp::zero());
  for (size_t c = 0; c < sets.uniq.size(); ++c) {
    const PolyRef& pr = qs[sets.uniq[c]]; int si = sets.poly_set[c];
    if (q_polys[si].empty()) q_polys[si] = *pr.poly;
    else for (size_t i = 0; i < n; ++i) q_polys[si][i] = q_polys[si][i] * x1 + (*pr.poly)[i];
    q_blinds[si] = q_blinds[si] * x1 + pr.blind;
  }
  std::vector<Fp> q_prime;
  for (size_t si = 0; si < ns; ++si) {
    std::vector<Fp> poly = q_polys[si];
    for (int rot : sets.point_sets[si]) poly = kate_division(poly, dom.rotate_omega(x, rot));
    poly.resize(n, Fp::zero());
    if (q_prime.empty()) q_prime = poly; else for (size_t i = 0; i < n; ++i) q_prime[i] = q_prime[i] * x2 + poly[i];
  }
  Fp q_prime_blind = rnd(seed, pidx, R_QPRIME_BLIND, 0);
  tr.write_point(key.commit(key.g, q_prime, q_prime_blind).to_affine());
  Fp x3 = tr.squeeze();
  for (size_t si = 0; si < ns; ++si) tr.write_scalar(eval_polynomial(q_polys[si].data(), n, x3));
  Fp x4 = tr.squeeze();
  std::vector<Fp> p_poly = q_prime; Fp p_blind = q_prime_blind;
  for (size_t si = 0; si < ns; ++si) { for (size_t i = 0; i < n; ++i) p_poly[i] = p_poly[i] * x4 + q_polys[si][i]; p_blind = p_blind * x4 + q_blinds[si]; }
  ph.next("A ipa (incl.)");
  ipa_prove(key, tr, seed, pidx, p_poly, p_blind, x3);
  ph.next(nullptr);
  PhaseTimer::report();
  if (tr.bad) return 4;
  out = tr.proof;
  return 0;
}

// ------------------------------------------------------------------ verifier (plonk/verifier.rs, SingleVerifier)
struct VQuery { int comm; int rot; Fp eval; };  // comm: index into the commitment table (-1 = h msm)
static int verify(const Key& key, const uint8_t* instance_bytes, const uint32_t* instance_len, const uint8_t* proof, size_t proof_len) {
  const Desc& d = key.d; const Domain<Fp>& dom = key.dom;
  size_t n = key.n, bf = d.bf, P = d.perm.size(), nsets = key.nsets, chunk = key.chunk_len, nl = d.lookups.size();
  Transcript tr; tr.rd = proof; tr.rd_len = proof_len;
  tr.common_scalar(d.vk_repr);
  std::vector<Pt> comms;  // commitment table
  std::vector<int> inst_c(d.ni), adv_c(d.na);
  { size_t off = 0;
    for (uint32_t c = 0; c < d.ni; ++c) {
      if (instance_len[c] > n - (bf + 1)) return 2;
      std::vector<Fp> v(n, Fp::zero());
      for (uint32_t i = 0; i < instance_len[c]; ++i) { if (!Fp::canonical_ok(instance_bytes + 32 * (off + i))) return 2; v[i] = Fp::from_bytes(instance_bytes + 32 * (off + i)); }
      off += instance_len[c];
      Pt p = key.commit(key.gl, v, Fp::one()).to_affine();
      tr.common_point(p); inst_c[c] = (int)comms.size(); comms.push_back(p);
    } }
  auto rp = [&](int& idx) { Pt p; if (!tr.read_point(p)) return false; idx = (int)comms.size(); comms.push_back(p); return true; };
  for (uint32_t c = 0; c < d.na; ++c) if (!rp(adv_c[c])) return 5;
  Fp theta = tr.squeeze();
  std::vector<int> lk_pin_c(nl), lk_ptab_c(nl), lz_c(nl), pz_c(nsets);
  for (size_t l = 0; l < nl; ++l) { if (!rp(lk_pin_c[l]) || !rp(lk_ptab_c[l])) return 5; }
  Fp beta = tr.squeeze(), gamma = tr.squeeze();
  for (size_t s = 0; s < nsets; ++s) if (!rp(pz_c[s])) return 5;
  for (size_t l = 0; l < nl; ++l) if (!rp(lz_c[l])) return 5;
  int random_c; if (!rp(random_c)) return 5;
  Fp y = tr.squeeze();
  size_t npieces = dom.quotient_poly_degree;
  std::vector<int> h_c(npieces); for (size_t p = 0; p < npieces; ++p) if (!rp(h_c[p])) return 5;
  Fp x = tr.squeeze();
  auto rs = [&](Fp& s) { return tr.read_scalar(s); };
  std::vector<Fp> inst_ev(d.iq.size()), adv_ev(d.aq.size()), fix_ev(d.fq.size()), sig_ev(P);
  for (auto& e : inst_ev) if (!rs(e)) return 5;
  for (auto& e : adv_ev) if (!rs(e)) return 5;
  for (auto& e : fix_ev) if (!rs(e)) return 5;
  Fp random_eval; if (!rs(random_eval)) return 5;
  for (auto& e : sig_ev) if (!rs(e)) return 5;
  std::vector<Fp> pz_ev(nsets), pz_next(nsets), pz_last(nsets);
  for (size_t s = 0; s < nsets; ++s) { if (!rs(pz_ev[s]) || !rs(pz_next[s])) return 5; if (s + 1 < nsets && !rs(pz_last[s])) return 5; }
  std::vector<Fp> lz_ev(nl), lz_next(nl), pin_ev(nl), pin_inv(nl), ptab_ev(nl);
  for (size_t l = 0; l < nl; ++l) if (!rs(lz_ev[l]) || !rs(lz_next[l]) || !rs(pin_ev[l]) || !rs(pin_inv[l]) || !rs(ptab_ev[l])) return 5;
  // ---- expected h(x)
  Fp one = Fp::one(), xn = x.pow_u64(n);
  // l_i_range(x, xn, -(bf+1)..=0): l_i(x) = (x^n - 1)/n * w^i / (x - w^i)
  auto l_at = [&](int rot) { Fp wi = dom.rotate_omega(one, rot); return (xn - one) * dom.n_inv * wi * (x - wi).inv(); };
  Fp l_last = l_at(-(int)(bf + 1)), l_blind = Fp::zero(), l_0 = l_at(0);
  for (int r = -(int)bf; r <= -1; ++r) l_blind = l_blind + l_at(r);
  std::vector<Fp> v;
  d.eval_nodes(v, [&](uint32_t q) { return adv_ev[q]; }, [&](uint32_t q) { return fix_ev[q]; }, [&](uint32_t q) { return inst_ev[q]; });
  Fp acc = Fp::zero();
  for (uint32_t r : d.roots) acc = acc * y + v[r];
  Fp active = one - (l_last + l_blind), delta = Fp::delta();
  auto col_eval = [&](const tb_column& c) -> Fp {
    int qi = d.query_index(c.kind == TB_COL_ADVICE ? d.aq : c.kind == TB_COL_FIXED ? d.fq : d.iq, c.index, 0);
    if (qi < 0) return Fp::zero();
    return c.kind == TB_COL_ADVICE ? adv_ev[qi] : c.kind == TB_COL_FIXED ? fix_ev[qi] : inst_ev[qi]; };
  if (nsets) {
    acc = acc * y + l_0 * (one - pz_ev[0]);
    acc = acc * y + (pz_ev[nsets - 1].sqr() - pz_ev[nsets - 1]) * l_last;
    for (size_t s = 1; s < nsets; ++s) acc = acc * y + (pz_ev[s] - pz_last[s - 1]) * l_0;
    for (size_t s = 0; s < nsets; ++s) {
      size_t c0 = s * chunk, c1 = std::min(P, c0 + chunk);
      Fp left = pz_next[s], right = pz_ev[s], cd = beta * x * delta.pow_u64(s * chunk);
      for (size_t c = c0; c < c1; ++c) { Fp e = col_eval(d.perm[c]); left = left * (e + beta * sig_ev[c] + gamma); right = right * (e + cd + gamma); cd = cd * delta; }
      acc = acc * y + (left - right) * active;
    }
  }
  for (size_t l = 0; l < nl; ++l) {
    Fp a = Fp::zero(), t = Fp::zero();
    for (uint32_t r : d.lookups[l].in) a = a * theta + v[r];
    for (uint32_t r : d.lookups[l].tab) t = t * theta + v[r];
    acc = acc * y + l_0 * (one - lz_ev[l]);
    acc = acc * y + l_last * (lz_ev[l].sqr() - lz_ev[l]);
    acc = acc * y + (lz_next[l] * (pin_ev[l] + beta) * (ptab_ev[l] + gamma) - lz_ev[l] * (a + beta) * (t + gamma)) * active;
    acc = acc * y + l_0 * (pin_ev[l] - ptab_ev[l]);
    acc = acc * y + (pin_ev[l] - ptab_ev[l]) * (pin_ev[l] - pin_inv[l]) * active;
  }
  Fp expected_h = acc * (xn - one).inv();
  // h commitment = sum_i xn^i * h_i
  JPt hj = JPt::identity(); { u64 c[4]; xn.to_canonical(c); for (size_t p = npieces; p-- > 0;) hj = hj.mul(c).add_affine(comms[h_c[p]]); }
  int hmsm_c = (int)comms.size(); comms.push_back(hj.to_affine());
  std::vector<int> fix_c(d.nf), sig_c(P);
  for (uint32_t c = 0; c < d.nf; ++c) { fix_c[c] = (int)comms.size(); comms.push_back(key.fixed_comms[c]); }
  for (size_t c = 0; c < P; ++c) { sig_c[c] = (int)comms.size(); comms.push_back(key.sig_comms[c]); }
  // ---- queries (same order as the prover)
  std::vector<VQuery> qs; int last_rot = -(int)(bf + 1);
  for (size_t i = 0; i < d.iq.size(); ++i) qs.push_back({inst_c[d.iq[i].column], d.iq[i].rotation, inst_ev[i]});
  for (size_t i = 0; i < d.aq.size(); ++i) qs.push_back({adv_c[d.aq[i].column], d.aq[i].rotation, adv_ev[i]});
  for (size_t s = 0; s < nsets; ++s) { qs.push_back({pz_c[s], 0, pz_ev[s]}); qs.push_back({pz_c[s], 1, pz_next[s]}); }
  for (size_t s = nsets; s-- > 0;) if (s + 1 < nsets) qs.push_back({pz_c[s], last_rot, pz_last[s]});
  for (size_t l = 0; l < nl; ++l) {
    qs.push_back({lz_c[l], 0, lz_ev[l]}); qs.push_back({lk_pin_c[l], 0, pin_ev[l]}); qs.push_back({lk_ptab_c[l], 0, ptab_ev[l]});
    qs.push_back({lk_pin_c[l], -1, pin_inv[l]}); qs.push_back({lz_c[l], 1, lz_next[l]});
  }
  for (size_t i = 0; i < d.fq.size(); ++i) qs.push_back({fix_c[d.fq[i].column], d.fq[i].rotation, fix_ev[i]});
  for (size_t c = 0; c < P; ++c) qs.push_back({sig_c[c], 0, sig_ev[c]});
  qs.push_back({hmsm_c, 0, expected_h});
  qs.push_back({random_c, 0, random_eval});
  // ---- multiopen::verify_proof
  Fp x1 = tr.squeeze(), x2 = tr.squeeze();
  std::vector<const void*> ids; for (auto& q : qs) ids.push_back((const void*)(uintptr_t)(q.comm + 1));
  Sets sets = intermediate_sets(qs, ids);
  size_t ns = sets.point_sets.size();
  std::vector<JPt> q_comm(ns, JPt::identity()); std::vector<std::vector<Fp>> q_evals(ns);
  for (size_t si = 0; si < ns; ++si) q_evals[si].assign(sets.point_sets[si].size(), Fp::zero());
  u64 x1c[4]; x1.to_canonical(x1c);
  for (size_t c = 0; c < sets.uniq.size(); ++c) {
    int comm = qs[sets.uniq[c]].comm, si = sets.poly_set[c];
    q_comm[si] = q_comm[si].mul(x1c).add_affine(comms[comm]);
    for (size_t pi = 0; pi < sets.point_sets[si].size(); ++pi) {
      Fp e = Fp::zero(); bool found = false;
      for (auto& q : qs) if (q.comm == comm && q.rot == sets.point_sets[si][pi]) { e = q.eval; found = true; }
      if (!found) return 6;
      q_evals[si][pi] = q_evals[si][pi] * x1 + e;
    }
  }
  Pt q_prime; if (!tr.read_point(q_prime)) return 5;
  Fp x3 = tr.squeeze();
  std::vector<Fp> u(ns); for (auto& e : u) if (!rs(e)) return 5;
  Fp msm_eval = Fp::zero();
  for (size_t si = 0; si < ns; ++si) {
    size_t m = sets.point_sets[si].size();
    std::vector<Fp> pts(m); for (size_t i = 0; i < m; ++i) pts[i] = dom.rotate_omega(x, sets.point_sets[si][i]);
    // r(x3) by Lagrange interpolation through (pts, q_evals)
    Fp r_eval = Fp::zero();
    for (size_t i = 0; i < m; ++i) { Fp num = one, den = one; for (size_t j = 0; j < m; ++j) if (j != i) { num = num * (x3 - pts[j]); den = den * (pts[i] - pts[j]); } r_eval = r_eval + q_evals[si][i] * num * den.inv(); }
    Fp e = u[si] - r_eval;
    for (size_t i = 0; i < m; ++i) e = e * (x3 - pts[i]).inv();
    msm_eval = msm_eval * x2 + e;
  }
  Fp x4 = tr.squeeze();
  JPt pm = JPt::from_affine(q_prime); Fp vv = msm_eval; u64 x4c[4]; x4.to_canonical(x4c);
  for (size_t si = 0; si < ns; ++si) { pm = pm.mul(x4c).add(q_comm[si]); vv = vv * x4 + u[si]; }
  // ---- commitment::verify_proof (IPA)
  uint32_t k = d.k;
  Pt s_comm; if (!tr.read_point(s_comm)) return 5;
  Fp xi = tr.squeeze(), z = tr.squeeze();
  u64 c4[4];
  xi.to_canonical(c4); JPt lhs = pm.add(JPt::from_affine(s_comm).mul(c4));
  vv.neg().to_canonical(c4); lhs = lhs.add(JPt::from_affine(key.g[0]).mul(c4));  // - [v] G_0
  std::vector<Fp> us(k);
  std::vector<Pt> Ls(k), Rs(k);
  for (uint32_t j = 0; j < k; ++j) { if (!tr.read_point(Ls[j]) || !tr.read_point(Rs[j])) return 5; us[j] = tr.squeeze(); }
  for (uint32_t j = 0; j < k; ++j) {
    us[j].inv().to_canonical(c4); lhs = lhs.add(JPt::from_affine(Ls[j]).mul(c4));
    us[j].to_canonical(c4); lhs = lhs.add(JPt::from_affine(Rs[j]).mul(c4));
  }
  Fp cc, ff; if (!rs(cc) || !rs(ff)) return 5;
  if (tr.rd_pos != proof_len) return 7;  // trailing bytes
  Fp b = one; { Fp cur = x3; for (uint32_t j = k; j-- > 0;) { b = b * (one + us[j] * cur); cur = cur * cur; } }
  std::vector<Fp> s(n, Fp::zero()); s[0] = one;
  { size_t len = 1; for (uint32_t j = k; j-- > 0;) { for (size_t i = 0; i < len; ++i) s[len + i] = s[i] * us[j]; len <<= 1; } }
  for (auto& e : s) e = e * cc;
  // [c] G'_0 + [c b z] U + [f] W  with G'_0 = <s, g>
  JPt rhs = msm<Fq, Fp>(s.data(), key.g.data(), n);
  (cc * b * z).to_canonical(c4); rhs = rhs.add(JPt::from_affine(key.u).mul(c4));
  ff.to_canonical(c4); rhs = rhs.add(JPt::from_affine(key.w).mul(c4));
  Pt la = lhs.to_affine(), ra = rhs.to_affine();
  if (tr.bad) return 5;
  if (la.inf != ra.inf) return 1;
  if (!la.inf && (la.x != ra.x || la.y != ra.y)) return 1;
  return 0;
}

}  // namespace orc

using namespace orc;
extern "C" {
void* orc_keygen(const tb_cs_desc* cs, const uint8_t* g, const uint8_t* g_lagrange, const uint8_t* w, const uint8_t* u, const uint8_t* fixed,
                 const uint8_t* sigma) { return keygen(cs, g, g_lagrange, w, u, fixed, sigma); }
void orc_key_free(void* key) { delete (Key*)key; }
// returns 0 on success; *proof_len in: capacity, out: bytes written.  2 = InstanceTooLarge, 3 = ConstraintSystemFailure
int orc_prove(void* key, const uint8_t* advice, const uint8_t* instance, const uint32_t* instance_len, const uint8_t* seed, uint32_t proof_index,
              uint8_t* proof, size_t* proof_len) {
  std::vector<uint8_t> out;
  int rc = prove(*(Key*)key, advice, instance, instance_len, seed, proof_index, out);
  if (rc) return rc;
  if (out.size() > *proof_len) return -1;
  memcpy(proof, out.data(), out.size()); *proof_len = out.size();
  return 0;
}
// 0 = accept
int orc_verify(void* key, const uint8_t* instance, const uint32_t* instance_len, const uint8_t* proof, size_t proof_len) {
  return verify(*(Key*)key, instance, instance_len, proof, proof_len);
}
int orc_key_commitments(void* key_, uint8_t* fixed_comms, uint8_t* sigma_comms) {
  Key* key = (Key*)key_;
  for (size_t i = 0; i < key->fixed_comms.size(); ++i) affine_to_bytes(key->fixed_comms[i], fixed_comms + 64 * i);
  for (size_t i = 0; i < key->sig_comms.size(); ++i) affine_to_bytes(key->sig_comms[i], sigma_comms + 64 * i);
  return 0;
}
int orc_rnd(const uint8_t* seed, uint32_t proof, uint32_t tag, uint32_t idx, uint8_t* out) { rnd(seed, proof, tag, idx).to_bytes(out); return 0; }
int orc_blake2b(const uint8_t* data, size_t len, const char* personal16, uint8_t* out64) { Blake2b b(64, personal16); b.update(data, len); b.finalize(out64); return 0; }
}
