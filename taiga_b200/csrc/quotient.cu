// Row-parallel evaluation of the PLONKish gate / lookup / permutation constraint polynomial (the quotient numerator)
// and the grand-product helpers, for sm_100a.
//
// Replaces the h(X) construction of halo2_proofs `plonk::create_proof` + `vanishing::Argument::construct`,
// `permutation::Argument::commit` and `lookup::Argument::commit_product` (EXT; SURVEY.md §8a rows H3-H5, App. A.1
// step 8, App. E.3/E.6).  The extended domain is never materialised per column: for each of the R = 2^(ext_k-k)
// sub-cosets zeta*w_ext^k1*<w> the per-proof columns are NTT'd onto that sub-coset (n rows), every constraint is
// evaluated one thread per row, and the result is scaled by the (constant on the sub-coset) 1/(X^n - 1).
//
// Algorithmic bytes per sub-coset row: 32*(C+1), C = distinct column-cosets read.
#include <map>
#define TB_NOINLINE_MUL 0  // loop-structured kernels: small code, keep the multiply inline
#include "common.cuh"
#include "prover_kernels.cuh"

namespace tb {

// ---------------------------------------------------------------- expression compiler (host)
namespace {
struct Compiler {
  const tb_cs_desc* cs;
  std::vector<int> refc;          // remaining uses per node
  std::vector<int> reg_of;        // register holding node value (-1 = none)
  std::vector<int> free_regs; int next_reg = 0, max_regs = 0;
  std::vector<QInstr> code;
  struct Opnd { int kind; uint32_t v; int node; };

  explicit Compiler(const tb_cs_desc* c) : cs(c), refc(c->num_nodes, 0), reg_of(c->num_nodes, -1) {}
  void count(uint32_t node, std::vector<char>& seen) {
    refc[node]++;
    if (seen[node]) return;
    seen[node] = 1;
    const tb_expr_node& nd = cs->nodes[node];
    if (nd.op == TB_EX_NEG || nd.op == TB_EX_SCALE) count(nd.a, seen);
    else if (nd.op == TB_EX_ADD || nd.op == TB_EX_MUL) { count(nd.a, seen); count(nd.b, seen); }
  }
  int alloc() {
    int r;
    if (!free_regs.empty()) { r = free_regs.back(); free_regs.pop_back(); } else r = next_reg++;
    if (next_reg > max_regs) max_regs = next_reg;
    return r;
  }
  void release(const Opnd& o) {
    if (o.node < 0) return;
    if (--refc[o.node] == 0 && reg_of[o.node] >= 0) { free_regs.push_back(reg_of[o.node]); reg_of[o.node] = -1; }
  }
  Opnd emit(uint32_t node) {
    const tb_expr_node& nd = cs->nodes[node];
    switch (nd.op) {
      case TB_EX_CONST: return {K_CONST, nd.a, (int)node};
      case TB_EX_ADVICE: return {K_ADV, nd.a, (int)node};
      case TB_EX_FIXED: return {K_FIX, nd.a, (int)node};
      case TB_EX_INSTANCE: return {K_INST, nd.a, (int)node};
      default: break;
    }
    if (reg_of[node] >= 0) return {K_REG, (uint32_t)reg_of[node], (int)node};
    int op; Opnd oa, ob; bool binary = true;
    if (nd.op == TB_EX_NEG) { oa = emit(nd.a); ob = {K_CONST, 0, -1}; op = Q_NEG; binary = false; }
    else if (nd.op == TB_EX_SCALE) { oa = emit(nd.a); ob = {K_CONST, nd.b, -1}; op = Q_MUL; }
    else if (nd.op == TB_EX_MUL) { oa = emit(nd.a); ob = emit(nd.b); op = Q_MUL; }
    else {  // ADD, with a - b peephole when the negation is used only here
      const tb_expr_node& na = cs->nodes[nd.a]; const tb_expr_node& nb = cs->nodes[nd.b];
      if (nb.op == TB_EX_NEG && refc[nd.b] == 1 && reg_of[nd.b] < 0) {
        oa = emit(nd.a); refc[nd.b]--; ob = emit(nb.a); op = Q_SUB;
      } else if (na.op == TB_EX_NEG && refc[nd.a] == 1 && reg_of[nd.a] < 0) {
        oa = emit(nd.b); refc[nd.a]--; ob = emit(na.a); op = Q_SUB;
      } else { oa = emit(nd.a); ob = emit(nd.b); op = Q_ADD; }
    }
    // operands of leaves carry node ids only for refcounting; leaves hold no register
    release(oa); if (binary) release(ob);
    int r = alloc();
    code.push_back(q_make(op, r, oa.kind, oa.v, ob.kind, ob.v));
    reg_of[node] = r;
    return {K_REG, (uint32_t)r, (int)node};
  }
};
void finish_program(Compiler& c, QProgram* out) {
  out->host = c.code; out->nregs = c.max_regs < 1 ? 1 : c.max_regs; out->ninstr = (int)c.code.size();
  TB_REQUIRE(out->nregs <= 96, "constraint expressions need too many live temporaries");
  if (out->dev) cudaFree(out->dev);
  out->dev = nullptr;
  if (out->ninstr) {
    TB_CUDA(cudaMalloc(&out->dev, out->ninstr * sizeof(QInstr)));
    TB_CUDA(cudaMemcpy(out->dev, out->host.data(), out->ninstr * sizeof(QInstr), cudaMemcpyHostToDevice));
  }
}
}  // namespace

void q_compile_gates(const tb_cs_desc* cs, QProgram* out) {
  Compiler c(cs);
  std::vector<char> seen(cs->num_nodes, 0);
  for (uint32_t i = 0; i < cs->num_constraints; ++i) c.count(cs->constraint_roots[i], seen);
  for (uint32_t i = 0; i < cs->num_constraints; ++i) {
    Compiler::Opnd o = c.emit(cs->constraint_roots[i]);
    c.code.push_back(q_make(Q_FOLD_Y, 0, o.kind, o.v, K_CONST, 0));
    c.release(o);
  }
  finish_program(c, out);
}

void q_compile_gates_split(const tb_cs_desc* cs, int parts, std::vector<QProgram>* out, std::vector<int>* counts) {
  // cost of a root = instructions of its stand-alone program; contiguous groups with roughly equal cumulative cost
  std::vector<size_t> cost(cs->num_constraints);
  size_t total = 0;
  for (uint32_t i = 0; i < cs->num_constraints; ++i) {
    Compiler c(cs); std::vector<char> seen(cs->num_nodes, 0);
    c.count(cs->constraint_roots[i], seen);
    Compiler::Opnd o = c.emit(cs->constraint_roots[i]); (void)o;
    cost[i] = c.code.size() + 1; total += cost[i];
  }
  if (parts > (int)cs->num_constraints) parts = cs->num_constraints ? (int)cs->num_constraints : 1;
  out->clear(); counts->clear();
  uint32_t r0 = 0; size_t acc = 0;
  for (int p = 0; p < parts; ++p) {
    uint32_t r1 = r0;
    size_t target = total * (p + 1) / parts;
    while (r1 < cs->num_constraints && (acc < target || p == parts - 1)) acc += cost[r1++];
    if (p == parts - 1) r1 = cs->num_constraints;
    Compiler c(cs); std::vector<char> seen(cs->num_nodes, 0);
    for (uint32_t i = r0; i < r1; ++i) c.count(cs->constraint_roots[i], seen);
    for (uint32_t i = r0; i < r1; ++i) {
      Compiler::Opnd o = c.emit(cs->constraint_roots[i]);
      c.code.push_back(q_make(Q_FOLD_Y, 0, o.kind, o.v, K_CONST, 0));
      c.release(o);
    }
    out->emplace_back();
    finish_program(c, &out->back());
    counts->push_back((int)(r1 - r0));
    r0 = r1;
  }
}

void q_compile_lookups(const tb_cs_desc* cs, QProgram* out) {
  Compiler c(cs);
  std::vector<char> seen(cs->num_nodes, 0);
  for (uint32_t l = 0; l < cs->num_lookups; ++l)
    for (uint32_t e = 0; e < cs->lookups[l].num_exprs; ++e) { c.count(cs->lookups[l].input_roots[e], seen); c.count(cs->lookups[l].table_roots[e], seen); }
  for (uint32_t l = 0; l < cs->num_lookups; ++l) {
    c.code.push_back(q_make(Q_LK_BEGIN, 0, K_CONST, 0, K_CONST, 0));
    for (uint32_t e = 0; e < cs->lookups[l].num_exprs; ++e) {
      Compiler::Opnd o = c.emit(cs->lookups[l].input_roots[e]);
      c.code.push_back(q_make(Q_FOLD_A, 0, o.kind, o.v, K_CONST, 0)); c.release(o);
    }
    for (uint32_t e = 0; e < cs->lookups[l].num_exprs; ++e) {
      Compiler::Opnd o = c.emit(cs->lookups[l].table_roots[e]);
      c.code.push_back(q_make(Q_FOLD_S, 0, o.kind, o.v, K_CONST, 0)); c.release(o);
    }
    c.code.push_back(q_make(Q_LK_STORE, 0, K_CONST, l, K_CONST, 0));
  }
  finish_program(c, out);
}

// ---------------------------------------------------------------- interpreter kernel
__global__ void q_interp_kernel(QPartList pl, int nregs, QData d) {
  const QInstr* __restrict__ prog = pl.prog[blockIdx.z];
  const int ninstr = pl.ninstr[blockIdx.z];
  extern __shared__ uint4 q_smem[];
  const int T = blockDim.x, tid = threadIdx.x;
  uint4* rlo = q_smem;
  uint4* rhi = q_smem + (size_t)nregs * T;
  const int row = blockIdx.x * T + tid, b = blockIdx.y;
  if (row >= d.n) return;
  const int nm = d.n - 1;
  const Fp* adv = d.adv + (long long)b * d.adv_pstride;
  const Fp* inst = d.inst + (long long)b * d.inst_pstride;
  const Fp y = d.chal[(long long)b * d.chal_stride + d.y_slot];
  const Fp theta = d.chal[(long long)b * d.chal_stride + d.theta_slot];
  Fp acc = Fp::zero(), accA = Fp::zero(), accS = Fp::zero();

  auto fetch = [&](int kind, uint32_t v) -> Fp {
    switch (kind) {
      case K_REG: { uint4 x = rlo[v * T + tid], z = rhi[v * T + tid]; Fp r;
        r.l[0] = x.x; r.l[1] = x.y; r.l[2] = x.z; r.l[3] = x.w; r.l[4] = z.x; r.l[5] = z.y; r.l[6] = z.z; r.l[7] = z.w; return r; }
      case K_ADV: { int2 q = d.aq[v]; return ldg_fe(adv + (size_t)q.x * d.n + ((row + q.y + d.n) & nm)); }
      case K_FIX: { int2 q = d.fq[v]; return ldg_fe(d.fix + ((size_t)q.x * d.R + d.k1) * d.n + ((row + q.y + d.n) & nm)); }
      case K_INST: { int2 q = d.iq[v]; return ldg_fe(inst + (size_t)q.x * d.n + ((row + q.y + d.n) & nm)); }
      default: return ldg_fe(d.consts + v);
    }
  };
  for (int pc = 0; pc < ninstr; ++pc) {
    const QInstr in = prog[pc];
    const int op = in.w0 & 0xff, dst = (in.w0 >> 8) & 0xff, ak = (in.w0 >> 16) & 0xff, bk = in.w0 >> 24;
    Fp r;
    switch (op) {
      case Q_MOV: r = fetch(ak, in.a); break;
      case Q_NEG: r = fetch(ak, in.a).neg(); break;
      case Q_ADD: r = fetch(ak, in.a) + fetch(bk, in.b); break;
      case Q_SUB: r = fetch(ak, in.a) - fetch(bk, in.b); break;
      case Q_MUL: r = fetch(ak, in.a) * fetch(bk, in.b); break;
      case Q_FOLD_Y: acc = acc * y + fetch(ak, in.a); continue;
      case Q_LK_BEGIN: accA = Fp::zero(); accS = Fp::zero(); continue;
      case Q_FOLD_A: accA = accA * theta + fetch(ak, in.a); continue;
      case Q_FOLD_S: accS = accS * theta + fetch(ak, in.a); continue;
      case Q_LK_STORE: {
        size_t o = (size_t)b * d.lk_pstride + (size_t)in.a * d.n + row;
        st_fe(d.lkA + o, accA); st_fe(d.lkS + o, accS); continue; }
      default: continue;
    }
    rlo[dst * T + tid] = make_uint4(r.l[0], r.l[1], r.l[2], r.l[3]);
    rhi[dst * T + tid] = make_uint4(r.l[4], r.l[5], r.l[6], r.l[7]);
  }
  if (d.gate_out) st_fe(d.gate_out + (long long)blockIdx.z * pl.part_stride + (long long)b * d.gate_pstride + row, acc);
}

static void q_launch(Ctx* c, const QPartList& pl, int nregs, const QData& d, int B) {
  ProfScope prof_scope(c, PC_QUOT_GATES);
  static bool attr = false;
  if (!attr) { TB_CUDA(cudaFuncSetAttribute(q_interp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); attr = true; }
  int T = (96 * 1024) / (nregs * 32);
  const int tmax = tb_tune("TB_Q_THREADS", 128);
  T = T >= tmax ? tmax : (T / 32) * 32;
  TB_REQUIRE(T >= 32, "constraint program register file does not fit shared memory");
  if (d.n < T) T = d.n < 32 ? 32 : d.n;
  size_t smem = (size_t)nregs * T * 32;
  q_interp_kernel<<<dim3((d.n + T - 1) / T, B, pl.nparts), T, smem, c->stream>>>(pl, nregs, d);
  TB_LAUNCH_CHECK(); c->launches++;
}
void q_run(Ctx* c, const QProgram& prog, const QData& d, int B) {
  QPartList pl; memset(&pl, 0, sizeof(pl));
  pl.prog[0] = prog.dev; pl.ninstr[0] = prog.ninstr; pl.nparts = 1; pl.part_stride = 0;
  q_launch(c, pl, prog.nregs, d, B);
}
void q_run_parts(Ctx* c, const std::vector<QProgram>& progs, QData d, long long part_stride, int B) {
  QPartList pl; memset(&pl, 0, sizeof(pl));
  int nregs = 1;
  pl.nparts = (int)progs.size(); pl.part_stride = part_stride;
  for (int p = 0; p < pl.nparts; ++p) { pl.prog[p] = progs[p].dev; pl.ninstr[p] = progs[p].ninstr; nregs = nregs > progs[p].nregs ? nregs : progs[p].nregs; }
  q_launch(c, pl, nregs, d, B);
}

}  // namespace tb
