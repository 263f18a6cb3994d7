"""Host-side circuit description: a small mirror of halo2's `ConstraintSystem` / `Assignment` / keygen so that
tests and the benchmark can assemble exactly what the Rust shim would hand over the C ABI for a Taiga circuit
(`tb_cs_desc`, fixed columns, permutation sigma values, advice table, instance).

Mirrors (EXT halo2_proofs, called from taiga_halo2/src/circuit/compliance_circuit.rs:77-172 `configure` and
taiga_halo2/src/circuit/resource_logic_circuit.rs:321-410): `ConstraintSystem::{advice_column, fixed_column,
instance_column, selector, enable_equality, create_gate, lookup, degree, blinding_factors}`, `Expression`,
`permutation::keygen::Assembly`.  Selectors are plain fixed columns (no selector compression).
Pure host logic: no field arithmetic is offloaded here and nothing in this file is on the timed path.
"""
import ctypes
import hashlib

import numpy as np

P = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001
DELTA = pow(5, 1 << 32, P)
ROOT = pow(5, (P - 1) >> 32, P)

ADVICE, FIXED, INSTANCE = 0, 1, 2
EX_CONST, EX_ADVICE, EX_FIXED, EX_INSTANCE, EX_NEG, EX_ADD, EX_MUL, EX_SCALE = range(8)


class Column:
    def __init__(self, kind, index):
        self.kind, self.index = kind, index

    def __eq__(self, o):
        return (self.kind, self.index) == (o.kind, o.index)

    def __hash__(self):
        return hash((self.kind, self.index))

    def __repr__(self):
        return "%s%d" % ("afi"[self.kind], self.index)


class Expr:
    """halo2 `Expression<F>` node in the constraint system's hash-consed DAG."""

    def __init__(self, cs, node, degree):
        self.cs, self.node, self.degree = cs, node, degree

    def _lift(self, o):
        return o if isinstance(o, Expr) else self.cs.constant(o)

    def __add__(self, o):
        o = self._lift(o)
        return Expr(self.cs, self.cs._node(EX_ADD, self.node, o.node), max(self.degree, o.degree))

    __radd__ = __add__

    def __neg__(self):
        return Expr(self.cs, self.cs._node(EX_NEG, self.node, 0), self.degree)

    def __sub__(self, o):
        return self + (-self._lift(o))

    def __rsub__(self, o):
        return self._lift(o) + (-self)

    def __mul__(self, o):
        if isinstance(o, int):  # Expression::Scaled
            return Expr(self.cs, self.cs._node(EX_SCALE, self.node, self.cs._const(o)), self.degree)
        return Expr(self.cs, self.cs._node(EX_MUL, self.node, o.node), self.degree + o.degree)

    __rmul__ = __mul__

    def square(self):
        return self * self


class ConstraintSystem:
    def __init__(self):
        self.num_advice = self.num_fixed = self.num_instance = 0
        self.advice_queries, self.fixed_queries, self.instance_queries = [], [], []
        self.num_advice_queries = []
        self.perm_columns = []
        self.constants, self._const_index = [], {}
        self.nodes, self._node_index = [], {}
        self.gates = []       # (name, [Expr])
        self.lookups = []     # [(input Expr, table Expr), ...]
        self.constants_column = None

    # ---- columns
    def advice_column(self):
        self.num_advice += 1
        self.num_advice_queries.append(0)
        return Column(ADVICE, self.num_advice - 1)

    def fixed_column(self):
        self.num_fixed += 1
        return Column(FIXED, self.num_fixed - 1)

    def instance_column(self):
        self.num_instance += 1
        return Column(INSTANCE, self.num_instance - 1)

    def selector(self):
        return self.fixed_column()

    def enable_equality(self, col):
        self.query_any_index(col, 0)
        if col not in self.perm_columns:
            self.perm_columns.append(col)

    def enable_constant(self, col):
        assert col.kind == FIXED
        self.constants_column = col
        self.enable_equality(col)

    # ---- expression pool
    def _const(self, v):
        v %= P
        if v not in self._const_index:
            self._const_index[v] = len(self.constants)
            self.constants.append(v)
        return self._const_index[v]

    def _node(self, op, a, b):
        key = (op, a, b)
        if key not in self._node_index:
            self._node_index[key] = len(self.nodes)
            self.nodes.append(key)
        return self._node_index[key]

    def constant(self, v):
        return Expr(self, self._node(EX_CONST, self._const(v), 0), 0)

    def query_any_index(self, col, rot):
        qs = (self.advice_queries, self.fixed_queries, self.instance_queries)[col.kind]
        key = (col.index, rot)
        if key in qs:
            return qs.index(key)
        qs.append(key)
        if col.kind == ADVICE:
            self.num_advice_queries[col.index] += 1
        return len(qs) - 1

    def query(self, col, rot=0):
        """query_advice / query_fixed / query_instance / query_selector."""
        qi = self.query_any_index(col, rot)
        return Expr(self, self._node((EX_ADVICE, EX_FIXED, EX_INSTANCE)[col.kind], qi, 0), 1)

    def create_gate(self, name, polys):
        polys = list(polys)
        assert polys
        self.gates.append((name, polys))

    def lookup(self, pairs):
        pairs = list(pairs)
        assert pairs
        self.lookups.append(pairs)
        return len(self.lookups) - 1

    # ---- derived quantities (halo2 plonk/circuit.rs)
    def degree(self):
        d = 3   # permutation.required_degree() is 3 whether or not any column is equality-enabled (halo2 plonk/circuit.rs)
        for lk in self.lookups:
            ind = max([1] + [i.degree for i, _ in lk])
            td = max([1] + [t.degree for _, t in lk])
            d = max(d, max(4, 2 + ind + td))
        for _, polys in self.gates:
            for p in polys:
                d = max(d, p.degree)
        return d

    def blinding_factors(self):
        return max(3, max(self.num_advice_queries + [1])) + 2

    def minimum_rows(self):
        return self.blinding_factors() + 3


class Assignment:
    """Witness + fixed assignment for one circuit instance (halo2 `Assignment`, simple sequential floor plan)."""

    def __init__(self, cs, k):
        self.cs, self.k, self.n = cs, k, 1 << k
        self.usable = self.n - (cs.blinding_factors() + 1)
        self.advice = [dict() for _ in range(cs.num_advice)]
        self.fixed = [dict() for _ in range(cs.num_fixed)]
        self.instance = [[] for _ in range(cs.num_instance)]
        self.copies = []
        self._const_cells = {}
        self._const_row = 0

    def assign(self, col, row, value):
        assert 0 <= row < self.usable, "row %d outside usable rows" % row
        (self.advice if col.kind == ADVICE else self.fixed)[col.index][row] = value % P
        return (col, row)

    def enable(self, selector, row):
        self.assign(selector, row, 1)

    def copy(self, a, b):
        for col, row in (a, b):
            assert col in self.cs.perm_columns, "column %r not equality-enabled" % col
        self.copies.append((a, b))

    def constant_cell(self, value):
        """assign_advice_from_constant support: a cell of the constants column holding `value`."""
        value %= P
        if value not in self._const_cells:
            col = self.cs.constants_column
            self._const_cells[value] = self.assign(col, self._const_row, value)
            self._const_row += 1
        return self._const_cells[value]

    def set_instance(self, col, values):
        self.instance[col.index] = [v % P for v in values]

    def value(self, cell):
        col, row = cell
        if col.kind == INSTANCE:
            vals = self.instance[col.index]
            return vals[row] if row < len(vals) else 0
        return (self.advice if col.kind == ADVICE else self.fixed)[col.index].get(row, 0)


def _col_bytes(dicts, n):
    """{row: value} per column -> [columns, n, 32] little-endian bytes (one join + one scatter per column)."""
    out = np.zeros((len(dicts), n, 32), np.uint8)
    for c, d in enumerate(dicts):
        if not d:
            continue
        rows = np.fromiter(d.keys(), dtype=np.int64, count=len(d))
        vals = np.frombuffer(b"".join([int(v).to_bytes(32, "little") for v in d.values()]), np.uint8).reshape(len(d), 32)
        out[c, rows] = vals
    return out


def build_sigma(cs, asg):
    """permutation::keygen::Assembly::{copy, build_pk}: sigma[col][row] = delta^col' * omega^row'."""
    n, k = asg.n, asg.k
    cols = cs.perm_columns
    m = len(cols)
    idx = {c: i for i, c in enumerate(cols)}
    mapping = [[(i, j) for j in range(n)] for i in range(m)]
    aux = [[(i, j) for j in range(n)] for i in range(m)]
    sizes = [[1] * n for _ in range(m)]
    for (ca, ra), (cb, rb) in asg.copies:
        a, b = (idx[ca], ra), (idx[cb], rb)
        if aux[a[0]][a[1]] == aux[b[0]][b[1]]:
            continue
        la, lb = aux[a[0]][a[1]], aux[b[0]][b[1]]
        if sizes[la[0]][la[1]] < sizes[lb[0]][lb[1]]:
            a, b, la, lb = b, a, lb, la
        sizes[la[0]][la[1]] += sizes[lb[0]][lb[1]]
        # relabel b's cycle with a's leader
        i, j = b
        while True:
            aux[i][j] = la
            i, j = mapping[i][j]
            if (i, j) == b:
                break
        ta = mapping[a[0]][a[1]]
        mapping[a[0]][a[1]] = mapping[b[0]][b[1]]
        mapping[b[0]][b[1]] = ta
    omega = pow(ROOT, 1 << (32 - k), P)
    om = [1] * n
    for j in range(1, n):
        om[j] = om[j - 1] * omega % P
    dl = [pow(DELTA, i, P) for i in range(m)]
    sigma = np.zeros((m, n, 32), np.uint8)
    ident = {}
    for i in range(m):
        rowbuf = bytearray(32 * n)
        for j in range(n):
            pi, pj = mapping[i][j]
            key = (pi, pj)
            v = ident.get(key)
            if v is None:
                v = (dl[pi] * om[pj] % P).to_bytes(32, "little")
                ident[key] = v
            rowbuf[32 * j:32 * j + 32] = v
        sigma[i] = np.frombuffer(bytes(rowbuf), np.uint8).reshape(n, 32)
    return sigma


# ---------------------------------------------------------------- ctypes mirror of include/taiga_b200.h descriptor types
class TbQuery(ctypes.Structure):
    _fields_ = [("column", ctypes.c_uint32), ("rotation", ctypes.c_int32)]


class TbColumn(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_uint32), ("index", ctypes.c_uint32)]


class TbExprNode(ctypes.Structure):
    _fields_ = [("op", ctypes.c_uint32), ("a", ctypes.c_uint32), ("b", ctypes.c_uint32)]


class TbLookup(ctypes.Structure):
    _fields_ = [("num_exprs", ctypes.c_uint32), ("input_roots", ctypes.POINTER(ctypes.c_uint32)), ("table_roots", ctypes.POINTER(ctypes.c_uint32))]


class TbCsDesc(ctypes.Structure):
    _fields_ = [("k", ctypes.c_uint32), ("num_advice", ctypes.c_uint32), ("num_fixed", ctypes.c_uint32), ("num_instance", ctypes.c_uint32),
                ("cs_degree", ctypes.c_uint32), ("blinding_factors", ctypes.c_uint32),
                ("num_advice_queries", ctypes.c_uint32), ("advice_queries", ctypes.POINTER(TbQuery)),
                ("num_fixed_queries", ctypes.c_uint32), ("fixed_queries", ctypes.POINTER(TbQuery)),
                ("num_instance_queries", ctypes.c_uint32), ("instance_queries", ctypes.POINTER(TbQuery)),
                ("num_perm_columns", ctypes.c_uint32), ("perm_columns", ctypes.POINTER(TbColumn)),
                ("num_constants", ctypes.c_uint32), ("constants", ctypes.POINTER(ctypes.c_uint8)),
                ("num_nodes", ctypes.c_uint32), ("nodes", ctypes.POINTER(TbExprNode)),
                ("num_constraints", ctypes.c_uint32), ("constraint_roots", ctypes.POINTER(ctypes.c_uint32)),
                ("num_lookups", ctypes.c_uint32), ("lookups", ctypes.POINTER(TbLookup)),
                ("vk_transcript_repr", ctypes.c_uint8 * 32)]


class CircuitKeyData:
    """Everything keygen produces for one circuit: descriptor + fixed column values + sigma values."""

    def __init__(self, cs, k, fixed_asg, name="circuit"):
        self.cs, self.k, self.n, self.name = cs, k, 1 << k, name
        assert self.n >= cs.minimum_rows()
        self.degree = cs.degree()
        self.blinding_factors = cs.blinding_factors()
        self.fixed = _col_bytes(fixed_asg.fixed, self.n)              # [num_fixed, n, 32]
        self.sigma = build_sigma(cs, fixed_asg)                        # [P, n, 32]
        self._keep = []
        self.desc = self._make_desc()

    def _arr(self, ctype, items):
        arr = (ctype * max(1, len(items)))(*items)
        self._keep.append(arr)
        return arr

    def _make_desc(self):
        cs = self.cs
        d = TbCsDesc()
        d.k, d.num_advice, d.num_fixed, d.num_instance = self.k, cs.num_advice, cs.num_fixed, cs.num_instance
        d.cs_degree, d.blinding_factors = self.degree, self.blinding_factors
        for name, qs in (("advice", cs.advice_queries), ("fixed", cs.fixed_queries), ("instance", cs.instance_queries)):
            setattr(d, "num_%s_queries" % name, len(qs))
            setattr(d, "%s_queries" % name, self._arr(TbQuery, [TbQuery(c, r) for c, r in qs]))
        d.num_perm_columns = len(cs.perm_columns)
        d.perm_columns = self._arr(TbColumn, [TbColumn(c.kind, c.index) for c in cs.perm_columns])
        cbytes = b"".join(int(v).to_bytes(32, "little") for v in cs.constants) or bytes(32)
        d.num_constants = len(cs.constants)
        d.constants = self._arr(ctypes.c_uint8, list(cbytes))
        d.num_nodes = len(cs.nodes)
        d.nodes = self._arr(TbExprNode, [TbExprNode(*nd) for nd in cs.nodes])
        roots = [p.node for _, polys in cs.gates for p in polys]
        d.num_constraints = len(roots)
        d.constraint_roots = self._arr(ctypes.c_uint32, roots)
        lks = []
        for lk in cs.lookups:
            ins = self._arr(ctypes.c_uint32, [i.node for i, _ in lk])
            tabs = self._arr(ctypes.c_uint32, [t.node for _, t in lk])
            lks.append(TbLookup(len(lk), ins, tabs))
        d.num_lookups = len(lks)
        d.lookups = self._arr(TbLookup, lks)
        # Stand-in for vk.transcript_repr (the real one is a BLAKE2b of Rust's Debug string of the pinned vk,
        # taiga_halo2/src/resource_logic_vk.rs:33-48, owned by the Rust side): hash of the structural description.
        h = hashlib.blake2b(digest_size=64, person=b"TaigaB200-VkRepr")
        h.update(repr((self.k, cs.num_advice, cs.num_fixed, cs.num_instance, cs.advice_queries, cs.fixed_queries, cs.instance_queries,
                       [(c.kind, c.index) for c in cs.perm_columns], cs.constants, cs.nodes, roots,
                       [[(i.node, t.node) for i, t in lk] for lk in cs.lookups])).encode())
        h.update(hashlib.sha256(self.fixed.tobytes()).digest() + hashlib.sha256(self.sigma.tobytes()).digest())
        self.vk_repr = int.from_bytes(h.digest(), "little") % P
        d.vk_transcript_repr = (ctypes.c_uint8 * 32)(*self.vk_repr.to_bytes(32, "little"))
        return d

    def witness_arrays(self, asg):
        """advice [num_advice, n, 32] bytes, instance bytes (columns concatenated), instance_len uint32[num_instance]."""
        adv = _col_bytes(asg.advice, self.n)
        inst = b"".join(int(v).to_bytes(32, "little") for col in asg.instance for v in col)
        inst_arr = np.frombuffer(inst, np.uint8).copy() if inst else np.zeros(32, np.uint8)
        lens = np.array([len(c) for c in asg.instance] or [0], dtype=np.uint32)
        return adv, inst_arr, lens

    def proof_size(self):
        """SURVEY App. D accounting (bytes)."""
        cs = self.cs
        nsets = -(-len(cs.perm_columns) // (self.degree - 2)) if cs.perm_columns else 0
        L = len(cs.lookups)
        commits = cs.num_advice + 2 * L + nsets + L + 1 + (self.degree - 1)
        evals = len(cs.instance_queries) + len(cs.advice_queries) + len(cs.fixed_queries) + 1 + len(cs.perm_columns) + max(0, 3 * nsets - 1) + 5 * L
        return 32 * (commits + evals + 1 + self.num_point_sets() + 1 + 2 * self.k + 2)

    def num_point_sets(self):
        cs = self.cs
        bf = self.blinding_factors
        nsets = -(-len(cs.perm_columns) // (self.degree - 2)) if cs.perm_columns else 0
        rots = {}

        def add(key, rot):
            rots.setdefault(key, set()).add(rot)
        for c, r in cs.instance_queries:
            add(("i", c), r)
        for c, r in cs.advice_queries:
            add(("a", c), r)
        for s in range(nsets):
            add(("pz", s), 0), add(("pz", s), 1)
            if s + 1 < nsets:
                add(("pz", s), -(bf + 1))
        for l in range(len(cs.lookups)):
            add(("lz", l), 0), add(("lz", l), 1), add(("lin", l), 0), add(("lin", l), -1), add(("ltab", l), 0)
        for c, r in cs.fixed_queries:
            add(("f", c), r)
        for c in range(len(cs.perm_columns)):
            add(("sig", c), 0)
        add("h", 0), add("rand", 0)
        return len({frozenset(v) for v in rots.values()})
