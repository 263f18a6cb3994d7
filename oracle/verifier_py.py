"""ORACLE (test infrastructure, never shipped): an INDEPENDENT halo2 / IPA verifier in pure Python.

oracle/plonk.cpp holds a prover AND a verifier written in one pass from one reading of halo2, so their agreement
cannot catch a shared misreading.  This file restates `plonk::verify_proof` (SingleVerifier), `poly::multiopen::verify_proof`
and `poly::commitment::verify_proof` of the zcash/halo2 0.3 lineage a second time -- big-integer field arithmetic,
Jacobian curve arithmetic and a bucket MSM written here, hashlib for BLAKE2b -- sharing no code with plonk.cpp,
field.hpp or prims.hpp.  It follows SURVEY.md Appendix A (A.1 steps 1-10 for the transcript order, A.2 for multiopen
and the inner product argument, A.3 for the transcript, A.4 for the final check) and is the stand-in for
`Proof::verify` (taiga_halo2/src/proof.rs:45-54), which cannot run here (no Rust toolchain).

What it pins: the golden proofs of tests/golden/ (made by plonk.cpp, reproduced byte for byte by the CUDA prover) are
accepted by a verifier that was written separately.  What it cannot pin: that both restatements match the real crate.

Inputs: a taiga_b200.circuit.CircuitKeyData (constraint system, k, vk transcript representation), the SRS as affine
byte arrays, the verifying key commitments (fixed columns, permutation sigmas; 64-byte affine points -- the verifier's
INPUT in halo2 too), the instance columns and the proof bytes.
"""
import hashlib

P = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001   # circuit field = Vesta scalar field
Q = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001   # Vesta base field
ROOT_2_32 = pow(5, (P - 1) >> 32, P)
DELTA = pow(5, 1 << 32, P)

EX_CONST, EX_ADVICE, EX_FIXED, EX_INSTANCE, EX_NEG, EX_ADD, EX_MUL, EX_SCALE = range(8)
ADVICE, FIXED, INSTANCE = 0, 1, 2


class Reject(Exception):
    pass


# ---------------------------------------------------------------- Vesta: y^2 = x^3 + 5 over Fq, Jacobian (X, Y, Z), identity Z = 0
def _from_affine(b):
    x, y = int.from_bytes(bytes(b[:32]), "little"), int.from_bytes(bytes(b[32:64]), "little")
    if x == 0 and y == 0:
        return (0, 1, 0)
    if (y * y - x * x * x - 5) % Q:
        raise Reject("point not on Vesta")
    return (x, y, 1)


def _dbl(p):
    X, Y, Z = p
    if Z == 0 or Y == 0:
        return (0, 1, 0)
    A = X * X % Q
    B = Y * Y % Q
    C = B * B % Q
    D = 2 * ((X + B) * (X + B) - A - C) % Q
    E = 3 * A % Q
    X3 = (E * E - 2 * D) % Q
    return (X3, (E * (D - X3) - 8 * C) % Q, 2 * Y * Z % Q)


def _add(p, q):
    if p[2] == 0:
        return q
    if q[2] == 0:
        return p
    X1, Y1, Z1 = p
    X2, Y2, Z2 = q
    Z1Z1, Z2Z2 = Z1 * Z1 % Q, Z2 * Z2 % Q
    U1, U2 = X1 * Z2Z2 % Q, X2 * Z1Z1 % Q
    S1, S2 = Y1 * Z2 * Z2Z2 % Q, Y2 * Z1 * Z1Z1 % Q
    if U1 == U2:
        return _dbl(p) if S1 == S2 else (0, 1, 0)
    H, R = (U2 - U1) % Q, (S2 - S1) % Q
    HH = H * H % Q
    HHH = H * HH % Q
    V = U1 * HH % Q
    X3 = (R * R - HHH - 2 * V) % Q
    return (X3, (R * (V - X3) - S1 * HHH) % Q, Z1 * Z2 * H % Q)


def _neg(p):
    return (p[0], (-p[1]) % Q, p[2])


def _mul(p, k):
    k %= P
    acc = (0, 1, 0)
    for bit in bin(k)[2:] if k else "":
        acc = _dbl(acc)
        if bit == "1":
            acc = _add(acc, p)
    return acc


def _is_identity(p):
    return p[2] == 0


def msm(scalars, points, c=None):
    """sum scalars[i] * points[i] (bucket method, unsigned c-bit windows); points Jacobian."""
    n = len(scalars)
    if n == 0:
        return (0, 1, 0)
    if c is None:
        c = 4 if n < 64 else (8 if n < 4096 else 12)
    acc = (0, 1, 0)
    for w in reversed(range((255 + c - 1) // c + 1)):
        for _ in range(c):
            acc = _dbl(acc)
        buckets = [None] * (1 << c)
        sh = w * c
        for s, pt in zip(scalars, points):
            d = (s >> sh) & ((1 << c) - 1)
            if d:
                buckets[d] = pt if buckets[d] is None else _add(buckets[d], pt)
        run, tot = (0, 1, 0), (0, 1, 0)
        for d in range((1 << c) - 1, 0, -1):
            if buckets[d] is not None:
                run = _add(run, buckets[d])
            tot = _add(tot, run)
        acc = _add(acc, tot)
    return acc


def _sqrt_q(a):
    """square root in Fq (2-adicity 32, Tonelli-Shanks), None if a is not a square."""
    a %= Q
    if a == 0:
        return 0
    if pow(a, (Q - 1) // 2, Q) != 1:
        return None
    s, t = 32, (Q - 1) >> 32
    z = pow(5, t, Q)            # 5 generates the multiplicative group: 5^t has order 2^32
    m, cc, tt, r = s, z, pow(a, t, Q), pow(a, (t + 1) // 2, Q)
    while tt != 1:
        i, t2 = 0, tt
        while t2 != 1:
            t2 = t2 * t2 % Q
            i += 1
        b = pow(cc, 1 << (m - i - 1), Q)
        m, cc = i, b * b % Q
        tt, r = tt * cc % Q, r * b % Q
    return r


def decompress(enc):
    """pasta encoding: x little-endian, bit 255 = parity of y, all zeros = identity."""
    v = int.from_bytes(bytes(enc), "little")
    sign, x = v >> 255, v & ((1 << 255) - 1)
    if x == 0 and sign == 0:
        return (0, 1, 0)
    if x >= Q:
        raise Reject("non-canonical x")
    y = _sqrt_q(x * x * x + 5)
    if y is None:
        raise Reject("x is not on the curve")
    if (y & 1) != sign:
        y = Q - y
    return (x, y, 1)


def _to_affine(p):
    if p[2] == 0:
        return (0, 0)
    zi = pow(p[2], Q - 2, Q)
    zi2 = zi * zi % Q
    return (p[0] * zi2 % Q, p[1] * zi2 * zi % Q)


# ---------------------------------------------------------------- transcript (Blake2bRead + Challenge255)
class Transcript:
    def __init__(self, proof):
        self.h = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")
        self.buf, self.pos = bytes(proof), 0

    def common_point(self, p):
        if _is_identity(p):
            raise Reject("identity in transcript")
        x, y = _to_affine(p)
        self.h.update(b"\x01" + x.to_bytes(32, "little") + y.to_bytes(32, "little"))

    def common_scalar(self, s):
        self.h.update(b"\x02" + (s % P).to_bytes(32, "little"))

    def read_point(self):
        if self.pos + 32 > len(self.buf):
            raise Reject("proof too short")
        p = decompress(self.buf[self.pos:self.pos + 32])
        self.pos += 32
        self.common_point(p)
        return p

    def read_scalar(self):
        if self.pos + 32 > len(self.buf):
            raise Reject("proof too short")
        v = int.from_bytes(self.buf[self.pos:self.pos + 32], "little")
        self.pos += 32
        if v >= P:
            raise Reject("non-canonical scalar")
        self.common_scalar(v)
        return v

    def squeeze(self):
        self.h.update(b"\x00")
        return int.from_bytes(self.h.copy().digest(), "little") % P


# ---------------------------------------------------------------- helpers on the evaluation domain
def _inv(a):
    return pow(a % P, P - 2, P)


def lagrange_interpolate(xs, ys):
    """coefficients (low to high) of the polynomial through (xs[i], ys[i])."""
    n = len(xs)
    coeffs = [0] * n
    for i in range(n):
        num = [1]
        den = 1
        for j in range(n):
            if j == i:
                continue
            num = [((num[t - 1] if t > 0 else 0) - xs[j] * (num[t] if t < len(num) else 0)) % P for t in range(len(num) + 1)]
            den = den * (xs[i] - xs[j]) % P
        s = ys[i] * _inv(den) % P
        for t in range(n):
            coeffs[t] = (coeffs[t] + s * num[t]) % P
    return coeffs


def _poly_eval(c, x):
    acc = 0
    for v in reversed(c):
        acc = (acc * x + v) % P
    return acc


def verify(kd, srs, fixed_commitments, sigma_commitments, instance_columns, proof, msm_big=None):
    """True iff `proof` is accepted.  instance_columns: list (one per instance column) of lists of ints.
    msm_big(scalars, which) may supply the n-term MSM over srs['g'] (default: the bucket MSM above, ~10 s at k = 15)."""
    try:
        return _verify(kd, srs, fixed_commitments, sigma_commitments, instance_columns, proof, msm_big)
    except Reject:
        return False


def _verify(kd, srs, fixed_commitments, sigma_commitments, instance_columns, proof, msm_big):
    cs, k, n = kd.cs, kd.k, kd.n
    d = cs.degree()
    bf = cs.blinding_factors()
    omega = pow(ROOT_2_32, 1 << (32 - k), P)
    omega_inv = _inv(omega)
    P_cols = list(cs.perm_columns)
    chunk = d - 2
    nsets = -(-len(P_cols) // chunk) if P_cols else 0
    L = len(cs.lookups)
    g = [_from_affine(b) for b in srs["g"]] if msm_big is None else None
    g0 = _from_affine(srs["g"][0])
    W, U = _from_affine(srs["w"]), _from_affine(srs["u"])
    fixed_c = [_from_affine(b) for b in fixed_commitments]
    sigma_c = [_from_affine(b) for b in sigma_commitments]
    if len(instance_columns) != cs.num_instance:
        raise Reject("instance columns")

    tr = Transcript(proof)
    tr.common_scalar(kd.vk_repr)                                                   # A.1 step 1
    # A.1 step 2: commit_lagrange(instance, Blind::default() = 1), hashed but not read from the proof
    inst_c = []
    for col in instance_columns:
        if len(col) > n - (bf + 1):
            raise Reject("InstanceTooLarge")
        pts = [_from_affine(srs["g_lagrange"][i]) for i in range(len(col))]
        c = _add(msm([v % P for v in col], pts), W)
        tr.common_point(c)
        inst_c.append(c)
    adv_c = [tr.read_point() for _ in range(cs.num_advice)]                        # step 3
    theta = tr.squeeze()                                                           # step 4
    lk_perm = [(tr.read_point(), tr.read_point()) for _ in range(L)]
    beta, gamma = tr.squeeze(), tr.squeeze()                                       # step 5
    pz_c = [tr.read_point() for _ in range(nsets)]
    lz_c = [tr.read_point() for _ in range(L)]                                     # step 6
    random_c = tr.read_point()                                                     # step 7
    y = tr.squeeze()
    h_c = [tr.read_point() for _ in range(d - 1)]                                  # step 8
    x = tr.squeeze()                                                               # step 9
    inst_ev = [tr.read_scalar() for _ in cs.instance_queries]
    adv_ev = [tr.read_scalar() for _ in cs.advice_queries]
    fix_ev = [tr.read_scalar() for _ in cs.fixed_queries]
    random_ev = tr.read_scalar()
    sig_ev = [tr.read_scalar() for _ in P_cols]
    pz_ev = []
    for s in range(nsets):
        e, en = tr.read_scalar(), tr.read_scalar()
        el = tr.read_scalar() if s + 1 < nsets else None
        pz_ev.append((e, en, el))
    lk_ev = [tuple(tr.read_scalar() for _ in range(5)) for _ in range(L)]          # Z(x), Z(wx), A'(x), A'(w^-1 x), S'(x)

    # ---- A.4: recompute the numerator of h at x
    xn = pow(x, n, P)
    if xn == 1:
        raise Reject("x in the domain")

    def l_i(i):   # Lagrange basis polynomial of row i (mod n) at x
        wi = pow(omega, i % n, P)
        return (xn - 1) * wi % P * _inv(n * (x - wi)) % P
    l_last = l_i(-(bf + 1))
    l_blind = sum(l_i(-r) for r in range(1, bf + 1)) % P
    l_0 = l_i(0)
    active = (1 - l_last - l_blind) % P

    memo = {}

    def ev(node):
        if node in memo:
            return memo[node]
        op, a, b = cs.nodes[node]
        if op == EX_CONST:
            v = cs.constants[a]
        elif op == EX_ADVICE:
            v = adv_ev[a]
        elif op == EX_FIXED:
            v = fix_ev[a]
        elif op == EX_INSTANCE:
            v = inst_ev[a]
        elif op == EX_NEG:
            v = -ev(a)
        elif op == EX_ADD:
            v = ev(a) + ev(b)
        elif op == EX_MUL:
            v = ev(a) * ev(b)
        else:
            v = ev(a) * cs.constants[b]
        memo[node] = v % P
        return memo[node]

    def col_eval(col):   # evaluation of a permutation column at x (its Rotation::cur query)
        qs = (cs.advice_queries, cs.fixed_queries, cs.instance_queries)[col.kind]
        evs = (adv_ev, fix_ev, inst_ev)[col.kind]
        return evs[qs.index((col.index, 0))]

    terms = []
    for _, polys in cs.gates:
        for p in polys:
            terms.append(ev(p.node))
    if nsets:
        terms.append(l_0 * (1 - pz_ev[0][0]) % P)
        zl = pz_ev[-1][0]
        terms.append(l_last * (zl * zl - zl) % P)
        for s in range(1, nsets):
            terms.append(l_0 * (pz_ev[s][0] - pz_ev[s - 1][2]) % P)
        for s in range(nsets):
            cols = P_cols[s * chunk:(s + 1) * chunk]
            left, right = pz_ev[s][1], pz_ev[s][0]
            cur = beta * x % P * pow(DELTA, s * chunk, P) % P
            for ci, col in enumerate(cols):
                v = col_eval(col)
                left = left * (v + beta * sig_ev[s * chunk + ci] + gamma) % P
                right = right * (v + cur + gamma) % P
                cur = cur * DELTA % P
            terms.append((left - right) * active % P)
    for l, lk in enumerate(cs.lookups):
        z, zn, ap, apm, sp = lk_ev[l]
        a_c = t_c = 0
        for inp, tab in lk:
            a_c = (a_c * theta + ev(inp.node)) % P
            t_c = (t_c * theta + ev(tab.node)) % P
        terms.append(l_0 * (1 - z) % P)
        terms.append(l_last * (z * z - z) % P)
        terms.append((zn * (ap + beta) % P * (sp + gamma) - z * (a_c + beta) % P * (t_c + gamma)) * active % P)
        terms.append(l_0 * (ap - sp) % P)
        terms.append((ap - sp) * (ap - apm) % P * active % P)
    num = 0
    for t in terms:
        num = (num * y + t) % P
    expected_h = num * _inv(xn - 1) % P

    # h commitment = sum_i xn^i * piece_i
    h_comm = (0, 1, 0)
    for piece in reversed(h_c):
        h_comm = _add(_mul(h_comm, xn), piece)

    # ---- A.1 step 10: the queries (commitment, rotation, eval), in the verifier's order
    last_rot = -(bf + 1)
    queries = []
    for qi, (c, r) in enumerate(cs.instance_queries):
        queries.append((("i", c), inst_c[c], r, inst_ev[qi]))
    for qi, (c, r) in enumerate(cs.advice_queries):
        queries.append((("a", c), adv_c[c], r, adv_ev[qi]))
    for s in range(nsets):
        queries.append((("pz", s), pz_c[s], 0, pz_ev[s][0]))
        queries.append((("pz", s), pz_c[s], 1, pz_ev[s][1]))
    for s in reversed(range(nsets)):
        if s + 1 < nsets:
            queries.append((("pz", s), pz_c[s], last_rot, pz_ev[s][2]))
    for l in range(L):
        z, zn, ap, apm, sp = lk_ev[l]
        queries.append((("lz", l), lz_c[l], 0, z))
        queries.append((("la", l), lk_perm[l][0], 0, ap))
        queries.append((("ls", l), lk_perm[l][1], 0, sp))
        queries.append((("la", l), lk_perm[l][0], -1, apm))
        queries.append((("lz", l), lz_c[l], 1, zn))
    for qi, (c, r) in enumerate(cs.fixed_queries):
        queries.append((("f", c), fixed_c[c], r, fix_ev[qi]))
    for ci in range(len(P_cols)):
        queries.append((("sig", ci), sigma_c[ci], 0, sig_ev[ci]))
    queries.append((("h",), h_comm, 0, expected_h))
    queries.append((("rand",), random_c, 0, random_ev))

    # ---- A.2 multiopen: group the commitments by their SET of evaluation points
    x1, x2 = tr.squeeze(), tr.squeeze()
    order, rots_of, comm_of, eval_at = [], {}, {}, {}
    for key, comm, rot, e in queries:
        if key not in rots_of:
            order.append(key)
            rots_of[key], comm_of[key] = [], comm
        if rot not in rots_of[key]:
            rots_of[key].append(rot)
        eval_at[(key, rot)] = e
    set_keys, set_of = [], {}
    for key in order:
        fs = frozenset(rots_of[key])
        if fs not in set_keys:
            set_keys.append(fs)
        set_of[key] = set_keys.index(fs)
    # the points of a set in the order halo2 lists them: order of first appearance over all queries
    rot_order = []
    for _, _, rot, _ in queries:
        if rot not in rot_order:
            rot_order.append(rot)
    set_rots = [[r for r in rot_order if r in fs] for fs in set_keys]
    q_comm = [(0, 1, 0)] * len(set_keys)
    q_evs = [[0] * len(r) for r in set_rots]
    for key in order:
        s = set_of[key]
        q_comm[s] = _add(_mul(q_comm[s], x1), comm_of[key])
        for i, r in enumerate(set_rots[s]):
            q_evs[s][i] = (q_evs[s][i] * x1 + eval_at[(key, r)]) % P
    f_comm = tr.read_point()
    x3 = tr.squeeze()
    u = [tr.read_scalar() for _ in set_keys]
    x4 = tr.squeeze()

    def point(rot):
        return x * pow(omega if rot >= 0 else omega_inv, abs(rot), P) % P
    f_eval = 0
    for s in range(len(set_keys)):
        pts = [point(r) for r in set_rots[s]]
        r_eval = _poly_eval(lagrange_interpolate(pts, q_evs[s]), x3)
        den = 1
        for pt in pts:
            den = den * (x3 - pt) % P
        f_eval = (f_eval * x2 + (u[s] - r_eval) * _inv(den)) % P
    final_c, v = f_comm, f_eval
    for s in range(len(set_keys)):
        final_c = _add(_mul(final_c, x4), q_comm[s])
        v = (v * x4 + u[s]) % P

    # ---- inner product argument: open final_c at x3 to v
    s_comm = tr.read_point()
    xi, z = tr.squeeze(), tr.squeeze()
    rounds = []
    for _ in range(k):
        lj, rj = tr.read_point(), tr.read_point()
        rounds.append((lj, rj, tr.squeeze()))
    c, f = tr.read_scalar(), tr.read_scalar()
    if tr.pos != len(tr.buf):
        raise Reject("trailing bytes")
    us = [r[2] for r in rounds]
    lhs = _add(final_c, _mul(g0, -v))
    lhs = _add(lhs, _mul(s_comm, xi))
    for lj, rj, uj in rounds:
        lhs = _add(lhs, _add(_mul(lj, _inv(uj)), _mul(rj, uj)))
    # b = prod_j (1 + u_{k-1-j} x3^(2^j)),   s_i = prod over the set bits t of i of u_{k-1-t}
    b, cur = 1, x3
    for uj in reversed(us):
        b = b * (1 + uj * cur) % P
        cur = cur * cur % P
    svec = [1]
    for uj in reversed(us):
        svec = svec + [t * uj % P for t in svec]
    if msm_big is not None:
        gs = msm_big(svec)
    else:
        gs = msm(svec, g)
    rhs = _add(_mul(gs, c), _add(_mul(U, c * b % P * z % P), _mul(W, f)))
    return _to_affine(lhs) == _to_affine(rhs)
