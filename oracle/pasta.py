"""ORACLE (test infrastructure, never shipped): Python big-int restatement of the Pasta
fields/curves, the halo2 evaluation-domain conventions, naive MSM/NTT and the BLAKE2b
transcript used by Taiga's prover path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

The arithmetic lives in third-party crates that are NOT vendored under /root/reference:
  * pasta_curves 0.5.1 (heliaxdev fork, branch `taiga`; /root/reference/taiga_halo2/Cargo.toml:10)
  * halo2_proofs  (heliaxdev/halo2 branch `taiga`; /root/reference/taiga_halo2/Cargo.toml:14-15)
so this file restates their published algorithms.  It is pinned by the reference's own
fixture /root/reference/taiga_halo2/params/params_15 (see tests/test_oracle_fixture.py and
SURVEY.md App. B.2) and by the call sites taiga_halo2/src/proof.rs:25-54.
Small cases only (pure Python loops): k <= 10 circuits, KAT generation.
"""
import hashlib
import struct

# ---------------------------------------------------------------- fields
P = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001  # Fp: Pallas base / Vesta scalar
Q = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001  # Fq: Vesta base / Pallas scalar
S = 32                                   # 2-adicity of both fields
GEN = 5                                  # multiplicative generator of both fields
ROOT_P = pow(GEN, (P - 1) >> S, P)       # Fp::ROOT_OF_UNITY (2^32-th root)
ROOT_Q = pow(GEN, (Q - 1) >> S, Q)
DELTA_P = pow(GEN, 1 << S, P)            # Fp::DELTA
DELTA_Q = pow(GEN, 1 << S, Q)
ZETA_P = 0x12CCCA834ACDBA712CAAD5DC57AAB1B01D1F8BD237AD31491DAD5EBDFDFE4AB9  # Fp::ZETA
ZETA_Q = 0x06819A58283E528E511DB4D81CF70F5A0FED467D47C033AF2AA9D2E050AA0E4F  # Fq::ZETA
CURVE_B = 5                              # both curves: y^2 = x^3 + 5

assert pow(ZETA_P, 3, P) == 1 and ZETA_P != 1
assert pow(ZETA_Q, 3, Q) == 1 and ZETA_Q != 1
assert ROOT_P == 0x2BCE74DEAC30EBDA362120830561F81AEA322BF2B7BB7584BDAD6FABD87EA32F
assert ROOT_Q == 0x2DE6A9B8746D3F589E5C4DFD492AE26E9BB97EA3C106F049A70E2C1102B6D05F


def inv(a, m):
    return pow(a, m - 2, m)


def sqrt_mod(a, m):
    """Tonelli-Shanks; returns a root or None."""
    a %= m
    if a == 0:
        return 0
    if pow(a, (m - 1) // 2, m) != 1:
        return None
    qq, s = m - 1, 0
    while qq % 2 == 0:
        qq //= 2
        s += 1
    z = pow(GEN, qq, m)  # 5 is a non-residue generator
    mm, c, t, r = s, z, pow(a, qq, m), pow(a, (qq + 1) // 2, m)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % m
            i += 1
        b = pow(c, 1 << (mm - i - 1), m)
        mm, c = i, b * b % m
        t, r = t * c % m, r * b % m
    return r


def to_bytes(x):
    return int(x).to_bytes(32, "little")


def from_bytes(b):
    return int.from_bytes(b, "little")


def from_uniform_bytes(b64, m):
    """pasta_curves `from_uniform_bytes` / halo2 Challenge255: 64 LE bytes reduced mod m."""
    return int.from_bytes(b64, "little") % m


# ---------------------------------------------------------------- curves (affine tuples, None = identity)
class Curve:
    def __init__(self, base, scalar, name):
        self.fb, self.fs, self.name = base, scalar, name

    def is_on_curve(self, pt):
        if pt is None:
            return True
        x, y = pt
        return (y * y - x * x * x - CURVE_B) % self.fb == 0

    def neg(self, pt):
        return None if pt is None else (pt[0], (-pt[1]) % self.fb)

    def add(self, a, b):
        m = self.fb
        if a is None:
            return b
        if b is None:
            return a
        x1, y1 = a
        x2, y2 = b
        if x1 == x2:
            if (y1 + y2) % m == 0:
                return None
            lam = 3 * x1 * x1 * inv(2 * y1, m) % m
        else:
            lam = (y2 - y1) * inv(x2 - x1, m) % m
        x3 = (lam * lam - x1 - x2) % m
        return (x3, (lam * (x1 - x3) - y1) % m)

    def mul(self, k, pt):
        k %= self.fs
        acc = None
        while k:
            if k & 1:
                acc = self.add(acc, pt)
            pt = self.add(pt, pt)
            k >>= 1
        return acc

    def msm(self, scalars, points):
        acc = None
        for s, p_ in zip(scalars, points):
            acc = self.add(acc, self.mul(s, p_))
        return acc

    # pasta_curves compressed encoding: x LE, bit 255 = parity of y; identity = 32 zero bytes
    def decompress(self, b):
        v = from_bytes(b)
        sign = v >> 255
        x = v & ((1 << 255) - 1)
        if x == 0 and sign == 0:
            return None
        y = sqrt_mod((x * x * x + CURVE_B) % self.fb, self.fb)
        if y is None or x >= self.fb:
            raise ValueError("not on curve")
        if (y & 1) != sign:
            y = self.fb - y
        return (x, y)

    def compress(self, pt):
        if pt is None:
            return bytes(32)
        x, y = pt
        return to_bytes(x | ((y & 1) << 255))


VESTA = Curve(Q, P, "vesta")    # commitment group of Taiga's proofs (proof.rs:25-27)
PALLAS = Curve(P, Q, "pallas")
VESTA_GEN = (Q - 1, 2)
PALLAS_GEN = (P - 1, 2)


# ---------------------------------------------------------------- evaluation domain (halo2 poly/domain.rs, EXT)
def omega(k, m=P):
    root = ROOT_P if m == P else ROOT_Q
    return pow(root, 1 << (S - k), m)


def bitrev(i, bits):
    return int(format(i, "0%db" % bits)[::-1], 2) if bits else 0


def ntt(a, w, m=P):
    """In-order radix-2 NTT: out[k] = sum_i a[i] w^(ik)."""
    n = len(a)
    if n == 1:
        return list(a)
    bits = n.bit_length() - 1
    a = [a[bitrev(i, bits)] for i in range(n)]
    length = 2
    while length <= n:
        wl = pow(w, n // length, m)
        for s in range(0, n, length):
            t = 1
            for j in range(length // 2):
                u, v = a[s + j], a[s + j + length // 2] * t % m
                a[s + j], a[s + j + length // 2] = (u + v) % m, (u - v) % m
                t = t * wl % m
        length *= 2
    return a


def intt(a, w, m=P):
    n = len(a)
    ninv = inv(n, m)
    return [x * ninv % m for x in ntt(a, inv(w, m), m)]


def ntt_naive(a, w, m=P):
    n = len(a)
    return [sum(a[i] * pow(w, i * k, m) for i in range(n)) % m for k in range(n)]


def coeff_to_extended(coeffs, k, ext_k, m=P):
    """halo2 EvaluationDomain::coeff_to_extended: scale coeff i by zeta^(i mod 3), zero-pad, NTT(2^ext_k)."""
    zeta = ZETA_P if m == P else ZETA_Q
    zp = [1, zeta, zeta * zeta % m]
    a = [c * zp[i % 3] % m for i, c in enumerate(coeffs)] + [0] * ((1 << ext_k) - len(coeffs))
    return ntt(a, omega(ext_k, m), m)


def extended_to_coeff(evals, ext_k, m=P):
    zeta = ZETA_P if m == P else ZETA_Q
    zi = [1, zeta * zeta % m, zeta]  # inverse powers
    a = intt(evals, omega(ext_k, m), m)
    return [c * zi[i % 3] % m for i, c in enumerate(a)]


def eval_poly(coeffs, x, m=P):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % m
    return acc


# ---------------------------------------------------------------- params_15 fixture (constant.rs:128-139)
def read_params(path_or_bytes, curve=VESTA, limit=None):
    """Layout: u32 k | g[n] | g_lagrange[n] | w | u, 32-byte compressed points."""
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    k = struct.unpack("<I", data[:4])[0]
    n = 1 << k
    assert len(data) == 4 + 32 * (2 * n + 2)
    cnt = n if limit is None else limit

    def pt(i):
        return curve.decompress(data[4 + 32 * i: 4 + 32 * (i + 1)])

    g = [pt(i) for i in range(cnt)]
    gl = [pt(n + i) for i in range(cnt)]
    return {"k": k, "n": n, "g": g, "g_lagrange": gl, "w": pt(2 * n), "u": pt(2 * n + 1)}


# ---------------------------------------------------------------- BLAKE2b transcript (halo2 transcript.rs, EXT)
class Transcript:
    """Blake2bWrite<_, vesta::Affine, Challenge255>: personal 'Halo2-Transcript', 64-byte digest."""

    def __init__(self, scalar_mod=P):
        self.h = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")
        self.proof = bytearray()
        self.m = scalar_mod

    def common_point(self, pt):
        if pt is None:
            raise ValueError("cannot write points at infinity to the transcript")
        self.h.update(b"\x01" + to_bytes(pt[0]) + to_bytes(pt[1]))

    def common_scalar(self, s):
        self.h.update(b"\x02" + to_bytes(s))

    def write_point(self, pt, curve=VESTA):
        self.common_point(pt)
        self.proof += curve.compress(pt)

    def write_scalar(self, s):
        self.common_scalar(s)
        self.proof += to_bytes(s)

    def squeeze(self):
        self.h.update(b"\x00")
        return from_uniform_bytes(self.h.copy().digest(), self.m)
