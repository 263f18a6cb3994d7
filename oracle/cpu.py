"""ORACLE (test infrastructure, never shipped): ctypes binding of oracle/liboracle.so, the threaded C++
CPU restatement.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FP, FQ = 0, 1
VESTA, PALLAS = 0, 1


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def set_threads(n):
    return lib().orc_set_threads(int(n))


def ints_to_bytes(xs):
    """list of Python ints -> uint8 array [len, 32] (canonical LE)."""
    return np.frombuffer(b"".join(int(x).to_bytes(32, "little") for x in xs), dtype=np.uint8).reshape(-1, 32).copy()


def bytes_to_ints(a):
    a = _u8(a).reshape(-1, 32)
    return [int.from_bytes(r.tobytes(), "little") for r in a]


def field_op(field, op, a, b=None):
    out = np.zeros(32, np.uint8)
    a = _u8(a)
    rc = lib().orc_field_op(field, op, _p(a), _p(_u8(b)) if b is not None else None, _p(out))
    return rc, out


def field_consts(field, k):
    out = np.zeros(192, np.uint8)
    lib().orc_field_consts(field, k, _p(out))
    v = bytes_to_ints(out)
    return {"omega": v[0], "delta": v[1], "zeta": v[2], "R": v[3], "R2": v[4], "inv64": v[5]}


def from_uniform(field, b64):
    out = np.zeros(32, np.uint8)
    lib().orc_from_uniform(field, _p(_u8(np.frombuffer(b64, np.uint8))), _p(out))
    return out


def ntt(field, data, inverse=False):
    """data: uint8 [n,32] canonical; returns new array."""
    d = _u8(data).copy()
    n = d.size // 32
    logn = n.bit_length() - 1
    assert 1 << logn == n
    lib().orc_ntt(field, logn, int(inverse), _p(d))
    return d.reshape(n, 32)


def coeff_to_extended(k, cs_degree, coeffs):
    n = 1 << k
    ext_k = k
    while (1 << ext_k) < n * (cs_degree - 1):
        ext_k += 1
    out = np.zeros((1 << ext_k, 32), np.uint8)
    lib().orc_coeff_to_extended(k, cs_degree, _p(_u8(coeffs)), _p(out))
    return out


def extended_to_coeff(k, cs_degree, evals):
    n = 1 << k
    out = np.zeros((n * (cs_degree - 1), 32), np.uint8)
    lib().orc_extended_to_coeff(k, cs_degree, _p(_u8(evals)), _p(out))
    return out


def msm(curve, scalars, points):
    s, p = _u8(scalars), _u8(points)
    n = s.size // 32
    assert p.size == 64 * n
    out = np.zeros(64, np.uint8)
    lib().orc_msm(curve, ctypes.c_size_t(n), _p(s), _p(p), _p(out))
    return out


def decompress(curve, enc):
    e = _u8(enc)
    n = e.size // 32
    out = np.zeros((n, 64), np.uint8)
    rc = lib().orc_decompress(curve, ctypes.c_size_t(n), _p(e), _p(out))
    if rc:
        raise ValueError("point not on curve")
    return out


def compress(curve, pts):
    p = _u8(pts)
    n = p.size // 64
    out = np.zeros((n, 32), np.uint8)
    lib().orc_compress(curve, ctypes.c_size_t(n), _p(p), _p(out))
    return out


def point_add(curve, a, b):
    out = np.zeros(64, np.uint8)
    lib().orc_point_op(curve, 0, _p(_u8(a)), _p(_u8(b)), _p(out))
    return out


def point_mul(curve, a, k_bytes):
    out = np.zeros(64, np.uint8)
    lib().orc_point_op(curve, 1, _p(_u8(a)), _p(_u8(k_bytes)), _p(out))
    return out


# ---------------------------------------------------------------- halo2 prover / verifier restatement (oracle/plonk.cpp)
class OracleKey:
    """keygen result for one circuit (CircuitKeyData from taiga_b200.circuit) over an SRS dict of affine byte arrays."""

    def __init__(self, keydata, srs):
        L = lib()
        L.orc_keygen.restype = ctypes.c_void_p
        n = keydata.n
        g, gl = _u8(srs["g"]), _u8(srs["g_lagrange"])
        assert g.size == 64 * n and gl.size == 64 * n, "SRS size must match the circuit's k"
        self.keydata, self.srs = keydata, srs
        self._fixed, self._sigma = _u8(keydata.fixed), _u8(keydata.sigma)
        self._h = ctypes.c_void_p(L.orc_keygen(ctypes.byref(keydata.desc), _p(g), _p(gl), _p(_u8(srs["w"])), _p(_u8(srs["u"])),
                                               _p(self._fixed), _p(self._sigma)))

    def prove(self, advice, instance, instance_len, seed, proof_index=0):
        L = lib()
        buf = np.zeros(1 << 16, np.uint8)
        ln = ctypes.c_size_t(buf.size)
        seed = _u8(np.frombuffer(seed, np.uint8))
        rc = L.orc_prove(self._h, _p(_u8(advice)), _p(_u8(instance)), _p(np.ascontiguousarray(instance_len, dtype=np.uint32)), _p(seed),
                         ctypes.c_uint32(proof_index), _p(buf), ctypes.byref(ln))
        if rc:
            raise RuntimeError("oracle prover failed rc=%d (2 = InstanceTooLarge, 3 = ConstraintSystemFailure)" % rc)
        return buf[: ln.value].tobytes()

    def verify(self, instance, instance_len, proof):
        pb = _u8(np.frombuffer(proof, np.uint8))
        return lib().orc_verify(self._h, _p(_u8(instance)), _p(np.ascontiguousarray(instance_len, dtype=np.uint32)), _p(pb), ctypes.c_size_t(pb.size))

    def commitments(self):
        kd = self.keydata
        f = np.zeros((max(1, kd.cs.num_fixed), 64), np.uint8)
        s = np.zeros((max(1, len(kd.cs.perm_columns)), 64), np.uint8)
        lib().orc_key_commitments(self._h, _p(f), _p(s))
        return f[: kd.cs.num_fixed], s[: len(kd.cs.perm_columns)]

    def __del__(self):
        try:
            if self._h:
                lib().orc_key_free(self._h)
                self._h = None
        except Exception:
            pass


def rnd(seed, proof, tag, idx):
    out = np.zeros(32, np.uint8)
    lib().orc_rnd(_p(_u8(np.frombuffer(seed, np.uint8))), ctypes.c_uint32(proof), ctypes.c_uint32(tag), ctypes.c_uint32(idx), _p(out))
    return int.from_bytes(out.tobytes(), "little")


def blake2b(data, personal16):
    out = np.zeros(64, np.uint8)
    d = _u8(np.frombuffer(data, np.uint8)) if len(data) else np.zeros(1, np.uint8)
    lib().orc_blake2b(_p(d), ctypes.c_size_t(len(data)), ctypes.c_char_p(personal16), _p(out))
    return out.tobytes()


def synthetic_srs(k, seed=1):
    """A structurally valid SRS for small test circuits: g[i] = [s_i] G with known s_i (insecure, test only),
    g_lagrange = group inverse-DFT of g (as in params_15), w, u random multiples."""
    from . import pasta as o
    import random
    rnd_ = random.Random(seed)
    n = 1 << k
    s = [rnd_.randrange(1, o.P) for _ in range(n)]
    sl = o.intt(s, o.omega(k))
    G = ints_to_bytes(list(o.VESTA_GEN)).reshape(64)
    g = np.stack([point_mul(VESTA, G, ints_to_bytes([v])[0]) for v in s])
    gl = np.stack([point_mul(VESTA, G, ints_to_bytes([v])[0]) for v in sl])
    w = point_mul(VESTA, G, ints_to_bytes([rnd_.randrange(1, o.P)])[0])
    u = point_mul(VESTA, G, ints_to_bytes([rnd_.randrange(1, o.P)])[0])
    return {"k": k, "n": n, "g": g, "g_lagrange": gl, "w": w, "u": u}
