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
