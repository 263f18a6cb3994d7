"""CPU: the product's own field/curve source (taiga_b200/csrc/field.cuh, curve.cuh - the code the kernels compile)
instantiated for the host and checked against the Python big-int oracle."""
import ctypes
import os
import random
import subprocess

import pytest

from conftest import ROOT
from oracle import pasta as o


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shim") / "host_shim.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, os.path.join(ROOT, "tests", "host_shim.cpp")])
    return ctypes.CDLL(so)


def _b32(x):
    return (ctypes.c_uint8 * 32).from_buffer_copy(int(x).to_bytes(32, "little"))


def _b64(p):
    return (ctypes.c_uint8 * 64).from_buffer_copy(bytes(64) if p is None else int(p[0]).to_bytes(32, "little") + int(p[1]).to_bytes(32, "little"))


def test_field(shim):
    rnd = random.Random(5)

    def fo(f, op, a, b=0):
        out = (ctypes.c_uint8 * 32)()
        shim.hs_field(f, op, _b32(a), _b32(b), out)
        return int.from_bytes(bytes(out), "little")

    for f, m in [(0, o.P), (1, o.Q)]:
        edge = [0, 1, m - 1, m - 2, (1 << 256) % m, 2 ** 32 - 1, 2 ** 32, (1 << 255) % m, m >> 1]
        vals = edge + [rnd.randrange(m) for _ in range(200)]
        for a in vals:
            for b in rnd.sample(vals, 4) + edge[:4]:
                assert fo(f, 0, a, b) == (a + b) % m and fo(f, 1, a, b) == (a - b) % m and fo(f, 2, a, b) == a * b % m
            assert fo(f, 4, a) == (-a) % m and fo(f, 5, a) == a * a % m
        # inversion (divsteps, field.cuh::inv): every edge value, small values, powers of two and many random ones
        for a in vals + list(range(2, 40)) + [1 << i for i in range(1, 254, 7)] + [m - (1 << i) for i in range(1, 250, 11)] + [rnd.randrange(m) for _ in range(3000)]:
            assert fo(f, 3, a) == pow(a, m - 2, m), hex(a)


def test_curve(shim):
    rnd = random.Random(6)

    def po(c, op, a, b):
        out = (ctypes.c_uint8 * 64)()
        shim.hs_point(c, op, _b64(a), b if isinstance(b, ctypes.Array) else _b64(b), out)
        r = bytes(out)
        x, y = int.from_bytes(r[:32], "little"), int.from_bytes(r[32:], "little")
        return None if x == 0 and y == 0 else (x, y)

    for c, cv, G in [(0, o.VESTA, o.VESTA_GEN), (1, o.PALLAS, o.PALLAS_GEN)]:
        A, B = cv.mul(12345, G), cv.mul(99999, G)
        assert po(c, 0, A, B) == cv.add(A, B) and po(c, 0, A, A) == cv.add(A, A) and po(c, 0, A, cv.neg(A)) is None
        assert po(c, 0, A, None) == A and po(c, 0, None, B) == B
        assert po(c, 2, A, None) == cv.add(A, A) and po(c, 4, A, None) == cv.add(A, A)
        assert po(c, 3, A, B) == cv.add(cv.mul(2, A), cv.mul(3, B)) and po(c, 3, A, A) == cv.mul(5, A)
        for k in [0, 1, 2, 15, 16, cv.fs - 1, rnd.randrange(cv.fs)]:
            assert po(c, 1, A, _b32(k)) == cv.mul(k, A)
