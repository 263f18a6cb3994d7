"""GPU parity tests for the NTT / MSM kernels, through the C ABI, against the CPU oracle (bit-exact) and the
reference's SRS fixture."""
import random

import numpy as np
import pytest

from oracle import pasta as o
from taiga_b200 import lib

pytestmark = pytest.mark.gpu


def rand_scalars(rnd, n, m, kind="uniform"):
    if kind == "uniform":
        return [rnd.randrange(m) for _ in range(n)]
    if kind == "witness":  # 30% zero, 30% one, 20% < 2^8, 8% < 2^32, 12% uniform
        out = []
        for _ in range(n):
            r = rnd.random()
            out.append(0 if r < .3 else 1 if r < .6 else rnd.randrange(256) if r < .8 else rnd.randrange(1 << 32) if r < .88 else rnd.randrange(m))
        return out
    if kind == "same":
        v = rnd.randrange(m)
        return [v] * n
    if kind == "edge":
        base = [0, 1, 2, m - 1, m - 2, (1 << 254), (1 << 255) % m, (1 << 128) - 1, 0x8000, 0x7FFF, 0x8001, 0xFFFF, 0x10000]
        return [base[i % len(base)] for i in range(n)]
    raise ValueError(kind)


def make_points(c, curve, n, rnd, with_specials=True):
    """n affine points: a short random-walk of oracle point additions (distinct points), plus identity / repeats / negations."""
    cv = o.VESTA if curve == c.VESTA else o.PALLAS
    G = o.VESTA_GEN if curve == c.VESTA else o.PALLAS_GEN
    base = c.ints_to_bytes(list(cv.mul(rnd.randrange(cv.fs), G))).reshape(64)
    step = c.ints_to_bytes(list(cv.mul(rnd.randrange(cv.fs), G))).reshape(64)
    pts = np.zeros((n, 64), np.uint8)
    cur = base
    for i in range(n):
        pts[i] = cur
        cur = c.point_add(curve, cur, step)
    if with_specials and n >= 8:
        pts[3] = 0                      # identity
        pts[5] = pts[4]                 # duplicate
        neg = c.bytes_to_ints(pts[6].reshape(2, 32))
        pts[7] = c.ints_to_bytes([neg[0], (-neg[1]) % cv.fb]).reshape(64)  # negation of its neighbour
    return pts


def test_mont_roundtrip(gpu_ctx):
    import torch
    rnd = random.Random(10)
    for f, m in [(lib.TB_FP, o.P), (lib.TB_FQ, o.Q)]:
        vals = [0, 1, m - 1, (1 << 256) % m] + [rnd.randrange(m) for _ in range(1000)]
        host = np.frombuffer(b"".join(v.to_bytes(32, "little") for v in vals), np.uint8).copy()
        t = torch.from_numpy(host).cuda()
        gpu_ctx.dev_to_mont(f, t, len(vals))
        gpu_ctx.sync()
        mont = t.cpu().numpy().reshape(-1, 32)
        for i in (0, 1, 2, 3, 17):
            assert int.from_bytes(mont[i].tobytes(), "little") == vals[i] * (1 << 256) % m
        gpu_ctx.dev_from_mont(f, t, len(vals))
        gpu_ctx.sync()
        assert t.cpu().numpy().tobytes() == host.tobytes()


@pytest.mark.parametrize("logn", [1, 2, 3, 5, 8, 10, 11, 12, 13, 14, 15, 16, 17, 18])
def test_ntt_matches_oracle(gpu_ctx, oracle_cpu, logn):
    c = oracle_cpu
    rnd = random.Random(100 + logn)
    n = 1 << logn
    for f, m in [(lib.TB_FP, o.P), (lib.TB_FQ, o.Q)]:
        if f == lib.TB_FQ and logn not in (3, 12, 15):
            continue
        x = c.ints_to_bytes(rand_scalars(rnd, n, m))
        x[0] = 0
        x[n - 1] = np.frombuffer((m - 1).to_bytes(32, "little"), np.uint8)
        fwd = gpu_ctx.ntt(f, x)
        assert fwd.tobytes() == c.ntt(f, x).tobytes()
        inv = gpu_ctx.ntt(f, x, inverse=True)
        assert inv.tobytes() == c.ntt(f, x, inverse=True).tobytes()
        assert gpu_ctx.ntt(f, fwd, inverse=True).tobytes() == x.tobytes()


def test_ntt_batch_and_coset(gpu_ctx, oracle_cpu):
    c = oracle_cpu
    rnd = random.Random(7)
    for logn in (4, 9, 15):
        n = 1 << logn
        B = 3
        xs = [c.ints_to_bytes(rand_scalars(rnd, n, o.P)) for _ in range(B)]
        got = gpu_ctx.ntt(lib.TB_FP, np.concatenate(xs), batch=B).reshape(B, n, 32)
        for b in range(B):
            assert got[b].tobytes() == c.ntt(c.FP, xs[b]).tobytes()
        # halo2 zeta-coset: forward == NTT of coefficients scaled by zeta^(i mod 3); inverse undoes it
        zp = [1, o.ZETA_P, o.ZETA_P * o.ZETA_P % o.P]
        a = c.bytes_to_ints(xs[0])
        scaled = c.ints_to_bytes([v * zp[i % 3] % o.P for i, v in enumerate(a)])
        cos = gpu_ctx.ntt(lib.TB_FP, xs[0], coset=True)
        assert cos.tobytes() == c.ntt(c.FP, scaled).tobytes()
        assert gpu_ctx.ntt(lib.TB_FP, cos, inverse=True, coset=True).tobytes() == xs[0].tobytes()


@pytest.mark.parametrize("logn", [20, 22, 23])
def test_ntt_large_properties(gpu_ctx, oracle_cpu, logn):
    """BASELINE sweep sizes: bit-exact against the threaded oracle, and inverse(forward(x)) == x."""
    c = oracle_cpu
    n = 1 << logn
    rng = np.random.default_rng(logn)
    x = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    x[:, 31] &= 0x3F  # < 2^254 < p: canonical
    y = gpu_ctx.ntt(lib.TB_FP, x)
    assert y.tobytes() == c.ntt(c.FP, x).tobytes()
    assert gpu_ctx.ntt(lib.TB_FP, y, inverse=True).tobytes() == x.tobytes()


@pytest.mark.parametrize("curve", [lib.TB_VESTA, lib.TB_PALLAS])
@pytest.mark.parametrize("n,kind,window", [(1, "uniform", 0), (2, "edge", 4), (33, "edge", 5), (1000, "uniform", 0), (1000, "witness", 7),
                                           (4096, "same", 0), (4096, "uniform", 13), (1 << 15, "uniform", 0), (1 << 15, "witness", 0),
                                           (1 << 16, "uniform", 16), (1 << 16, "witness", 0)])
def test_msm_matches_oracle(gpu_ctx, oracle_cpu, curve, n, kind, window):
    c = oracle_cpu
    rnd = random.Random(n * 31 + window)
    m = o.P if curve == lib.TB_VESTA else o.Q
    pts = make_points(c, curve, n, rnd)
    sc = c.ints_to_bytes(rand_scalars(rnd, n, m, kind))
    got = gpu_ctx.msm(curve, sc, pts, window_bits=window)
    assert got[0].tobytes() == c.msm(curve, sc, pts).tobytes()


def test_msm_batch(gpu_ctx, oracle_cpu):
    c = oracle_cpu
    rnd = random.Random(77)
    n, B = 3000, 4
    pts = make_points(c, c.VESTA, n, rnd)
    kinds = ["uniform", "witness", "same", "edge"]
    scs = [c.ints_to_bytes(rand_scalars(rnd, n, o.P, k)) for k in kinds]
    scs[2][:] = 0  # an all-zero vector commits to the identity
    got = gpu_ctx.msm(lib.TB_VESTA, np.concatenate(scs), pts, batch=B)
    for b in range(B):
        assert got[b].tobytes() == c.msm(c.VESTA, scs[b], pts).tobytes()
    assert got[2].tobytes() == bytes(64)


@pytest.mark.parametrize("logn", [18, 20, 22])
def test_msm_large(gpu_ctx, oracle_cpu, srs_fixture, logn):
    """BASELINE sweep sizes (variable-base Pippenger), bases = the SRS points repeated with a twist; bit-exact vs oracle."""
    c = oracle_cpu
    n = 1 << logn
    reps = n // srs_fixture["n"]
    pts = np.concatenate([srs_fixture["g"], srs_fixture["g_lagrange"]] * ((reps + 1) // 2))[:n]
    rng = np.random.default_rng(logn)
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    sc[:, 31] &= 0x3F
    got = gpu_ctx.msm(lib.TB_VESTA, sc, pts)
    assert got[0].tobytes() == c.msm(c.VESTA, sc, pts).tobytes()


def test_srs_commit_and_fixture_identities(gpu_ctx, gpu_srs, oracle_cpu, srs_fixture):
    """Params::commit / commit_lagrange on the fixed-base path, and the identity MSM(v, g_lagrange) == MSM(iNTT(v), g)
    that ties the CUDA NTT + MSM to the reference's own fixture."""
    c, s = oracle_cpu, srs_fixture
    n = s["n"]
    rnd = random.Random(5)
    ones = c.ints_to_bytes([1] * n)
    assert gpu_srs.commit(ones, lagrange=True)[0].tobytes() == s["g"][0].tobytes()
    v = c.ints_to_bytes(rand_scalars(rnd, n, o.P))
    wit = c.ints_to_bytes(rand_scalars(rnd, n, o.P, "witness"))
    blinds = c.ints_to_bytes([rnd.randrange(o.P), 0])
    got = gpu_srs.commit(np.concatenate([v, wit]), blinds=blinds, lagrange=True, batch=2)
    for b, vec in enumerate((v, wit)):
        ref = c.msm(c.VESTA, vec, s["g_lagrange"])
        ref = c.point_add(c.VESTA, ref, c.point_mul(c.VESTA, s["w"], blinds[b]))
        assert got[b].tobytes() == ref.tobytes()
    coeffs = gpu_ctx.ntt(lib.TB_FP, v, inverse=True)
    assert gpu_srs.commit(coeffs, lagrange=False)[0].tobytes() == gpu_srs.commit(v, lagrange=True)[0].tobytes()
    assert gpu_srs.commit(coeffs, lagrange=False)[0].tobytes() == c.msm(c.VESTA, coeffs, s["g"]).tobytes()
