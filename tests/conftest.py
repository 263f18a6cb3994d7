import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_cpu():
    from oracle import cpu
    cpu.build()
    return cpu


@pytest.fixture(scope="session")
def srs_fixture():
    """Taiga's params_15 decompressed to affine (tests/golden/make_fixtures.py)."""
    raw = np.fromfile(os.path.join(GOLDEN, "srs_k15_affine.bin"), dtype=np.uint8)
    n = 1 << 15
    assert raw.size == 64 * (2 * n + 2)
    pts = raw.reshape(-1, 64)
    return {"k": 15, "n": n, "g": pts[:n], "g_lagrange": pts[n:2 * n], "w": pts[2 * n], "u": pts[2 * n + 1]}


@pytest.fixture(scope="session")
def gpu_ctx():
    from taiga_b200 import lib
    ctx = lib.Context(0)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def gpu_srs(gpu_ctx, srs_fixture):
    s = srs_fixture
    return gpu_ctx.load_srs(s["k"], s["g"], s["g_lagrange"], s["w"], s["u"])
