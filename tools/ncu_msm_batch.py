import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taiga_b200 import lib
N15 = 1 << 15
raw = np.fromfile(os.path.join(ROOT, "tests", "golden", "srs_k15_affine.bin"), dtype=np.uint8).reshape(-1, 64)
ctx = lib.Context(0)
srs = ctx.load_srs(15, raw[:N15], raw[N15:2 * N15], raw[2 * N15], raw[2 * N15 + 1])
rng = np.random.default_rng(1)
K = int(os.environ.get("K", 296))
s = rng.integers(0, 256, size=(K, N15, 32), dtype=np.uint8); s[:, :, 31] &= 0x3F
bl = np.zeros((K, 32), np.uint8)
os.environ.setdefault("TB_MSM_BA_MIN_TERMS", "0")
srs.commit(s, bl, lagrange=True, batch=K)
