"""Generates the committed golden fixtures from the reference's own artefacts.  Run in the build container
(needs /root/reference); the outputs travel with the repo because /root/reference does not exist on the GPU box.

  srs_k15_affine.bin : Taiga's SRS /root/reference/taiga_halo2/params/params_15 (constant.rs:128-139), decompressed
                       to affine (x||y, 32-byte LE each): g[2^15] | g_lagrange[2^15] | w | u.
  srs_k15_kat.json   : the three fixture identities of SURVEY.md App. B.2 + sha256 of both files.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import cpu as c, pasta as o  # noqa: E402

SRC = "/root/reference/taiga_halo2/params/params_15"


def main():
    data = open(SRC, "rb").read()
    k = int.from_bytes(data[:4], "little")
    n = 1 << k
    assert k == 15 and len(data) == 4 + 32 * (2 * n + 2)
    pts = c.decompress(c.VESTA, np.frombuffer(data[4:], np.uint8))
    assert c.compress(c.VESTA, pts).tobytes() == data[4:]
    open(os.path.join(HERE, "srs_k15_affine.bin"), "wb").write(pts.tobytes())
    g, gl = pts[:n], pts[n:2 * n]
    om = o.omega(k)
    wv = [1]
    for _ in range(n - 1):
        wv.append(wv[-1] * om % o.P)
    kat = {
        "source": "taiga_halo2/params/params_15 @ de70468",
        "k": k,
        "sha256_params_15": hashlib.sha256(data).hexdigest(),
        "sha256_srs_k15_affine": hashlib.sha256(pts.tobytes()).hexdigest(),
        "g0_x": hex(c.bytes_to_ints(g[0][:32])[0]),
        "sum_g_lagrange_equals_g0": c.msm(c.VESTA, c.ints_to_bytes([1] * n), gl).tobytes() == g[0].tobytes(),
        "ninv_sum_g_equals_g_lagrange0": c.msm(c.VESTA, c.ints_to_bytes([o.inv(n, o.P)] * n), g).tobytes() == gl[0].tobytes(),
        "sum_omega_i_g_lagrange_i_equals_g1": c.msm(c.VESTA, c.ints_to_bytes(wv), gl).tobytes() == g[1].tobytes(),
    }
    assert kat["sum_g_lagrange_equals_g0"] and kat["ninv_sum_g_equals_g_lagrange0"] and kat["sum_omega_i_g_lagrange_i_equals_g1"]
    json.dump(kat, open(os.path.join(HERE, "srs_k15_kat.json"), "w"), indent=1)
    print(json.dumps(kat, indent=1))


if __name__ == "__main__":
    main()
