"""Golden PROOF vectors: what the CPU oracle (oracle/plonk.cpp) produces for fixed circuits, witnesses and seeds.

They do not come from the Rust reference (no toolchain; its proofs are randomised by OsRng anyway) -- they freeze the
restated algorithm, so that (a) any later change to the oracle that alters a single proof byte is caught on CPU
(tests/test_oracle_prover.py) and (b) the CUDA prover is compared against committed bytes as well as against the live
oracle (tests/test_gpu_prover.py, tests/test_gpu_taiga_shapes.py use exactly the inputs below).

  proof_k6_plonk.bin            mini circuit standard_plonk(k=6, n_lookups=2), witness 100, synthetic SRS (seed 6), proof index 5
  proof_k15_compliance_shape.bin Compliance-shaped circuit (degree 17, 4480 B) over Taiga's params_15, witness 41, proof index 1
  proof_k15_vp_shape.bin        Resource-Logic-shaped circuit (degree 9) over params_15, witness 41, proof index 1
  proofs.json                   sizes and sha256
Run:  python tests/golden/make_proof_fixtures.py     (CPU only, ~1 min)
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

MINI_SEED = bytes((7 * i + 1) & 0xFF for i in range(32))
TAIGA_SEED = bytes(range(100, 132))


def srs15():
    raw = np.fromfile(os.path.join(HERE, "srs_k15_affine.bin"), dtype=np.uint8).reshape(-1, 64)
    n = 1 << 15
    return {"k": 15, "n": n, "g": raw[:n], "g_lagrange": raw[n:2 * n], "w": raw[2 * n], "u": raw[2 * n + 1]}


def cases(oracle_cpu):
    """name -> (key data, srs, (advice, instance, lens), seed, proof index)"""
    from taiga_b200 import circuits_mini as cm, circuits_taiga as ct
    kd, make = cm.standard_plonk(k=6, wide=False, n_lookups=2)
    yield "proof_k6_plonk.bin", kd, oracle_cpu.synthetic_srs(6, seed=6), kd.witness_arrays(make(100)), MINI_SEED, 5
    s = srs15()
    for name, compliance in (("proof_k15_compliance_shape.bin", True), ("proof_k15_vp_shape.bin", False)):
        kd, make = ct.build(compliance)
        yield name, kd, s, kd.witness_arrays(make(41)), TAIGA_SEED, 1


def main():
    from oracle import cpu as oc
    oc.build()
    meta = {}
    for name, kd, srs, (adv, inst, lens), seed, idx in cases(oc):
        key = oc.OracleKey(kd, srs)
        proof = key.prove(adv, inst, lens, seed, proof_index=idx)
        assert key.verify(inst, lens, proof) == 0 and len(proof) == kd.proof_size()
        open(os.path.join(HERE, name), "wb").write(proof)
        meta[name] = {"circuit": kd.name, "bytes": len(proof), "sha256": hashlib.sha256(proof).hexdigest(), "proof_index": idx}
        print(name, meta[name])
    json.dump(meta, open(os.path.join(HERE, "proofs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
