"""GPU parity at the reference's real size (k = 15, Taiga's own SRS fixture): the Compliance-shaped (degree 17,
extended domain 2^19, 4480-byte proofs as in taiga_api.rs:109) and the Resource-Logic-shaped (degree 9) circuits.
Proofs from the CUDA prover must equal the CPU oracle's byte for byte and be accepted by its verifier."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

from taiga_b200 import circuits_taiga as ct

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("compliance", [True, False])
def test_taiga_shape_proofs_bit_identical(gpu_ctx, gpu_srs, oracle_cpu, srs_fixture, compliance):
    kd, make = ct.build(compliance)
    if compliance:
        assert kd.degree == 17 and kd.proof_size() == 4480  # proof size published by the reference (taiga_api.rs:109)
    else:
        assert kd.degree == 9
    okey = oracle_cpu.OracleKey(kd, srs_fixture)
    pk = gpu_srs.load_circuit(kd)
    assert pk.proof_len == kd.proof_size()
    B = 2
    wit = [kd.witness_arrays(make(40 + b)) for b in range(B)]
    adv = np.stack([w[0] for w in wit])
    inst = np.stack([w[1] for w in wit])
    lens = wit[0][2]
    seed = bytes(range(100, 132))
    proofs = pk.prove_batch(adv, inst, lens, seed, first_proof_index=0)
    for b in range(B):
        assert okey.verify(wit[b][1], lens, proofs[b]) == 0, "oracle verifier rejected the GPU proof"
    ref = okey.prove(wit[1][0], wit[1][1], lens, seed, proof_index=1)
    if proofs[1] != ref:
        first = next(i for i in range(len(ref)) if proofs[1][i] != ref[i])
        pytest.fail("GPU proof differs from the oracle at byte %d (element %d)" % (first, first // 32))
    # ... and the committed golden vector of the same inputs (tests/golden/make_proof_fixtures.py: witness 41, proof index 1)
    golden = open(os.path.join(GOLDEN, "proof_k15_compliance_shape.bin" if compliance else "proof_k15_vp_shape.bin"), "rb").read()
    assert proofs[1] == golden
    # tampering is rejected
    bad = bytearray(proofs[0]); bad[40] ^= 1
    assert okey.verify(wit[0][1], lens, bytes(bad)) != 0
    # the independent pure-Python verifier (oracle/verifier_py.py, no code shared with the C++ restatement) accepts the GPU proof too
    from oracle import verifier_py as vp
    fc, sc = okey.commitments()
    cols, off = [], 0
    for l in lens:
        cols.append([int.from_bytes(wit[0][1][32 * (off + i):32 * (off + i + 1)].tobytes(), "little") for i in range(int(l))])
        off += int(l)
    assert vp.verify(kd, srs_fixture, fc, sc, cols, proofs[0])
    # the product's own batched verifier gives the same verdicts
    assert pk.verify_batch(inst, lens, proofs) == [True, True]
    assert pk.verify_batch(inst, lens, [bytes(bad), ref]) == [False, True]
