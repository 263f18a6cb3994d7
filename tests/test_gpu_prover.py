"""GPU parity tests of the full prover (through tb_circuit_load / tb_prove_batch): proof bytes must equal the CPU
oracle's byte for byte for the same seed, and must be accepted by the oracle's verifier restatement."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

from taiga_b200 import circuits_mini as cm
from taiga_b200 import lib

pytestmark = pytest.mark.gpu

_SRS = {}


def small_srs(oracle_cpu, gpu_ctx, k):
    if k not in _SRS:
        s = oracle_cpu.synthetic_srs(k, seed=k)
        _SRS[k] = (s, gpu_ctx.load_srs(k, s["g"], s["g_lagrange"], s["w"], s["u"]))
    return _SRS[k]


@pytest.mark.parametrize("k,wide,nl", [(6, False, 2), (7, True, 1), (6, False, 0), (9, True, 2), (12, False, 1)])
def test_mini_circuit_proofs_bit_identical(gpu_ctx, oracle_cpu, k, wide, nl):
    kd, make = cm.standard_plonk(k=k, wide=wide, n_lookups=nl)
    srs, gsrs = small_srs(oracle_cpu, gpu_ctx, k)
    okey = oracle_cpu.OracleKey(kd, srs)
    pk = gsrs.load_circuit(kd)
    assert pk.proof_len == kd.proof_size()
    B = 3
    wit = [kd.witness_arrays(make(100 + b)) for b in range(B)]
    adv = np.stack([w[0] for w in wit])
    inst = np.stack([w[1] for w in wit])
    lens = wit[0][2]
    seed = bytes((7 * i + 1) & 0xFF for i in range(32))
    proofs = pk.prove_batch(adv, inst, lens, seed, first_proof_index=5)
    for b in range(B):
        ref = okey.prove(wit[b][0], wit[b][1], lens, seed, proof_index=5 + b)
        assert len(proofs[b]) == len(ref) == kd.proof_size()
        if proofs[b] != ref:
            first = next(i for i in range(len(ref)) if proofs[b][i] != ref[i])
            pytest.fail("proof %d differs from the oracle at byte %d (32-byte element %d)" % (b, first, first // 32))
        assert okey.verify(wit[b][1], lens, proofs[b]) == 0
    if (k, wide, nl) == (6, False, 2):   # committed golden vector (tests/golden/make_proof_fixtures.py: witness 100, proof index 5)
        assert proofs[0] == open(os.path.join(GOLDEN, "proof_k6_plonk.bin"), "rb").read()


def test_unsatisfied_lookup_is_reported(gpu_ctx, oracle_cpu):
    """halo2 returns Error::ConstraintSystemFailure when a lookup input is missing from the table."""
    kd, make = cm.standard_plonk(k=6, n_lookups=1)
    srs, gsrs = small_srs(oracle_cpu, gpu_ctx, 6)
    pk = gsrs.load_circuit(kd)
    asg = make(1)
    row = max(r for r, v in asg.fixed[6].items() if v == 1)  # a row with the lookup selector on
    asg.advice[0][row] = 999
    adv, inst, lens = kd.witness_arrays(asg)
    with pytest.raises(lib.ConstraintSystemFailure):
        pk.prove_batch(adv[None], inst[None], lens, bytes(32))
    with pytest.raises(RuntimeError):
        oracle_cpu.OracleKey(kd, srs).prove(adv, inst, lens, bytes(32))


def test_keygen_commitments_match_oracle(gpu_ctx, oracle_cpu):
    """keygen_vk on the device: fixed / sigma column commitments equal the oracle's (vk.fixed_commitments, permutation commitments)."""
    kd, _ = cm.standard_plonk(k=7, wide=True, n_lookups=1)
    srs, gsrs = small_srs(oracle_cpu, gpu_ctx, 7)
    of, os_ = oracle_cpu.OracleKey(kd, srs).commitments()
    gf, gs = gsrs.load_circuit(kd).commitments()
    assert gf.tobytes() == of.tobytes() and gs.tobytes() == os_.tobytes()


@pytest.mark.parametrize("k,wide,nl", [(6, False, 2), (7, True, 1), (6, False, 0)])
def test_device_verifier_agrees_with_oracle(gpu_ctx, oracle_cpu, k, wide, nl):
    """tb_verify_batch (product-side Proof::verify): accepts proofs made by the CUDA prover AND by the CPU oracle prover,
    rejects tampered proofs, truncated proofs and wrong public inputs - same verdicts as the oracle's verifier."""
    kd, make = cm.standard_plonk(k=k, wide=wide, n_lookups=nl)
    srs, gsrs = small_srs(oracle_cpu, gpu_ctx, k)
    okey = oracle_cpu.OracleKey(kd, srs)
    pk = gsrs.load_circuit(kd)
    wit = [kd.witness_arrays(make(200 + b)) for b in range(3)]
    adv = np.stack([w[0] for w in wit]); inst = np.stack([w[1] for w in wit]); lens = wit[0][2]
    proofs = pk.prove_batch(adv, inst, lens, bytes(range(32)))
    proofs[2] = okey.prove(wit[2][0], wit[2][1], lens, bytes(32), proof_index=77)   # a CPU-made proof (different blinding)
    assert pk.verify_batch(inst, lens, proofs) == [True, True, True]
    bad = list(proofs)
    t = bytearray(bad[0]); t[len(t) // 2] ^= 0x10; bad[0] = bytes(t)                   # flipped bit in an evaluation
    t = bytearray(bad[1]); t[5] ^= 1; bad[1] = bytes(t)                                # flipped bit in the first commitment
    verdict = pk.verify_batch(inst, lens, bad)
    assert verdict == [False, False, True]
    assert [okey.verify(wit[b][1], lens, bad[b]) == 0 for b in range(3)] == verdict
    wrong_inst = inst.copy(); wrong_inst[1, 0] ^= 1
    assert pk.verify_batch(wrong_inst, lens, proofs) == [True, False, True]
    asg = make(200); asg.advice[2][0] = (asg.advice[2][0] + 1) % cm.P                  # unsatisfying witness -> rejected proof
    adv2, _, _ = kd.witness_arrays(asg)
    p2 = pk.prove_batch(adv2[None], wit[0][1][None], lens, bytes(32))
    assert pk.verify_batch(wit[0][1][None], lens, p2) == [False]
