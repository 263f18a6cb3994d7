"""CPU: the INDEPENDENT pure-Python verifier (oracle/verifier_py.py, written from SURVEY.md Appendix A without sharing code
with oracle/plonk.cpp) accepts the committed golden proofs and rejects tampered ones.  The golden proofs were made by the
C++ restatement and are reproduced byte for byte by the CUDA prover, so two separately written readings of halo2 --
one prover, one verifier -- agree on the transcript order, the evaluation section, the gate / permutation / lookup
identities, multiopen and the inner product argument (the stand-in for `Proof::verify`, taiga_halo2/src/proof.rs:45-54)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

from oracle import verifier_py as vp


def _columns(inst, lens):
    cols, off = [], 0
    for l in lens:
        cols.append([int.from_bytes(inst[32 * (off + i):32 * (off + i + 1)].tobytes(), "little") for i in range(int(l))])
        off += int(l)
    return cols


def test_golden_k6_proof_accepted_and_tampering_rejected(oracle_cpu):
    from taiga_b200 import circuits_mini as cm
    kd, make = cm.standard_plonk(k=6, wide=False, n_lookups=2)
    srs = oracle_cpu.synthetic_srs(6, seed=6)
    adv, inst, lens = kd.witness_arrays(make(100))
    fc, sc = oracle_cpu.OracleKey(kd, srs).commitments()
    proof = open(os.path.join(GOLDEN, "proof_k6_plonk.bin"), "rb").read()
    cols = _columns(inst, lens)
    assert vp.verify(kd, srs, fc, sc, cols, proof)
    for pos in (0, 40, 33 * 32 + 5, len(proof) - 1):      # a commitment, an evaluation, an IPA round point, the last scalar
        bad = bytearray(proof)
        bad[pos] ^= 1
        assert not vp.verify(kd, srs, fc, sc, cols, bytes(bad))
    wrong = [list(c) for c in cols]
    wrong[0][0] += 1
    assert not vp.verify(kd, srs, fc, sc, wrong, proof)
    assert not vp.verify(kd, srs, fc, sc, cols, proof[:-32])
    # a proof of ANOTHER witness made by the C++ prover is accepted too (the verifier is not fitted to one vector)
    key = oracle_cpu.OracleKey(kd, srs)
    adv2, inst2, lens2 = kd.witness_arrays(make(7))
    assert vp.verify(kd, srs, fc, sc, _columns(inst2, lens2), key.prove(adv2, inst2, lens2, bytes(range(32)), proof_index=3))


@pytest.mark.parametrize("compliance,name", [(True, "proof_k15_compliance_shape.bin"), (False, "proof_k15_vp_shape.bin")])
def test_golden_k15_taiga_shape_proofs_accepted(oracle_cpu, srs_fixture, compliance, name):
    from taiga_b200 import circuits_taiga as ct
    kd, make = ct.build(compliance)
    adv, inst, lens = kd.witness_arrays(make(41))
    fc, sc = oracle_cpu.OracleKey(kd, srs_fixture).commitments()
    proof = open(os.path.join(GOLDEN, name), "rb").read()
    cols = _columns(inst, lens)
    assert vp.verify(kd, srs_fixture, fc, sc, cols, proof)
    bad = bytearray(proof)
    bad[2000] ^= 4
    assert not vp.verify(kd, srs_fixture, fc, sc, cols, bytes(bad))
    if compliance:   # the verifying-key commitments are inputs; one of them recomputed here: commit_lagrange(fixed column 0, blind 1)
        vals = [int.from_bytes(kd.fixed[0, i].tobytes(), "little") for i in range(kd.n)]
        gl = [vp._from_affine(b) for b in srs_fixture["g_lagrange"]]
        c = vp._to_affine(vp._add(vp.msm(vals, gl), vp._from_affine(srs_fixture["w"])))
        assert c == (int.from_bytes(fc[0][:32].tobytes(), "little"), int.from_bytes(fc[0][32:].tobytes(), "little"))
