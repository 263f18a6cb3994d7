"""GPU parity tests of the full prover (through tb_circuit_load / tb_prove_batch): proof bytes must equal the CPU
oracle's byte for byte for the same seed, and must be accepted by the oracle's verifier restatement."""
import numpy as np
import pytest

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
