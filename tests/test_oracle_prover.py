"""CPU: the oracle's halo2 prover/verifier restatement (oracle/plonk.cpp) on small circuits, its BLAKE2b / blinding PRF
against hashlib, and the structural facts of the Taiga-shaped circuits the reference documents (degree, proof size)."""
import hashlib

import pytest

from taiga_b200 import circuits_mini as cm
from taiga_b200 import circuits_taiga as ct
from taiga_b200.circuit import P


def test_blake2b_and_prf_match_hashlib(oracle_cpu):
    c = oracle_cpu
    for msg in (b"", b"abc", bytes(range(200)), b"x" * 128, b"y" * 129):
        assert c.blake2b(msg, b"Halo2-Transcript") == hashlib.blake2b(msg, digest_size=64, person=b"Halo2-Transcript").digest()
    seed = bytes(range(32))
    for proof, tag, idx in ((0, 1, 0), (5, 11, 32767), (1 << 20, 18, 14)):
        m = seed + proof.to_bytes(4, "little") + tag.to_bytes(4, "little") + idx.to_bytes(4, "little") + bytes(4)
        ref = int.from_bytes(hashlib.blake2b(m, digest_size=64, person=b"TaigaB200-Blind\0").digest(), "little") % P
        assert c.rnd(seed, proof, tag, idx) == ref


@pytest.mark.parametrize("k,wide,nl", [(6, False, 2), (7, True, 1), (6, False, 0)])
def test_prover_verifier_roundtrip(oracle_cpu, k, wide, nl):
    c = oracle_cpu
    kd, make = cm.standard_plonk(k=k, wide=wide, n_lookups=nl)
    key = c.OracleKey(kd, c.synthetic_srs(k))
    adv, inst, lens = kd.witness_arrays(make(5))
    seed = bytes(range(32))
    proof = key.prove(adv, inst, lens, seed)
    assert len(proof) == kd.proof_size()          # SURVEY App. D accounting
    assert key.verify(inst, lens, proof) == 0
    assert key.prove(adv, inst, lens, seed) == proof and key.prove(adv, inst, lens, bytes(32)) != proof  # seed-determined blinding
    bad = bytearray(proof); bad[100] ^= 1
    assert key.verify(inst, lens, bytes(bad)) != 0            # tampered proof
    assert key.verify(inst, lens, proof[:-32]) != 0           # truncated proof
    bad_inst = inst.copy(); bad_inst[0] ^= 1
    assert key.verify(bad_inst, lens, proof) != 0             # wrong public input
    asg = make(5)
    asg.advice[2][0] = (asg.advice[2][0] + 1) % P             # witness violating the arithmetic gate
    adv2, _, _ = kd.witness_arrays(asg)
    assert key.verify(inst, lens, key.prove(adv2, inst, lens, seed)) != 0


def test_lookup_failure_is_constraint_system_failure(oracle_cpu):
    c = oracle_cpu
    kd, make = cm.standard_plonk(k=6, n_lookups=1)
    key = c.OracleKey(kd, c.synthetic_srs(6))
    asg = make(1)
    row = max(r for r, v in asg.fixed[6].items() if v == 1)
    asg.advice[0][row] = 999                                   # not in the 0..15 table
    adv, inst, lens = kd.witness_arrays(asg)
    with pytest.raises(RuntimeError, match="rc=3"):
        key.prove(adv, inst, lens, bytes(32))


def test_taiga_shapes_structure():
    kd_c, make_c = ct.build(True)
    cs = kd_c.cs
    # 10 advice (all equality enabled) + instance + constants column = 12 permutation columns (compliance_circuit.rs:77-112)
    assert cs.num_advice == 10 and cs.num_instance == 1 and len(cs.perm_columns) == 12 and len(cs.lookups) == 1
    assert kd_c.degree == 17 and kd_c.blinding_factors == 5       # iso-map gate (curve/map_to_curve.rs:78-80); rotations -1/0/+1
    assert kd_c.proof_size() == 4480                               # taiga_api.rs:109: 4676 = 4 + 4480 + 6*32
    kd_v, _ = ct.build(False)
    assert kd_v.degree == 9 and len(kd_v.cs.perm_columns) == 12
    asg = make_c(3)
    assert len(asg.instance[0]) == 9                                # CompliancePublicInputs::to_instance, compliance.rs:62-78
    assert kd_c.shape.rows_used > (1 << 14)                         # blake2s forces k = 15


def test_golden_proofs_are_reproduced(oracle_cpu):
    """tests/golden/proof_*.bin (made by tests/golden/make_proof_fixtures.py): the oracle must still produce exactly these
    bytes for the same circuit, witness, SRS and seed, and accept them -- a change of one byte anywhere in the restated
    prover shows up here on CPU.  The GPU suites compare the CUDA prover with the same files."""
    import hashlib
    import importlib.util
    import json
    import os
    from conftest import GOLDEN
    spec = importlib.util.spec_from_file_location("make_proof_fixtures", os.path.join(GOLDEN, "make_proof_fixtures.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    meta = json.load(open(os.path.join(GOLDEN, "proofs.json")))
    seen = 0
    for name, kd, srs, (adv, inst, lens), seed, idx in mk.cases(oracle_cpu):
        want = open(os.path.join(GOLDEN, name), "rb").read()
        assert hashlib.sha256(want).hexdigest() == meta[name]["sha256"] and len(want) == meta[name]["bytes"] == kd.proof_size()
        key = oracle_cpu.OracleKey(kd, srs)
        assert key.prove(adv, inst, lens, seed, proof_index=idx) == want
        assert key.verify(inst, lens, want) == 0
        seen += 1
    assert seen == 3 and len(open(os.path.join(GOLDEN, "proof_k15_compliance_shape.bin"), "rb").read()) == 4480   # taiga_api.rs:109
