"""borsh wire format of the proof records: sizes the reference states (taiga_api.rs:104-131), round trips,
instance ordering (compliance.rs:62-78) and the error behaviour of the reference's deserialisers (InvalidData)."""
import random
import struct

import pytest

from oracle import pasta
from taiga_b200 import wire


def _rand_pi(rng):
    pt = pasta.PALLAS.mul(rng.randrange(1, pasta.Q), pasta.PALLAS_GEN)
    return wire.CompliancePublicInputs(rng.randrange(pasta.P), rng.randrange(pasta.P), rng.randrange(pasta.P), pt,
                                       bytes(rng.randrange(256) for _ in range(32)), bytes(rng.randrange(256) for _ in range(32))), pt


def _rl(rng, vk_len, proof_len):
    return (bytes(rng.randrange(256) for _ in range(vk_len)), bytes(rng.randrange(256) for _ in range(proof_len)),
            [rng.randrange(pasta.P) for _ in range(wire.RL_PUBLIC_INPUT_NUM)])


def test_compliance_verifying_info_is_4676_bytes():
    rng = random.Random(1)
    pi, _ = _rand_pi(rng)
    rec = wire.encode_compliance_verifying_info(bytes(wire.COMPLIANCE_PROOF_LEN), pi)
    assert len(rec) == wire.COMPLIANCE_VERIFYING_INFO_SIZE == 4676          # taiga_api.rs:109
    assert rec[:4] == struct.pack("<I", 4480)


def test_public_inputs_wire_order_vs_instance_order():
    rng = random.Random(2)
    pi, pt = _rand_pi(rng)
    b = pi.to_bytes()
    assert len(b) == 192
    assert int.from_bytes(b[0:32], "little") == pi.anchor and int.from_bytes(b[32:64], "little") == pi.nf   # anchor first on the wire
    inst = pi.to_instance()
    assert len(inst) == wire.COMPLIANCE_PUBLIC_INPUT_NUM
    assert inst[0] == pi.nf and inst[1] == pi.anchor and inst[2] == pi.cm                                    # nf first in the instance column
    assert (inst[3], inst[4]) == pt
    assert inst[5] == int.from_bytes(pi.input_rl_cm[:16], "little") and inst[6] == int.from_bytes(pi.input_rl_cm[16:], "little")
    assert inst[7] == int.from_bytes(pi.output_rl_cm[:16], "little") and inst[8] == int.from_bytes(pi.output_rl_cm[16:], "little")
    back = wire.CompliancePublicInputs.from_bytes(b)
    assert back.to_bytes() == b


def test_pallas_point_codec_matches_oracle():
    rng = random.Random(3)
    for _ in range(20):
        pt = pasta.PALLAS.mul(rng.randrange(1, pasta.Q), pasta.PALLAS_GEN)
        enc = wire.compress_pallas(pt)
        assert enc == pasta.PALLAS.compress(pt)
        assert wire.decompress_pallas(enc) == pt == pasta.PALLAS.decompress(enc)
    assert wire.decompress_pallas(bytes(32)) == (0, 0)
    for a in (0, 1, 4, 5, 25, pasta.P - 1):
        r = wire._sqrt_fp(a)
        assert (r is None) == (pasta.sqrt_mod(a, pasta.P) is None)
        if r is not None:
            assert r * r % pasta.P == a


def test_ptx_round_trip_and_size():
    rng = random.Random(4)
    vk_len, c_len, v_len = 1000, 4480, 4448
    comps = [(bytes(rng.randrange(256) for _ in range(c_len)), _rand_pi(rng)[0]) for _ in range(2)]
    dyn = (0, 1, 2, 0)
    sets = [(_rl(rng, vk_len, v_len), [_rl(rng, vk_len, v_len) for _ in range(d)]) for d in dyn]
    for sig, hints in ((None, b""), (rng.randrange(pasta.Q), b"hint bytes")):
        blob = wire.encode_ptx(comps, sets[:2], sets[2:], sig, hints)
        assert len(blob) == wire.ptx_size(2, c_len, dyn, vk_len, v_len, len(hints), sig is not None)
        d = wire.decode_ptx(blob, vk_len)
        assert [p for p, _ in d["compliances"]] == [p for p, _ in comps]
        assert [pi.to_bytes() for _, pi in d["compliances"]] == [pi.to_bytes() for _, pi in comps]
        got = d["inputs"] + d["outputs"]
        for (app, dl), (gapp, gdl) in zip(sets, got):
            assert tuple(gapp) == app and [tuple(x) for x in gdl] == dl
        assert d["binding_sig_r"] == sig and d["hints"] == hints
        assert wire.encode_ptx(d["compliances"], d["inputs"], d["outputs"], d["binding_sig_r"], d["hints"]) == blob


def test_reference_record_sizes():
    # taiga_api.rs:109-110: ComplianceVerifyingInfo 4676 B, ResourceLogicVerifyingInfo 158216 B.  The second implies the
    # size of halo2's vk.write output for the VP circuit once the proof length is known: vk = 158216 - 4 - proof - 22*32.
    rl_info = 158216
    for proof_len in (4448, 4480):
        vk_len = rl_info - 4 - proof_len - 32 * wire.RL_PUBLIC_INPUT_NUM
        assert len(wire.encode_rl_verifying_info(bytes(vk_len), bytes(proof_len), [0] * 22)) == rl_info
    # 2-in / 2-out ptx without dynamic proofs (the layout table of taiga_api.rs:104-122, plus borsh's Vec length prefixes)
    assert wire.ptx_size(2, 4480, (0, 0, 0, 0), 158216 - 4 - 4448 - 704, 4448) == 4 + 2 * 4676 + 8 + 4 * (158216 + 4) + 1 + 4


def test_malformed_records_are_rejected():
    rng = random.Random(5)
    pi, _ = _rand_pi(rng)
    good = pi.to_bytes()
    for off in (0, 32, 64):   # anchor, nf, cm not in field
        bad = good[:off] + (pasta.P).to_bytes(32, "little") + good[off + 32:]
        with pytest.raises(wire.WireError):
            wire.CompliancePublicInputs.from_bytes(bad)
    # delta: x with no square root on the curve
    x = next(x for x in range(2, 100) if pasta.sqrt_mod((x ** 3 + 5) % pasta.P, pasta.P) is None)
    with pytest.raises(wire.WireError):
        wire.CompliancePublicInputs.from_bytes(good[:96] + x.to_bytes(32, "little") + good[128:])
    blob = wire.encode_ptx([(b"\x01" * 10, pi)], [], [], None, b"")
    with pytest.raises(wire.WireError):
        wire.decode_ptx(blob[:-1], 0)
    with pytest.raises(wire.WireError):
        wire.decode_ptx(blob + b"\x00", 0)
    with pytest.raises(wire.WireError):
        wire.encode_rl_verifying_info(b"", b"", [0] * 21)
    with pytest.raises(wire.WireError):
        wire.encode_ptx([], [], [], pasta.Q, b"")


def test_resource_logic_commitment_vector():
    # the Blake2s chip equals blake2s_simd (blake2s.rs:1176-1210), so VPCommit vectors are regenerable;
    # VPCommit(1, 1) as computed there
    c = wire.resource_logic_commitment(1, 1)
    assert c.hex() == "f8ad5e9e7bd488260e34366e3d00e29c3fd242d7e97f5237b698ca0f0e58dd9f"
    lo, hi = wire.CompliancePublicInputs._rl_halves(c)
    assert lo < (1 << 128) and hi < (1 << 128) and lo | (hi << 128) == int.from_bytes(c, "little")
