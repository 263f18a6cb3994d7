"""borsh wire format of the proof records either side of `Proof::create`.

Host-side only (bytes in, bytes out): what a prover service needs to hand the GPU-made proofs back to Taiga in the
layout `partial_transaction_serialize` / `partial_transaction_deserialize` use
(taiga_halo2/src/taiga_api.rs:104-131; impls in shielded_ptx.rs:272-320, resource_logic_circuit.rs:175-213,
compliance.rs:82-125, proof.rs:19-22).  borsh conventions: `Vec<T>` = u32 LE count + items, `Vec<u8>` = u32 LE length +
bytes, fixed arrays raw, `Option` = 1 tag byte; field elements are their 32-byte little-endian canonical `to_repr()`.

    Proof                        = Vec<u8>                                          (proof.rs:19-22)
    CompliancePublicInputs       = anchor ‖ nf ‖ cm ‖ delta ‖ in_rl_cm ‖ out_rl_cm  (6 x 32 B; compliance.rs:82-93)
    ComplianceVerifyingInfo      = Proof ‖ CompliancePublicInputs                   (4 + 4480 + 192 = 4676 B; taiga_api.rs:109)
    ResourceLogicVerifyingInfo   = vk bytes ‖ Proof ‖ 22 x 32 B public inputs       (resource_logic_circuit.rs:175-189)
    ResourceLogicVerifyingInfoSet= ResourceLogicVerifyingInfo ‖ Vec<ResourceLogicVerifyingInfo>   (shielded_ptx.rs:53-61)
    ShieldedPartialTransaction   = Vec<ComplianceVerifyingInfo> ‖ Vec<Set> (inputs) ‖ Vec<Set> (outputs)
                                   ‖ Option<32 B scalar> ‖ Vec<u8> hints             (shielded_ptx.rs:272-294)

The verifying key is an opaque byte string here: halo2's `VerifyingKey::write` belongs to the un-vendored fork, and its
reader is self-delimiting only together with the circuit, so decoders take the vk length as a parameter.
"""
import hashlib
import struct

P = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001  # Pallas base field = circuit field
Q = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001  # Pallas scalar field
COMPLIANCE_PUBLIC_INPUT_NUM = 9       # constant.rs:54-62
RL_PUBLIC_INPUT_NUM = 22              # constant.rs:68-75
COMPLIANCE_PROOF_LEN = 4480           # taiga_api.rs:109: 4676 = 4 + 4480 + 6 * 32
COMPLIANCE_VERIFYING_INFO_SIZE = 4676


class WireError(ValueError):
    """Malformed record (the reference returns io::ErrorKind::InvalidData)."""


def _fe_bytes(x, modulus=P):
    if not 0 <= x < modulus:
        raise WireError("field element out of range")
    return int(x).to_bytes(32, "little")


def _fe_read(b, modulus=P, what="field element"):
    x = int.from_bytes(b, "little")
    if x >= modulus:
        raise WireError("%s not in field" % what)
    return x


class _Reader:
    def __init__(self, data):
        self.d, self.o = memoryview(bytes(data)), 0

    def take(self, n):
        if n < 0 or self.o + n > len(self.d):
            raise WireError("unexpected end of input")
        out = bytes(self.d[self.o:self.o + n])
        self.o += n
        return out

    def u32(self):
        return struct.unpack("<I", self.take(4))[0]

    def u8(self):
        return self.take(1)[0]

    def done(self):
        return self.o == len(self.d)


# ---------------------------------------------------------------- Proof
def encode_proof(proof):
    return struct.pack("<I", len(proof)) + bytes(proof)


def _read_proof(r):
    return r.take(r.u32())


# ---------------------------------------------------------------- Pallas point decoding (for delta -> instance rows 3, 4)
def _sqrt_fp(a):
    """Tonelli-Shanks in Fp (p - 1 = 2^32 * t)."""
    a %= P
    if a == 0:
        return 0
    if pow(a, (P - 1) // 2, P) != 1:
        return None
    s, t = 32, (P - 1) >> 32
    z = pow(5, t, P)  # 5 is a non-residue (multiplicative generator of Fp*)
    x, b, m = pow(a, (t + 1) // 2, P), pow(a, t, P), s
    while b != 1:
        i, b2 = 0, b
        while b2 != 1:
            b2 = b2 * b2 % P
            i += 1
        w = pow(z, 1 << (m - i - 1), P)
        x, z = x * w % P, w * w % P
        b, m = b * z % P, i
    return x


def decompress_pallas(b):
    """32-byte pasta_curves encoding -> affine (x, y) on y^2 = x^3 + 5 over Fp; identity (all zero bytes) -> (0, 0)."""
    v = int.from_bytes(b, "little")
    sign, x = v >> 255, v & ((1 << 255) - 1)
    if x == 0 and sign == 0:
        return (0, 0)
    if x >= P:
        raise WireError("delta not in field")
    y = _sqrt_fp((x * x % P * x + 5) % P)
    if y is None:
        raise WireError("delta is not a curve point")
    if (y & 1) != sign:
        y = P - y
    return (x, y)


def compress_pallas(pt):
    x, y = pt
    if x == 0 and y == 0:
        return bytes(32)
    return (x | ((y & 1) << 255)).to_bytes(32, "little")


# ---------------------------------------------------------------- ResourceLogicCommitment
def resource_logic_commitment(resource_logic, rcm):
    """ResourceLogicCommitment::commit (resource_logic_commitment.rs:18-27): BLAKE2s-256, personalisation "VPCommit", over the
    32-byte little-endian encodings of the resource-logic verifying-key hash and the commitment randomness (both Fp)."""
    return hashlib.blake2s(_fe_bytes(resource_logic) + _fe_bytes(rcm), digest_size=32, person=b"VPCommit").digest()


# ---------------------------------------------------------------- CompliancePublicInputs
class CompliancePublicInputs:
    """compliance.rs:50-60: anchor, nf, cm are Fp; delta a Pallas point; the two resource-logic commitments are 32 raw bytes."""

    def __init__(self, anchor, nf, cm, delta, input_rl_cm, output_rl_cm):
        self.anchor, self.nf, self.cm = int(anchor), int(nf), int(cm)
        self.delta = bytes(delta) if isinstance(delta, (bytes, bytearray)) else compress_pallas(delta)
        self.input_rl_cm, self.output_rl_cm = bytes(input_rl_cm), bytes(output_rl_cm)
        if len(self.delta) != 32 or len(self.input_rl_cm) != 32 or len(self.output_rl_cm) != 32:
            raise WireError("commitments are 32 bytes")

    def to_bytes(self):
        """Wire order (compliance.rs:82-93) -- NOT the instance order."""
        return (_fe_bytes(self.anchor) + _fe_bytes(self.nf) + _fe_bytes(self.cm) + self.delta + self.input_rl_cm + self.output_rl_cm)

    @classmethod
    def from_bytes(cls, b):
        if len(b) != 192:
            raise WireError("CompliancePublicInputs is 192 bytes")
        anchor = _fe_read(b[0:32], P, "anchor")
        nf = _fe_read(b[32:64], P, "nf")
        cm = _fe_read(b[64:96], P, "cm")
        decompress_pallas(b[96:128])  # rejects non-points like DeltaCommitment::from_bytes
        return cls(anchor, nf, cm, b[96:128], b[128:160], b[160:192])

    @staticmethod
    def _rl_halves(c):
        """ResourceLogicCommitment::to_public_inputs (resource_logic_commitment.rs:41-45): two 128-bit halves."""
        return [int.from_bytes(c[0:16], "little"), int.from_bytes(c[16:32], "little")]

    def to_instance(self):
        """The 9 instance-column values in row order (compliance.rs:62-78, constant.rs:54-62)."""
        dx, dy = decompress_pallas(self.delta)
        return [self.nf, self.anchor, self.cm, dx, dy] + self._rl_halves(self.input_rl_cm) + self._rl_halves(self.output_rl_cm)


def encode_compliance_verifying_info(proof, public_inputs):
    return encode_proof(proof) + public_inputs.to_bytes()


def _read_compliance_verifying_info(r):
    proof = _read_proof(r)
    return proof, CompliancePublicInputs.from_bytes(r.take(192))


# ---------------------------------------------------------------- ResourceLogicVerifyingInfo(+Set)
def encode_rl_verifying_info(vk_bytes, proof, public_inputs):
    if len(public_inputs) != RL_PUBLIC_INPUT_NUM:
        raise WireError("a resource-logic proof has %d public inputs" % RL_PUBLIC_INPUT_NUM)
    return bytes(vk_bytes) + encode_proof(proof) + b"".join(_fe_bytes(x) for x in public_inputs)


def _read_rl_verifying_info(r, vk_len):
    vk = r.take(vk_len)
    proof = _read_proof(r)
    pis = [_fe_read(r.take(32), P, "public input") for _ in range(RL_PUBLIC_INPUT_NUM)]
    return vk, proof, pis


def encode_rl_set(app, dynamic=()):
    """app / dynamic items: (vk_bytes, proof, public_inputs)."""
    out = encode_rl_verifying_info(*app) + struct.pack("<I", len(dynamic))
    for d in dynamic:
        out += encode_rl_verifying_info(*d)
    return out


def _read_rl_set(r, vk_len):
    app = _read_rl_verifying_info(r, vk_len)
    return app, [_read_rl_verifying_info(r, vk_len) for _ in range(r.u32())]


# ---------------------------------------------------------------- ShieldedPartialTransaction
def encode_ptx(compliances, inputs, outputs, binding_sig_r=None, hints=b""):
    """compliances: [(proof, CompliancePublicInputs)], inputs / outputs: [(app, [dynamic...])] (see encode_rl_set)."""
    out = struct.pack("<I", len(compliances))
    for proof, pi in compliances:
        out += encode_compliance_verifying_info(proof, pi)
    for sets in (inputs, outputs):
        out += struct.pack("<I", len(sets))
        for app, dyn in sets:
            out += encode_rl_set(app, dyn)
    out += b"\x00" if binding_sig_r is None else b"\x01" + _fe_bytes(binding_sig_r, Q)
    return out + struct.pack("<I", len(hints)) + bytes(hints)


def decode_ptx(data, vk_len):
    r = _Reader(data)
    compliances = [_read_compliance_verifying_info(r) for _ in range(r.u32())]
    inputs = [_read_rl_set(r, vk_len) for _ in range(r.u32())]
    outputs = [_read_rl_set(r, vk_len) for _ in range(r.u32())]
    tag = r.u8()
    binding_sig_r = None if tag == 0 else _fe_read(r.take(32), Q, "binding_sig_r")
    hints = r.take(r.u32())
    if not r.done():
        raise WireError("trailing bytes")
    return {"compliances": compliances, "inputs": inputs, "outputs": outputs, "binding_sig_r": binding_sig_r, "hints": hints}


def ptx_size(n_compliance, compliance_proof_len, rl_counts, vk_len, rl_proof_len, hints_len=0, with_binding_sig=False):
    """Byte size of a serialized ptx; rl_counts = number of dynamic proofs of each input/output resource (4 entries for 2-in/2-out)."""
    rl_info = vk_len + 4 + rl_proof_len + 32 * RL_PUBLIC_INPUT_NUM
    size = 4 + n_compliance * (4 + compliance_proof_len + 192) + 8
    size += sum(rl_info * (1 + d) + 4 for d in rl_counts)
    return size + (33 if with_binding_sig else 1) + 4 + hints_len
