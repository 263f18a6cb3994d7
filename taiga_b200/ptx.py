"""Host-side mirror of the reference's proving surface for the hot path.

  reference (taiga_halo2)                                       here
  ------------------------------------------------------------  -------------------------------------------
  Proof::create(pk, params, circuit, instance, rng)  proof.rs:25-42      Proof.create(pk, witness, seed)
  Proof::verify(vk, params, instance)                proof.rs:45-54      (oracle verifier in tests / bench only)
  SETUP_PARAMS_MAP / COMPLIANCE_PROVING_KEY          constant.rs:128-152 ProverService (SRS + both proving keys, device resident)
  ShieldedPartialTransaction::build                  shielded_ptx.rs:98-134
      2 x ComplianceVerifyingInfo::create + 4 x get_verifying_info, sequential
                                                                         ProverService.build_ptx_batch: all 2P Compliance proofs
                                                                         in one batched call, all 4P VP proofs in another

The circuits are the Taiga-shaped ones of circuits_taiga.py (the real ones need the Rust `synthesize`).  Witness
synthesis happens on the host before the call, exactly as `Circuit::synthesize` does in the reference; it is not part
of the proving hot path and not part of any timed region.
"""
import threading

import numpy as np

from . import circuits_taiga, lib

COMPLIANCE_PER_PTX = 2   # shielded_ptx.rs:107-113
VP_PER_PTX = 4           # taiga_api.rs:256-352 (ptx_example_test: 4 trivial application VPs)


class Proof:
    """`Proof(Vec<u8>)` (proof.rs:21)."""

    def __init__(self, data):
        self.data = bytes(data)

    @classmethod
    def create(cls, pk, advice, instance, instance_len, seed):
        """One proof (the reference's call shape).  pk: lib.ProvingKey."""
        return cls(pk.prove_batch(advice[None], instance[None], instance_len, seed)[0])

    def inner(self):
        return self.data


class ProverService:
    """SRS + Compliance / Resource-Logic proving keys resident on one GPU."""

    def __init__(self, device=0, srs_arrays=None):
        self.ctx = lib.Context(device)
        self.ctx2 = lib.Context(device)   # second stream: the VP batch overlaps the Compliance batch
        s = srs_arrays
        self.srs = self.ctx.load_srs(s["k"], s["g"], s["g_lagrange"], s["w"], s["u"])
        self.kd_c, self.make_c = circuits_taiga.build(True)
        self.kd_v, self.make_v = circuits_taiga.build(False)
        self.pk_c = self.srs.load_circuit(self.kd_c)
        self.pk_v = self.srs.load_circuit(self.kd_v)

    def synthesize_ptx(self, n_ptx, wseed=0):
        """Witness tables for n_ptx partial transactions: dict of stacked numpy arrays (host)."""
        cw = [self.kd_c.witness_arrays(self.make_c(wseed * 1000 + i)) for i in range(COMPLIANCE_PER_PTX * n_ptx)]
        vw = [self.kd_v.witness_arrays(self.make_v(wseed * 1000 + 500 + i)) for i in range(VP_PER_PTX * n_ptx)]
        return {
            "c_adv": np.stack([w[0] for w in cw]), "c_inst": np.stack([w[1] for w in cw]), "c_len": cw[0][2],
            "v_adv": np.stack([w[0] for w in vw]), "v_inst": np.stack([w[1] for w in vw]), "v_len": vw[0][2],
        }

    def build_ptx_batch(self, wit, seed, c_adv=None, v_adv=None, max_batch=64):
        """ShieldedPartialTransaction::build for a batch: returns (compliance proofs, vp proofs) as lists of bytes.
        c_adv / v_adv may override the advice buffers (e.g. pinned host or device-resident torch tensors)."""
        c_adv = wit["c_adv"] if c_adv is None else c_adv
        v_adv = wit["v_adv"] if v_adv is None else v_adv
        res = {}

        def run_vp():
            try:
                res["vp"] = self._prove_chunks(self.pk_v, v_adv, wit["v_inst"], wit["v_len"], seed, max_batch, 1 << 20, self.ctx2)
            except BaseException as ex:  # re-raised in the caller's thread
                res["err"] = ex
        th = threading.Thread(target=run_vp)
        th.start()   # ctypes releases the GIL inside tb_prove_batch, so both batches are enqueued concurrently on two streams
        cp = self._prove_chunks(self.pk_c, c_adv, wit["c_inst"], wit["c_len"], seed, max_batch, 0, self.ctx)
        th.join()
        if "err" in res:
            raise res["err"]
        return cp, res["vp"]

    @property
    def launch_count(self):
        return self.ctx.launch_count + self.ctx2.launch_count

    @staticmethod
    def _prove_chunks(pk, adv, inst, lens, seed, max_batch, index0, ctx):
        kd = pk.keydata
        per = kd.cs.num_advice * kd.n * 32
        if hasattr(adv, "data_ptr"):  # torch tensor (pinned host or device)
            total = adv.numel() // per
        else:
            total = adv.size // per
        out = []
        for s in range(0, total, max_batch):
            e = min(total, s + max_batch)
            if hasattr(adv, "data_ptr"):
                chunk = _TensorSlice(adv, s * per, (e - s) * per)
            else:
                chunk = adv.reshape(total, -1)[s:e]
            out += pk.prove_batch_raw(chunk, e - s, inst[s:e], lens, seed, index0 + s, ctx=ctx)
        return out


class _TensorSlice:
    """A byte range of a torch uint8 tensor, passed to the C ABI by address (host-pinned or device memory)."""

    def __init__(self, t, offset, nbytes):
        self.t, self.offset, self.nbytes = t, offset, nbytes

    def data_ptr(self):
        return self.t.data_ptr() + self.offset
