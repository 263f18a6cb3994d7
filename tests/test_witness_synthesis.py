"""CPU: the host witness synthesis of the Taiga-shaped circuits (the stand-in for the Rust `Circuit::synthesize`,
compliance_circuit.rs:174-327, that feeds the prover): the multi-process, shared-memory path of ProverService.synthesize_ptx
produces exactly the tables of the serial path, and replaying the cached fixed-base window tables does not change a witness."""
import numpy as np

from taiga_b200 import circuits_taiga as ct
from taiga_b200 import ptx


class _Svc:   # what synthesize_ptx touches of a ProverService (no GPU needed)
    pass


def test_parallel_synthesis_equals_serial():
    svc = _Svc()
    svc.kd_c, svc.make_c = ct.build(True)
    svc.kd_v, svc.make_v = ct.build(False)
    par = ptx.ProverService.synthesize_ptx(svc, 2, wseed=4, procs=3)
    ser = ptx.ProverService.synthesize_ptx(svc, 2, wseed=4, procs=1)
    for key in ("c_adv", "v_adv", "c_inst", "v_inst", "c_len", "v_len"):
        assert np.array_equal(par[key], ser[key]), key
    assert par["c_adv"].shape == (4, svc.kd_c.cs.num_advice, 1 << 15, 32) and par["v_adv"].shape == (8, svc.kd_v.cs.num_advice, 1 << 15, 32)
    # a fresh Shape (empty cache of fixed-base window tables) gives the same witness as one that replays its cache
    kd2, make2 = ct.build(True)
    a = kd2.witness_arrays(make2(4 * 100000 + 1))[0]
    assert np.array_equal(a, par["c_adv"][1])
    # distinct proofs get distinct witnesses
    assert not np.array_equal(par["c_adv"][0], par["c_adv"][1])
