"""ORACLE (test infrastructure, never shipped): one CPU prover process of bench.py's cpu_baseline / --impl reference leg.

bench.py starts W of these side by side, each with T threads (W*T = the host's hardware threads), so that the CPU
arm is timed the way a throughput-minded operator would run the reference prover on the box: independent proofs
in parallel, every core busy.  Protocol on stdin/stdout (one line each way):
  -> "ready"                       after keygen, witness generation and one warm-up proof (untimed, as in the reference's benches)
  <- "go"                          prove one Compliance-shaped and one VP-shaped proof, verify both
  -> "<compliance_s> <vp_s>"       wall seconds of the two create_proof calls
  <- "quit"
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N15 = 1 << 15


def main():
    threads = int(sys.argv[1])
    from oracle import cpu as oc
    from taiga_b200 import circuits_taiga as ct
    oc.set_threads(threads)
    raw = np.fromfile(os.path.join(ROOT, "tests", "golden", "srs_k15_affine.bin"), dtype=np.uint8).reshape(-1, 64)
    srs = {"k": 15, "n": N15, "g": raw[:N15], "g_lagrange": raw[N15:2 * N15], "w": raw[2 * N15], "u": raw[2 * N15 + 1]}
    keys = {}
    for comp in (True, False):
        kd, make = ct.build(comp)
        keys[comp] = (oc.OracleKey(kd, srs), kd.witness_arrays(make(3)))
    key, (adv, inst, lens) = keys[False]
    key.prove(adv, inst, lens, bytes(range(32)))   # untimed: first-touch of the heap and the thread pool
    print("ready", flush=True)
    for line in sys.stdin:
        if line.strip() != "go":
            break
        out = []
        for comp in (True, False):
            key, (adv, inst, lens) = keys[comp]
            t = time.time()
            proof = key.prove(adv, inst, lens, bytes(range(32)))
            out.append(time.time() - t)
            assert key.verify(inst, lens, proof) == 0
        print("%.6f %.6f" % tuple(out), flush=True)


if __name__ == "__main__":
    main()
