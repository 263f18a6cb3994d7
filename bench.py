#!/usr/bin/env python
"""bench.py - partial-transaction proofs/sec of the B200-native Taiga prover hot path.

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]`; for N > 1 it is launched under
torch.distributed.run, one rank per GPU.  Rank 0 prints ONE JSON line.

  step      one pass of the hot path over one batch: `--ptx P` shielded partial transactions per GPU
            (default 1 = BASELINE.json configs[1]: 2 Compliance("Action")-shaped + 4 Resource-Logic("VP")-shaped proofs,
            k = 15, Taiga's own SRS), i.e. ShieldedPartialTransaction::build (shielded_ptx.rs:98-134).
  value     whole-job ptx/s with the advice tables already resident in HBM when the timed region starts
            (CUDA events on the library's stream, max over ranks).
  e2e       the same metric through the C ABI with HOST (pinned) advice buffers: host->device copies of the advice
            tables and the device->host read of the proof bytes are inside the timed region.
  roofline  dominant kernel group of a step (by CUDA-event time), algorithmic bytes / its average duration vs the
            measured HBM peak (MEASURED_PEAKS.json); plus the MSM / NTT sweeps of BASELINE configs[3].
  cpu_baseline   the CPU oracle (threaded C++ restatement of the halo2 prover; the Rust reference cannot be built in
            this image) timed on this box's host cores on a bounded sample (1 Compliance + 1 VP proof -> ptx/s).
  --impl reference   times that CPU arm alone, same metric / config (see DESIGN.md "Reference arm").
Synthetic data: Taiga-shaped circuits with satisfying witnesses (taiga_b200/circuits_taiga.py); every proof of the last
timed step is checked with the oracle's verifier restatement outside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N15 = 1 << 15


def load_srs():
    raw = np.fromfile(os.path.join(ROOT, "tests", "golden", "srs_k15_affine.bin"), dtype=np.uint8).reshape(-1, 64)
    return {"k": 15, "n": N15, "g": raw[:N15], "g_lagrange": raw[N15:2 * N15], "w": raw[2 * N15], "u": raw[2 * N15 + 1]}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        try:
            p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                 stdout=subprocess.PIPE, text=True)
        except Exception:
            return
        self.proc = p
        for line in p.stdout:
            if self.stop_flag:
                break
            self.samples.append([x.strip() for x in line.split(",")])
        p.terminate()

    def summary(self):
        self.stop_flag = True
        if getattr(self, "proc", None):
            self.proc.terminate()
        sm = sorted(int(float(s[0])) for s in self.samples if s and s[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for i, nme in enumerate(names):
                if len(s) > 3 + i and s[3 + i].lower().startswith("active"):
                    reasons.add(nme)
        mx = [int(float(s[1])) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_prove_sample(srs, threads=None):
    """Oracle (port) on the host cores: one Compliance-shaped + one VP-shaped proof; returns seconds and ptx/s."""
    from oracle import cpu as oc
    from taiga_b200 import circuits_taiga as ct
    if threads:
        oc.set_threads(threads)
    cores = oc.set_threads(0)
    out = {}
    for comp in (True, False):
        kd, make = ct.build(comp)
        key = oc.OracleKey(kd, srs)
        adv, inst, lens = kd.witness_arrays(make(3))
        t = time.time()
        proof = key.prove(adv, inst, lens, bytes(range(32)))
        out["compliance" if comp else "vp"] = time.time() - t
        assert key.verify(inst, lens, proof) == 0
    sec_per_ptx = 2 * out["compliance"] + 4 * out["vp"]
    return {"value": 1.0 / sec_per_ptx, "unit": "ptx/s", "cores": cores, "kind": "port",
            "sample": "1 Compliance-shaped + 1 VP-shaped proof (k=15), serial 2C+4V extrapolation as in shielded_ptx.rs:107-125",
            "compliance_proof_s": round(out["compliance"], 3), "vp_proof_s": round(out["vp"], 3),
            "reference_published": {"compliance_proof_s": 3.1445, "vp_proof_s": 2.2328, "source": "taiga_halo2/benches/Perfromance.md:3,9 (hardware not stated)"}}


def run_reference(args, rank, world):
    """--impl reference: the CPU arm (oracle port; the Rust reference cannot be compiled here: no cargo, un-vendored git deps)."""
    if rank != 0:
        return
    srs = load_srs()
    times = []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_prove_sample(srs)
        if i >= args.warmup:
            times.append(1.0 / base["value"])
    ms = 1e3 * sum(times) / len(times)
    val = 1e3 / ms
    line = {"impl": "reference", "metric": "partial-tx proofs/sec", "value": val, "unit": "ptx/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32x8 (255-bit Montgomery integers)", "data": "synthetic",
            "config": {"workload": "1 shielded partial transaction = 2 Compliance-shaped + 4 VP-shaped Halo2/IPA proofs, k=15 (BASELINE configs[1]); CPU step = bounded sample 1C+1V extrapolated 2C+4V"},
            "cpu_baseline": dict(base, value=val), "e2e": {"value": val, "unit": "ptx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def sweep(ctx, hbm_peak, quick):
    """BASELINE configs[3]: standalone Vesta/Pallas MSM 2^16-2^22 and Fp NTT 2^17-2^23, device resident, algorithmic GB/s."""
    import torch
    from taiga_b200 import lib
    st = torch.cuda.ExternalStream(ctx.stream)
    out = {"msm": [], "ntt": []}
    rng = np.random.default_rng(0)
    srs = load_srs()
    msm_sizes = [16, 18, 20, 22] if quick else list(range(16, 23))
    ntt_sizes = [17, 19, 21, 23] if quick else list(range(17, 24))
    for lg in msm_sizes:
        n = 1 << lg
        pts = np.concatenate([srs["g"], srs["g_lagrange"]] * max(1, n // (2 * N15)))[:n]
        sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        sc[:, 31] &= 0x3F
        d_sc = torch.from_numpy(sc).cuda()
        d_pts = torch.from_numpy(np.ascontiguousarray(pts)).cuda()
        d_out = torch.zeros(64, dtype=torch.uint8, device="cuda")
        ctx.dev_to_mont(lib.TB_FP, d_sc, n)
        ctx.dev_to_mont(lib.TB_FQ, d_pts, 2 * n)
        for _ in range(2):
            ctx.dev_msm(lib.TB_VESTA, n, d_sc, d_pts, d_out)
        ctx.sync()
        reps = 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            ctx.dev_msm(lib.TB_VESTA, n, d_sc, d_pts, d_out)
        e1.record(st)
        ctx.sync()
        ms = e0.elapsed_time(e1) / reps
        gbs = 96.0 * n / (ms * 1e-3) / 1e9
        out["msm"].append({"log2_n": lg, "ms": round(ms, 3), "gpoints_per_s": round(n / ms / 1e6, 4), "alg_gbs": round(gbs, 2), "frac_hbm": round(gbs / hbm_peak, 5)})
        del d_sc, d_pts
    for lg in ntt_sizes:
        n = 1 << lg
        x = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        x[:, 31] &= 0x3F
        d_in = torch.from_numpy(x).cuda()
        d_out = torch.empty_like(d_in)
        d_scr = torch.empty_like(d_in)
        ctx.dev_to_mont(lib.TB_FP, d_in, n)
        for _ in range(2):
            ctx.dev_ntt(lib.TB_FP, lg, d_in, d_out, d_scr)
        ctx.sync()
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            ctx.dev_ntt(lib.TB_FP, lg, d_in, d_out, d_scr)
        e1.record(st)
        ctx.sync()
        ms = e0.elapsed_time(e1) / reps
        gbs = 64.0 * n / (ms * 1e-3) / 1e9
        out["ntt"].append({"log2_n": lg, "ms": round(ms, 4), "alg_gbs": round(gbs, 1), "frac_hbm": round(gbs / hbm_peak, 4)})
        del d_in, d_out, d_scr
    return out


# DRAM traffic per launch of the dominant kernels from the committed ncu captures (profiles/r01_ncu_summary.md, capture B:
# Compliance-shaped circuit, 2 proofs per launch = 20 advice MSMs resp. one sub-coset of 2 proofs)
# The contract is ONE JSON line on stdout.  Libraries may write to file descriptor 1 behind Python's back (NCCL prints its
# version banner there when the communicator is created), so the real stdout is set aside at import time, everything
# else that targets fd 1 is sent to stderr, and only emit() writes to the real one.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


NCU_TRAFFIC = {   # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, from profiles/r01_ncu_summary.md capture C (ncu --set full)
    "msm_accum": {"dram_bytes_per_launch": 47.56e6, "algorithmic_bytes_same_launch": 2 * 96 * N15,
                  "capture": "capture C, K = 2 dense commitments: the 42 MB fixed-base window table is streamed once per launch (amortised over K; K = 20 read 48.6 MB vs 63 MB algorithmic in capture B)"},
    "quotient_gates": {"dram_bytes_per_launch": 28.74e6, "algorithmic_bytes_same_launch": 32 * 28 * N15,
                       "capture": "capture C, one sub-coset of one proof, 8 constraint parts: every column-coset is read from DRAM once, the other parts hit L2"},
    "ntt": {"dram_bytes_per_launch": 15.92e6, "algorithmic_bytes_same_launch": 15 * 32 * N15,
            "capture": "capture C, one pass over 15 columns: DRAM read = algorithmic read, the writes stay in the 126 MB L2"},
}

ALG_BYTES_NOTE = {
    "ntt": "64*n per size-n transform (read + write once)",
    "msm_accum": "96 B per MSM term (64 B affine base + 32 B scalar), SURVEY 8d",
    "msm_sort": "96 B per MSM term", "msm_reduce": "96 B per MSM term",
    "quotient_gates": "32*(C+1) B per extended row, C = column-cosets read", "quotient_finish": "32*(C+1) B per extended row",
    "ipa_fold": "96 B per folded generator", "transcript": "-", "lookup_sort": "64 B per key", "poly": "64 B per coefficient",
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--ptx", type=int, default=1, help="partial transactions per GPU per step (1 = BASELINE configs[1], 64 = configs[2])")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--full-sweep", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--serial", action="store_true", help="one stream, no threads (for ncu launch lists; not a benchmark configuration)")
    ap.add_argument("--batch-probe", type=int, default=8, help="also time a batch of this many ptx per step (0 = off)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from taiga_b200 import ptx, shard
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    srs = load_srs()
    nw = 2 if args.ptx <= 2 else 1   # small batches: two streams per circuit so latency-bound phases overlap
    if args.serial:
        nw = 1
    svc = ptx.ProverService(local, srs, c_workers=int(os.environ.get("TB_C_WORKERS", nw)), v_workers=int(os.environ.get("TB_V_WORKERS", nw)), serial=args.serial)
    ctx = svc.ctx
    P = args.ptx
    wit = svc.synthesize_ptx(P, wseed=rank)
    h2d = wit["c_adv"].nbytes + wit["v_adv"].nbytes + wit["c_inst"].nbytes + wit["v_inst"].nbytes
    d2h = svc.pk_c.proof_len * 2 * P + svc.pk_v.proof_len * 4 * P
    c_pin, v_pin = torch.from_numpy(wit["c_adv"]).pin_memory(), torch.from_numpy(wit["v_adv"]).pin_memory()
    c_dev, v_dev = c_pin.cuda(), v_pin.cuda()
    st = torch.cuda.ExternalStream(ctx.stream)
    seed0 = bytes((rank * 37 + i) & 0xFF for i in range(32))

    def step(i, device_resident):
        seed = bytes((b + i) & 0xFF for b in seed0)
        proofs = svc.build_ptx_batch(wit, seed, c_dev if device_resident else c_pin, v_dev if device_resident else v_pin)
        if world > 1:  # the only collective on the path: gather the finished proof bytes (fixed-size records) over NCCL
            rec = shard.pack_records(proofs[0], proofs[1], svc.pk_c.proof_len, svc.pk_v.proof_len)
            shard.gather_records(rec, P * world, device="cuda")
        return proofs

    def timed(device_resident):
        for i in range(args.warmup):
            step(i, device_resident)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = svc.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(st)
        last = None
        for i in range(args.steps):
            last = step(100 + i, device_resident)
        e1.record(st)
        torch.cuda.synchronize()
        wall = time.time() - t0
        if world > 1:
            dist.barrier()
        ms = max(e0.elapsed_time(e1), 0.0)
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), svc.launch_count - l0, last

    sampler = ClockSampler(local)
    sampler.start()
    dev_ms, dev_wall_ms, launches, last = timed(True)
    e2e_ms, e2e_wall_ms, _, last_e2e = timed(False)
    clocks = sampler.summary()

    # one profiled step (CUDA events around every kernel group) for the share-of-step table and the roofline.  It runs the
    # workers one after the other: with the streams overlapped an event pair also times the wait for SMs held by the other
    # streams' kernels, and the shares would not be comparable with the (serialised) ncu launch list in profiles/.
    svc.prof_enable(True)
    was_serial, svc.serial = svc.serial, True
    step(999, True)
    svc.serial = was_serial
    prof = svc.prof_read()
    svc.prof_enable(False)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # acceptance (outside the timed region): every proof of the last e2e step under the oracle's verifier restatement
    accepted = None
    try:
        from oracle import cpu as oc
        kc, kv = oc.OracleKey(svc.kd_c, srs), oc.OracleKey(svc.kd_v, srs)
        accepted = all(kc.verify(wit["c_inst"][i], wit["c_len"], p) == 0 for i, p in enumerate(last_e2e[0])) and \
            all(kv.verify(wit["v_inst"][i], wit["v_len"], p) == 0 for i, p in enumerate(last_e2e[1]))
    except Exception as ex:  # pragma: no cover
        accepted = "verifier unavailable: %r" % (ex,)
    # ... and under the library's own batched device verifier (tb_verify_batch)
    try:
        accepted_dev = all(svc.pk_c.verify_batch(wit["c_inst"], wit["c_len"], list(last_e2e[0]))) and \
            all(svc.pk_v.verify_batch(wit["v_inst"], wit["v_len"], list(last_e2e[1]), ctx=svc.v_workers[0][0]))
    except Exception as ex:  # pragma: no cover
        accepted_dev = "device verifier failed: %r" % (ex,)

    hbm_peak, peak_kind = measured_peaks()
    total_ptx = P * world
    dev_step_ms, e2e_step_ms = dev_wall_ms / args.steps, e2e_wall_ms / args.steps
    value = total_ptx / (dev_step_ms * 1e-3)
    e2e_val = total_ptx / (e2e_step_ms * 1e-3)
    # dominant kernel group of the profiled step
    tot_prof = sum(v[0] for v in prof.values()) or 1.0
    # the roofline is quoted for the dominant SINGLE kernel (as in the ncu launch list, profiles/r01_launches_bench_serial_v2_summary.md);
    # the other groups bundle several short launches and the gaps between them, and are listed under per_kernel
    single = {"msm_accum": "msm_accum_kernel", "quotient_gates": "q_interp_kernel", "ntt": "ntt_pass_kernel"}
    cands = {k_: v_ for k_, v_ in prof.items() if k_ in single} or prof
    top = max(cands.items(), key=lambda kv_: kv_[1][0])
    n = N15
    nproofs_c, nproofs_v = 2 * P, 4 * P
    # algorithmic bytes of one profiled step per category (SURVEY 8d figures x units processed)
    msm_terms = nproofs_c * (33 + 30) * n + nproofs_v * (26 + 30) * n    # commitments + IPA rounds (2n terms, upper bound n each side)
    alg = {
        "msm_accum": 96.0 * msm_terms, "msm_sort": 96.0 * msm_terms, "msm_reduce": 96.0 * msm_terms,
        "ntt": 64.0 * n * (prof["ntt"][1] and (nproofs_c * (14 + 15 * 16 + 16) + nproofs_v * (15 + 16 * 8 + 8))),
        "quotient_gates": 32.0 * (svc.kd_c.cs.num_advice + svc.kd_c.cs.num_fixed + 2) * (1 << 19) * nproofs_c + 32.0 * (10 + svc.kd_v.cs.num_fixed + 2) * (1 << 18) * nproofs_v,
        "ipa_fold": 96.0 * n * (nproofs_c + nproofs_v),
    }
    top_name, (top_ms, top_groups) = top
    top_bytes = alg.get(top_name)
    roof = {"bound": "hbm", "kernel": top_name, "kernel_name": single.get(top_name), "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None, "traffic": None,
            "peak_source": peak_kind, "share_of_step": round(top_ms / tot_prof, 3), "launch_groups": top_groups, "avg_group_ms": round(top_ms / max(1, top_groups), 4),
            "algorithmic_bytes_per_step": top_bytes, "bytes_rule": ALG_BYTES_NOTE.get(top_name),
            "note": "255-bit modular arithmetic: the path is INT32-pipe bound, not HBM bound (SURVEY 8d); frac is reported against HBM as the metric asks"}
    if top_bytes:
        roof["achieved"] = round(top_bytes / (top_ms * 1e-3) / 1e9, 2)
        roof["frac"] = round(roof["achieved"] / hbm_peak, 5)
    if top_name in NCU_TRAFFIC:
        roof["traffic"] = NCU_TRAFFIC[top_name]["dram_bytes_per_launch"]
        roof["traffic_detail"] = NCU_TRAFFIC[top_name]
    # the same figures for every kernel group of the step (the dominant one is repeated above)
    roof["per_kernel"] = {k: {"ms": round(v[0], 3), "groups": v[1], "achieved_gbs": (round(alg[k] / (v[0] * 1e-3) / 1e9, 2) if alg.get(k) and v[0] > 0 else None),
                              "frac": (round(alg[k] / (v[0] * 1e-3) / 1e9 / hbm_peak, 5) if alg.get(k) and v[0] > 0 else None)} for k, v in prof.items()}
    line = {
        "metric": "partial-tx proofs/sec", "value": round(value, 4), "unit": "ptx/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dev_step_ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32x8 (255-bit Montgomery integers, Pasta Fp/Fq)", "data": "synthetic",
        "config": {"workload": "%d shielded partial transaction(s) per GPU per step = %d Compliance-shaped (degree 17, ext 2^19, 4480 B proofs) + %d VP-shaped (degree 9) Halo2/IPA proofs, k=15, Taiga params_15 SRS (BASELINE configs[%d])"
                   % (P, 2 * P, 4 * P, 1 if P == 1 else 2), "ptx_per_gpu": P, "parallelism": "independent ptx per GPU (no collective inside a proof; NCCL all_gather of proof bytes); %d CUDA streams per circuit" % nw,
                   "l2": "inputs (60 MiB advice per ptx + 0.9 GB resident key cosets) exceed L2; no explicit flush", "proofs_accepted_by_oracle_verifier": accepted, "proofs_accepted_by_device_verifier": accepted_dev},
        "e2e": {"value": round(e2e_val, 4), "unit": "ptx/s", "ms_per_step": round(e2e_step_ms, 3), "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "device_event_ms_per_step": round(dev_ms / args.steps, 3),
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": roof,
        "profile_share": {k: round(v[0] / tot_prof, 4) for k, v in sorted(prof.items(), key=lambda kv_: -kv_[1][0])},
        "profile_ms": {k: round(v[0], 3) for k, v in prof.items()},
        "kernel_time_over_step_time": round(tot_prof / dev_step_ms, 3),
    }
    if args.batch_probe and world == 1 and P == 1:
        # BASELINE configs[2]-style throughput probe: the same six witnesses tiled to `batch_probe` ptx (distinct blinding seeds per
        # proof, so distinct proofs), device resident, one stream per circuit
        bp = args.batch_probe
        wit_b = {k_: (np.concatenate([v_] * bp) if k_.endswith(("_adv", "_inst")) else v_) for k_, v_ in wit.items()}
        cb, vb = torch.from_numpy(wit_b["c_adv"]).cuda(), torch.from_numpy(wit_b["v_adv"]).cuda()
        svc.build_ptx_batch(wit_b, seed0, cb, vb)
        torch.cuda.synchronize()
        t0 = time.time()
        reps = 2
        for i in range(reps):
            svc.build_ptx_batch(wit_b, bytes((b + 7 + i) & 0xFF for b in seed0), cb, vb)
        torch.cuda.synchronize()
        dt = (time.time() - t0) / reps
        line["batch_probe"] = {"ptx_per_step": bp, "value": round(bp / dt, 3), "unit": "ptx/s", "ms_per_step": round(dt * 1e3, 2),
                               "note": "device-resident, witnesses of the P=1 step tiled %dx (distinct seeds)" % bp}
        del cb, vb
    if not args.no_sweep and world == 1:
        line["sweeps"] = sweep(ctx, hbm_peak, quick=not args.full_sweep)
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_prove_sample(srs)
        line["speedup_e2e_vs_cpu_port"] = round(e2e_val / line["cpu_baseline"]["value"], 2)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
