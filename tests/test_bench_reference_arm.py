"""bench.py --impl reference (the CPU arm the driver times beside ours): one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_threads_honours_the_affinity_mask():
    sys.path.insert(0, ROOT)
    import bench
    hw = bench.CpuFarm.host_threads()
    assert 1 <= hw <= len(os.sched_getaffinity(0))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, TB_CPU_PROVERS="1", TB_CPU_THREADS_PER_PROVER=str(min(8, len(os.sched_getaffinity(0)))))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "partial-tx proofs/sec" and line["unit"] == "ptx/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0
    assert line["value"] > 0 and abs(line["e2e"]["value"] - line["value"]) < 1e-12
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["provers"] == 1 and cb["cores"] == cb["threads_per_prover"] and "sample" in cb
    assert "64 shielded partial transaction" in line["config"]["workload"]
    # the rate is the one the two timed proofs imply (2 Compliance + 4 VP proofs per partial transaction)
    assert abs(line["value"] - 1.0 / (2 * cb["compliance_proof_s"] + 4 * cb["vp_proof_s"])) / line["value"] < 0.01


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""
