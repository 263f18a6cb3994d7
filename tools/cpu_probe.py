"""What the GPU box's host side looks like (cores, quota, load, memory) and how the CPU prover farm of bench.py scales on it."""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def sh(c):
    try:
        return subprocess.run(c, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as ex:
        return repr(ex)


info = {"affinity": len(os.sched_getaffinity(0)), "cpu_count": os.cpu_count(), "host_threads": bench.CpuFarm.host_threads(),
        "cpu.max": sh("cat /sys/fs/cgroup/cpu.max"), "cfs_quota": sh("cat /sys/fs/cgroup/cpu/cpu.cfs_quota_us"), "loadavg": sh("cat /proc/loadavg"),
        "lscpu": sh("lscpu | grep -E 'Model name|Socket|Core|Thread|^CPU\\(s\\)|MHz|L3'"), "mem": sh("free -g | head -2"), "cpu_stat": sh("cat /sys/fs/cgroup/cpu.stat"), "shm": sh("df -h /dev/shm | tail -1"), "top": sh("top -b -n 1 | head -15")}
print(json.dumps(info, indent=1), flush=True)
W, T = int(sys.argv[1]), int(sys.argv[2])
t = time.time()
farm = bench.CpuFarm(workers=W, threads=T)
print("farm of %d x %d ready in %.1f s" % (farm.workers, T, time.time() - t), flush=True)
for n in [int(a) for a in sys.argv[3:]]:
    r = farm.sample(n)
    print(n, "provers:", round(r["value"], 4), "ptx/s; C", r["compliance_proof_s"], "V", r["vp_proof_s"], "wall", r["sample_wall_s"], "load", sh("cat /proc/loadavg"), flush=True)
farm.close()
print("cpu.stat after:", sh("cat /sys/fs/cgroup/cpu.stat"))
