#!/bin/bash
# ncu evidence for profiles/: ONE `--set full` pass over 60 consecutive launches of the hot kernels in steady state (exported to
# CSV on the box: the reports themselves are too large to travel) and the launch list of one bench step.
set -x
mkdir -p gpurun_out
TUNE_EXTRA="{}" timeout 900 ncu --set full --clock-control none -k 'regex:msm_ba_bwd_kernel|msm_ba_fwd_kernel|q_interp_kernel|q_finish_kernel|msm_sort_kernel|ntt_pass_kernel|msm_linesum_kernel|msm_weighted_kernel' \
  -s 400 -c 60 -f -o /tmp/r02_ncu_hot python tools/tune.py 8 > gpurun_out/ncu_hot.log 2>&1
tail -2 gpurun_out/ncu_hot.log
ncu -i /tmp/r02_ncu_hot.ncu-rep --page raw --csv > gpurun_out/r02_ncu_hot_raw.csv 2>/dev/null
ls -la /tmp/r02_ncu_hot.ncu-rep gpurun_out/r02_ncu_hot_raw.csv
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file gpurun_out/r02_launches_bench_p64.csv python bench.py --steps 1 --warmup 0 --no-sweep --no-cpu --no-latency --serial > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log; wc -l gpurun_out/r02_launches_bench_p64.csv
du -sh gpurun_out
