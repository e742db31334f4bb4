#!/bin/bash
# Runs on the GPU box: one ncu capture (DRAM bytes, duration, issue activity, registers: a full --set with sources is 14 MB per kernel) of each kernel of the cfg2 and cfg4 workloads (their DRAM traffic per launch goes to
# profiles/ncu_traffic.json through scripts/update_traffic.py). Usage: scripts/profile_other.sh <tag>
TAG=${1:-r02o}
OUT=gpurun_out
mkdir -p $OUT
cap() { # workload name regex skip
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none -k regex:$3 --launch-skip $4 -c 1 -f -o $OUT/${TAG}_$1_$2 python bench.py --workload $1 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $OUT/${TAG}_$1_$2.log 2>&1
}
cap cfg4 k1 sweepKernel 12
cap cfg4 gram gramTilesKernel 12
cap cfg4 chol choleskyScheduledKernel 12
cap cfg2 k1 sweepKernel 12
cap cfg2 k2 gramCholeskyKernel 12
ls -la $OUT | grep ${TAG}_cfg
