#!/bin/bash
# Runs on the GPU box (gpurun): launch list + one full ncu capture of each hot kernel of the default bench workload.
# Usage: scripts/profile_gpu.sh <tag>     -> gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_{k1,k2}.ncu-rep
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_launches.log 2>&1
for spec in "k1:sweepKernel" "k2:gramCholeskyKernel"; do
  name=${spec%%:*}; regex=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:$regex --launch-skip 12 -c 1 -f -o $OUT/${TAG}_${name} $BENCH > $OUT/${TAG}_${name}.log 2>&1
done
# (gpurun merges at most 64 MiB back: PROFILE_MINIMAL=1 stops after the two kernels of the default path)
if [ -n "${PROFILE_MINIMAL:-}" ]; then ls -la $OUT | tail -6; exit 0; fi
# the three-kernel path's Gram / Cholesky kernels and the persistent kernel, for comparison
ncu --set full --clock-control none --import-source on -k regex:gramTilesKernel --launch-skip 4 -c 1 -f -o $OUT/${TAG}_gram $BENCH --fused-mode 1 > $OUT/${TAG}_gram.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:choleskyScheduledKernel --launch-skip 4 -c 1 -f -o $OUT/${TAG}_chol $BENCH --fused-mode 1 > $OUT/${TAG}_chol.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fusedSolveKernel -c 1 -f -o $OUT/${TAG}_persistent $BENCH --fused-mode 2 > $OUT/${TAG}_persistent.log 2>&1
ls -la $OUT | tail -12
