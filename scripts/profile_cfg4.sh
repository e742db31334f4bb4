#!/bin/bash
# Runs on the GPU box: one full ncu capture of each kernel of the cfg4 workload (2048 x bodyhands300, three-kernel path).
TAG=${1:-r02m}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --workload cfg4 --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
for spec in "cfg4_k1:sweepKernel" "cfg4_gram:gramTilesKernel" "cfg4_chol:choleskyScheduledKernel"; do
  name=${spec%%:*}; regex=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex --launch-skip 12 -c 1 -f -o $OUT/${TAG}_${name} $BENCH > $OUT/${TAG}_${name}.log 2>&1
done
ls -la $OUT | grep ${TAG}_cfg4
