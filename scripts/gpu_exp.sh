#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -60 > gpurun_out/test.log; tail -2 gpurun_out/test.log
for w in cfg3-shard cfg4; do timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload $w 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['value'], d['e2e']['value'], d['kernels_ms_per_iteration'])"; done
