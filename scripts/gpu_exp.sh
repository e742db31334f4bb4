#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -60 > gpurun_out/test.log; tail -2 gpurun_out/test.log
MB2_CHOL_PROFILE=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench.log 2>&1; grep "chol-profile" gpurun_out/bench.log | tail -1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('head', d['value'], d['e2e']['value'], d['kernels_ms_per_iteration'])"
