#!/bin/bash
# Runs on the GPU box: parity tests, then the device-resident bench of the default workload with per-kernel times.
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -4
