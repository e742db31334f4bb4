#!/bin/bash
# Runs on the GPU box: the driver's round-end sequence (tests, smoke, both bench arms with defaults).
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time python bench.py --impl reference ) 2>&1 | tail -6
( time python bench.py ) 2>&1 | tail -6
