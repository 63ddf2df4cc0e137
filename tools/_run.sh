#!/bin/bash
# session P: range-refit + fallback tests; split-step experiment
O=gpurun_out/sP; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "refit" > $O/pytest_refit.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_refit.log
timeout 600 python tools/split_step.py --parts 1,2,3,4,6 > $O/split.log 2>&1; cat $O/split.log | tail -8
