#!/bin/bash
# session O: every BASELINE config at full size on the final library
O=gpurun_out/sO; mkdir -p $O
timeout 1700 python tools/config_sweep.py --configs 1,2,3,4,5 --out $O/sweep.jsonl > $O/sweep.log 2>&1; echo "rc=$?"; tail -30 $O/sweep.log
