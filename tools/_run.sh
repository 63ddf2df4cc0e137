#!/bin/bash
O=gpurun_out/parity; mkdir -p $O; rm -f $O/parity.jsonl
timeout 200 python tools/parity_report.py --configs 1,2,4,3,5 --out $O/parity.jsonl > $O/parity.log 2>&1; echo "rc=$?"; tail -3 $O/parity.log | cut -c1-400
