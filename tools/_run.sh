#!/bin/bash
O=gpurun_out/s12; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; python -c "import json; d=json.load(open('$O/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['config']['hierarchy'])"
