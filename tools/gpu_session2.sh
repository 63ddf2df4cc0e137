#!/bin/bash
# Round-2 standard GPU session (run under gpurun): GPU suite, bench (both arms, as the driver runs them), ncu launch list and one
# full capture of the bounce launch of K1 (config 5 side measurement off under ncu).
# usage: tools/gpu_session2.sh <tag>
TAG=${1:-r2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt
(time timeout 900 python -m pytest tests -m gpu -x -q --durations=8) > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --config5 off > $OUT/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_traverse -s 8 -c 1 -o $OUT/k_traverse_bounce \
    python bench.py --steps 2 --warmup 3 --no-cpu --config5 off > $OUT/ncu_full.log 2>&1
tail -4 $OUT/pytest_gpu.log; cat $OUT/bench.json; cat $OUT/bench_ref.json; tail -3 $OUT/bench.err
