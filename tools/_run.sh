#!/bin/bash
# session N (8 GPUs): bench at N=8 exactly as the driver launches it
O=gpurun_out/sN; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench rc=$?"; cat $O/bench_n8.json; tail -3 $O/bench_n8.err
