#!/bin/bash
# session I (2 GPUs): NCCL test + bench at N=2 as the driver launches it
O=gpurun_out/sI; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpu" > $O/pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench rc=$?"; cat $O/bench_n2.json; tail -3 $O/bench_n2.err
