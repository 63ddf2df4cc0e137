O=gpurun_out/s4f; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpu" 2>&1 | tail -15 > $O/pytest2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench2.json 2> $O/bench2.err
echo rc=$? >> $O/bench2.err
cat $O/pytest2.log $O/bench2.json; tail -5 $O/bench2.err
