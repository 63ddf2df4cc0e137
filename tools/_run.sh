mkdir -p gpurun_out/m2
nvidia-smi -L > gpurun_out/m2/smi.txt; nproc >> gpurun_out/m2/smi.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpu or accumulate or config2" 2>&1 | tail -5 > gpurun_out/m2/pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/m2/bench2.json 2> gpurun_out/m2/bench2.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > gpurun_out/m2/bench1.json 2> gpurun_out/m2/bench1.err
cat gpurun_out/m2/smi.txt gpurun_out/m2/pytest.log; cat gpurun_out/m2/bench2.json gpurun_out/m2/bench1.json | cut -c1-900; tail -n 4 gpurun_out/m2/bench2.err
