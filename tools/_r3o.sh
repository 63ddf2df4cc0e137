mkdir -p gpurun_out/r3o
(time timeout 600 python -m pytest tests/test_gpu_group.py -x -q -m gpu) > gpurun_out/r3o/pytest_group.log 2>&1; tail -12 gpurun_out/r3o/pytest_group.log
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 3 --config5 off) > gpurun_out/r3o/bench_n2.json 2> gpurun_out/r3o/bench_n2.err
for f in bench_n2; do python -c "
import json; d=json.loads(open('gpurun_out/r3o/$f.json').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"; tail -2 gpurun_out/r3o/$f.err; done
