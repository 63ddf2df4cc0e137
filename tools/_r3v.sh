mkdir -p gpurun_out/r3v
(time timeout 240 python -m pytest tests/test_gpu_group.py tests/test_gpu_paths.py -x -q -m gpu) > gpurun_out/r3v/pytest.log 2>&1; tail -5 gpurun_out/r3v/pytest.log | cut -c1-300
(time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 3 --config5 off) > gpurun_out/r3v/bench_n2.json 2> gpurun_out/r3v/bench_n2.err
timeout 300 python bench.py --steps 20 --warmup 3 --config5 off --no-cpu > gpurun_out/r3v/bench_n1.json 2> gpurun_out/r3v/bench_n1.err
for f in bench_n2 bench_n1; do python -c "
import json; d=json.loads(open('gpurun_out/r3v/$f.json').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"; tail -1 gpurun_out/r3v/$f.err; done
