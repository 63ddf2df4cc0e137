mkdir -p gpurun_out/r3w
(time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 3 --config5 off) > gpurun_out/r3w/bench_n2.json 2> gpurun_out/r3w/bench_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/r3w/bench_n2.json').read().strip().splitlines()[-1]); print('bench_n2', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"; tail -1 gpurun_out/r3w/bench_n2.err
