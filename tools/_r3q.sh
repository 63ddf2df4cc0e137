mkdir -p gpurun_out/r3q
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 20 --warmup 3) > gpurun_out/r3q/bench_n8.json 2> gpurun_out/r3q/bench_n8.err
python -c "
import json; d=json.loads(open('gpurun_out/r3q/bench_n8.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d.get('weak'), d['config5']['value'], d['config5']['ms_per_frame'], d['config5']['e2e'])"; tail -3 gpurun_out/r3q/bench_n8.err
