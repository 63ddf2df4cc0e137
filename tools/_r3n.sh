mkdir -p gpurun_out/r3n
(time timeout 600 python -m pytest tests/test_gpu_paths.py -x -q -m gpu) > gpurun_out/r3n/pytest_paths.log 2>&1; tail -6 gpurun_out/r3n/pytest_paths.log
timeout 600 python bench.py --steps 20 --warmup 3 --config5 off > gpurun_out/r3n/bench_n1.json 2> gpurun_out/r3n/bench_n1.err
VT_BENCH_FRAMES_IN_FLIGHT=1 timeout 600 python bench.py --steps 20 --warmup 3 --config5 off --no-cpu > gpurun_out/r3n/bench_n1_fif1.json 2> gpurun_out/r3n/bench_n1_fif1.err
for f in bench_n1 bench_n1_fif1; do python -c "
import json; d=json.loads(open('gpurun_out/r3n/$f.json').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d.get('roofline',{}).get('kernel_ms'), d.get('parity',{}).get('ok'))"; tail -2 gpurun_out/r3n/$f.err; done
