mkdir -p gpurun_out/r3r
(time timeout 300 python -m pytest tests/test_gpu_paths.py -x -q -m gpu) > gpurun_out/r3r/pytest_paths.log 2>&1; tail -6 gpurun_out/r3r/pytest_paths.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r3r/bench_n1.json 2> gpurun_out/r3r/bench_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/r3r/bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config5']['value'], d['config5']['ms_per_frame'], d['config5']['e2e'], d['config5']['samples_in_flight'])"; tail -2 gpurun_out/r3r/bench_n1.err
