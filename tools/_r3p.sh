mkdir -p gpurun_out/r3p
(time timeout 240 python -m pytest tests/test_gpu_group.py -x -q -m gpu) > gpurun_out/r3p/pytest_group.log 2>&1; tail -25 gpurun_out/r3p/pytest_group.log | cut -c1-400
