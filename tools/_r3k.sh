mkdir -p gpurun_out/r3k
timeout 600 python tools/frames_in_flight_probe.py > gpurun_out/r3k/fif.jsonl 2> gpurun_out/r3k/fif.err; cat gpurun_out/r3k/fif.jsonl; tail -3 gpurun_out/r3k/fif.err
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r3k/pytest_gpu.log 2>&1; tail -5 gpurun_out/r3k/pytest_gpu.log
