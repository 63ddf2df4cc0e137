mkdir -p gpurun_out/c12
free -g | head -2 > gpurun_out/c12/mem.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/c12/pytest.log
timeout 1500 python tools/config_sweep.py --configs 1,2,4,3 --out gpurun_out/c12/sweep.jsonl > gpurun_out/c12/sweep.log 2>&1
timeout 1500 python tools/config_sweep.py --configs 5 --out gpurun_out/c12/sweep.jsonl > gpurun_out/c12/sweep5.log 2>&1
cat gpurun_out/c12/mem.txt gpurun_out/c12/pytest.log; tail -n 5 gpurun_out/c12/sweep.log | cut -c1-1500; tail -n 3 gpurun_out/c12/sweep5.log | cut -c1-2500
