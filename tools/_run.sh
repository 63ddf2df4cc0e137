mkdir -p gpurun_out/c10
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/c10/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/c10/bench.json 2> gpurun_out/c10/bench.err
for t in 131072 524288 1048576; do VT_WAVE_TILE=$t timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tile',$t, d['e2e'])" ; done > gpurun_out/c10/tiles.log 2>&1
cat gpurun_out/c10/pytest.log; cat gpurun_out/c10/bench.json; tail -n 3 gpurun_out/c10/bench.err; cat gpurun_out/c10/tiles.log
