mkdir -p gpurun_out/c9
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/c9/pytest.log
timeout 600 python tools/gpu_explore.py --quads 1582 --knobs "VT_LAYOUT=quad;VT_LAYOUT=compact;VT_LAYOUT=quad,VT_TRI_ROUND=4;VT_LAYOUT=quad,VT_TRI_ROUND=12;VT_LAYOUT=quad,VT_TRI_ROUND=16" > gpurun_out/c9/explore.log 2>&1
VT_LIB=$PWD/build/variants/lib_norun.so timeout 600 python tools/gpu_explore.py --quads 1582 --knobs "VT_LAYOUT=quad" > gpurun_out/c9/explore_norun.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/c9/bench.json 2> gpurun_out/c9/bench.err
cat gpurun_out/c9/pytest.log; grep -h knobs gpurun_out/c9/explore*.log | cut -c1-420; cat gpurun_out/c9/bench.json; tail -n 3 gpurun_out/c9/bench.err
