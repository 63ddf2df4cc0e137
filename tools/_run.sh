mkdir -p gpurun_out/c7
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -25 > gpurun_out/c7/pytest.log
timeout 600 python tools/gpu_explore.py --quads 1582 --knobs "VT_LAYOUT=exact;VT_LAYOUT=quad;VT_LAYOUT=quad,VT_TRI_ROUND=4;VT_LAYOUT=quad,VT_TRI_ROUND=12;VT_LAYOUT=quad,VT_REFILL=16;VT_LAYOUT=quad,VT_REFILL=28" > gpurun_out/c7/explore.log 2>&1
cat gpurun_out/c7/pytest.log | tail -12; grep knobs gpurun_out/c7/explore.log | cut -c1-420; tail -3 gpurun_out/c7/explore.log | cut -c1-300
