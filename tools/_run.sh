mkdir -p gpurun_out/c13
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/c13/pytest.log
timeout 600 python tools/gpu_explore.py --quads 1582 --knobs "VT_LAYOUT=quad" > gpurun_out/c13/explore.log 2>&1
for v in pfL2 pfL1 pfL2leaf pfL2far; do
VT_LIB=$PWD/build/variants/lib_$v.so timeout 600 python tools/gpu_explore.py --quads 1582 --knobs "VT_LAYOUT=quad" > gpurun_out/c13/explore_$v.log 2>&1
done
cat gpurun_out/c13/pytest.log; grep -H knobs gpurun_out/c13/explore*.log | cut -c1-440
