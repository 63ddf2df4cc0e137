timeout 600 python tools/diag_diff.py quad > gpurun_out/diag_diff.log 2>&1; tail -40 gpurun_out/diag_diff.log
