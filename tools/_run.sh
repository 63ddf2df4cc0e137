#!/bin/bash
# session G: K5 device refit — tests, timing at 6 M triangles, PCIe probe
O=gpurun_out/sG; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "refit" > $O/pytest_refit.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_refit.log
timeout 600 python tools/refit_time.py > $O/refit.log 2>&1; tail -3 $O/refit.log
timeout 300 python tools/pcie_probe.py > $O/pcie.log 2>&1; head -12 $O/pcie.log
