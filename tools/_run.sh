#!/bin/bash
# Scratch driver for one-off gpurun sessions (the sessions of this round are summarised in profiles/).
# Standard session: tools/gpu_session.sh <tag>; A/B of -D builds: tools/build_variants.sh + tools/ab_variants.sh;
# full-size correctness report: tools/parity_report.py; per-config timings: tools/config_sweep.py.
bash tools/gpu_session.sh "${1:-scratch}"
