#!/bin/bash
# session r4e: child order by entry + exit (VT_KEY_MID) — A/B on the bench scene (hit buffers diffed) and the collapse probe under it
bash tools/ab_variants.sh r4e mid0 mid1 2>&1 | grep -v "^$" | cut -c1-420
VT_LIB=$PWD/vistrace_b200/variants/lib_mid1.so timeout 240 python tools/collapse_probe.py --out gpurun_out/r4e/collapse_mid1.jsonl 2> gpurun_out/r4e/collapse.err | cut -c1-640
tail -n 2 gpurun_out/r4e/collapse.err
