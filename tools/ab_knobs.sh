#!/bin/bash
# A/B of library builds x knob sets on the bench scene (run under gpurun): tools/ab_knobs.sh <tag> "<knob sets>" name1 name2 ...
# <knob sets> is gpu_explore's --knobs syntax ("VT_REFILL=24;VT_REFILL=28").  The first variant's hit buffers are the reference the others are diffed against.
TAG=$1; KNOBS=$2; shift; shift
O=gpurun_out/$TAG; mkdir -p $O
rm -f /tmp/ab_ref_hits.npz
for v in "$@"; do
    VT_LIB=$PWD/vistrace_b200/variants/lib_$v.so timeout 900 python tools/gpu_explore.py --quads 1582 --knobs "$KNOBS" --ref-hits /tmp/ab_ref_hits.npz > $O/explore_$v.log 2>&1
    echo "== $v"; grep -E '^\{' $O/explore_$v.log | cut -c1-400
done
