#!/bin/bash
# A/B of library builds on the bench scene (run under gpurun): tools/ab_variants.sh <tag> name1 name2 ...
TAG=$1; shift
O=gpurun_out/$TAG; mkdir -p $O
[ -x tools/ubench/pipes.bin ] && tools/ubench/pipes.bin > $O/ubench_pipes.txt 2>&1
for v in "$@"; do
    VT_LIB=$PWD/vistrace_b200/variants/lib_$v.so timeout 600 python tools/gpu_explore.py --quads 1582 --ref-hits /tmp/ab_ref_hits.npz > $O/explore_$v.log 2>&1
    echo "== $v"; grep -E '^\{' $O/explore_$v.log
done
cat $O/ubench_pipes.txt
# e2e tile-size sweep on the default library (bench.py reads VT_WAVE_TILE through the library)
if [ -n "$AB_TILES" ]; then
  for t in $AB_TILES; do
    echo "== VT_WAVE_TILE=$t"; VT_WAVE_TILE=$t timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'])"
  done
fi
