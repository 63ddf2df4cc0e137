#!/bin/bash
# session U: K1 CTA size variants; refill / triangle-round thresholds on the shipped kernel
O=gpurun_out/sU; mkdir -p $O
timeout 900 python tools/gpu_explore.py --quads 1582 --ref-hits /tmp/ab_ref_hits.npz --knobs ";VT_TRI_ROUND=6;VT_TRI_ROUND=10;VT_REFILL=20;VT_REFILL=27;VT_REFILL=28,VT_TRI_ROUND=6" > $O/knobs.log 2>&1
echo "== default lib, knobs"; grep -E '^\{"knobs' $O/knobs.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['knobs'], d['primary_Mrays'], d['bounce_Mrays'], d['bounce_anyhit_Mrays'], d['ms'])"
for v in b64 b96 b256; do
  VT_LIB=$PWD/vistrace_b200/variants/lib_$v.so timeout 600 python tools/gpu_explore.py --quads 1582 --ref-hits /tmp/ab_ref_hits.npz > $O/explore_$v.log 2>&1
  echo "== $v"; grep -E '^\{' $O/explore_$v.log | cut -c1-330
  VT_LIB=$PWD/vistrace_b200/variants/lib_$v.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"
done
