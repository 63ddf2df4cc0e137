#!/bin/bash
O=gpurun_out/sT; mkdir -p $O
for cfg in "4 524288 0" "4 524288 5" "4 524288 3" "6 524288 3" "4 524288 6" "6 393216 4"; do
  set -- $cfg
  echo "== lanes=$1 tile=$2 ctas_per_sm=$3"; VT_WAVE_LANES=$1 VT_WAVE_TILE=$2 VT_WAVE_CTAS_PER_SM=$3 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['all_hit_records_variant']['ms_per_step'])"
done
