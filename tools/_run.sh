#!/bin/bash
# session L: e2e lanes (streams in flight) x tile size
O=gpurun_out/sL; mkdir -p $O
for cfg in "3 524288 65536" "4 524288 65536" "5 524288 65536" "6 524288 65536" "6 262144 32768" "8 262144 65536" "5 393216 131072" "2 524288 65536"; do
  set -- $cfg
  echo "== lanes=$1 tile=$2 first=$3"; VT_WAVE_LANES=$1 VT_WAVE_TILE=$2 VT_WAVE_FIRST=$3 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['all_hit_records_variant']['ms_per_step'])"
done
