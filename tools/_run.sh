#!/bin/bash
# session M: full GPU suite (incl. DXT ingestion test), bench with 4 lanes, first-tile sweep at 4 lanes
O=gpurun_out/sM; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
for cfg in "4 524288 65536" "4 524288 131072" "4 524288 262144" "4 655360 81920" "4 393216 65536"; do
  set -- $cfg
  echo "== lanes=$1 tile=$2 first=$3"; VT_WAVE_LANES=$1 VT_WAVE_TILE=$2 VT_WAVE_FIRST=$3 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['all_hit_records_variant']['ms_per_step'])"
done
