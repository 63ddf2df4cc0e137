#!/bin/bash
O=gpurun_out/sZ; mkdir -p $O
for lay in quad exact; do
  echo "== builder=ploc layout=$lay"; VT_LAYOUT=$lay timeout 600 python bench.py --builder ploc --steps 10 --warmup 3 --no-cpu 2> $O/err_$lay.log | tee $O/bench_ploc_$lay.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['node_visits_per_ray'], d['roofline']['tri_tests_per_ray'], d['roofline']['kernel_ms'], d['config']['hierarchy'])"
  tail -1 $O/err_$lay.log
done
