#!/bin/bash
# session D: K3 block-aggregated queue atomics; builder leaf-size / traversal-cost sweep
O=gpurun_out/sD; mkdir -p $O
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > $O/bench_queue.json 2> $O/bench_queue.err; echo "rc=$?"
cat $O/bench_queue.json
timeout 300 python -m pytest tests -m gpu -x -q -k "queue or bounce or accumulate" 2>&1 | tail -3
timeout 900 python tools/gpu_explore.py --quads 1582 --rebuild --knobs "VT_MAX_LEAF=4;VT_MAX_LEAF=1;VT_MAX_LEAF=2;VT_MAX_LEAF=3;VT_MAX_LEAF=4,VT_TRAV_COST=0.5;VT_MAX_LEAF=4,VT_TRAV_COST=2;VT_MAX_LEAF=2,VT_TRAV_COST=0.5;VT_MAX_LEAF=6,VT_TRAV_COST=1.5" > $O/leaf.log 2>&1
cat $O/leaf.log
