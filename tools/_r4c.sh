#!/bin/bash
# session r4c: config 4 (alpha-tested foliage) and config 2 (props) under the collapse / reinsertion builder options
OUT=gpurun_out/r4c; mkdir -p $OUT
run() { # name, configs, env...
  local name=$1 cfgs=$2; shift 2
  env "$@" timeout 200 python tools/config_sweep.py --configs $cfgs --no-agreement --cpu-rays 20000 --out $OUT/$name.jsonl > $OUT/$name.log 2>&1
  python - "$OUT/$name.jsonl" "$name" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    print(sys.argv[2], d["config"], {k: v["ms"] for k, v in d["stages"].items()}, d.get("primary_node_visits_tri_tests_per_ray"), "build_s", d["build_s"])
PY
}
run default 4 VT_NOP=1
run greedy 4 VT_COLLAPSE=greedy
run reinsert2 4,2 VT_REINSERT=2 VT_REINSERT_FRACTION=0.5
run greedy_reinsert2 4 VT_COLLAPSE=greedy VT_REINSERT=2 VT_REINSERT_FRACTION=0.5
run reinsert8 2 VT_REINSERT=8 VT_REINSERT_FRACTION=0.5
