#!/bin/bash
# session r4f: final standard session of round 2 (GPU suite, both bench arms, launch list, full capture) + configs 4 and 2 on the final library
bash tools/gpu_session2.sh r4f
VT_TIMING=1 timeout 200 python tools/config_sweep.py --configs 4,2 --cpu-rays 100000 --out gpurun_out/r4f/sweep_4_2.jsonl > gpurun_out/r4f/sweep_4_2.log 2>&1
grep -E "build_quads" gpurun_out/r4f/sweep_4_2.log | sort | uniq -c
python - <<'PY'
import json
for l in open("gpurun_out/r4f/sweep_4_2.jsonl"):
    d = json.loads(l)
    print(d["config"], d["wave"], {k: v["ms"] for k, v in d["stages"].items()}, d.get("primary_node_visits_tri_tests_per_ray"), d.get("layout_agreement_vs_exact"), d.get("gpu_over_cpu"))
PY
