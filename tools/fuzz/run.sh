#!/bin/bash
# Mutation fuzz of the host-only ingestion code (vt_vtf.cpp, vt_mdl.cpp, vt_bsp.cpp) under AddressSanitizer + UBSan: synthetic files of
# the test generators (tests/test_vtf.py, tests/mdl_files.py, tests/bsp_files.py) with random byte / word / offset mutations and
# truncations, every file read into an exact-size heap buffer; then the hierarchy code (hierarchy_driver.cpp) on hostile geometry and
# on node arrays with mutated child references, counts, bounds and primitive indices.  No GPU, no CUDA.
# usage: tools/fuzz/run.sh [seed] [files per parser]      (a finding prints an ASan / UBSan report and the script exits 1)
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"; ROOT="$HERE/../.."
SEED=${1:-1}; N=${2:-3000}
W=$(mktemp -d)
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer -fopenmp -ffp-contract=off \
    -I"$ROOT/vistrace_b200/csrc" -I"$ROOT/include" -I/usr/local/cuda/include \
    "$HERE/ingest_driver.cpp" "$ROOT/vistrace_b200/csrc/vt_bsp.cpp" "$ROOT/vistrace_b200/csrc/vt_mdl.cpp" "$ROOT/vistrace_b200/csrc/vt_vtf.cpp" -o "$W/driver"
python "$HERE/make_corpus.py" "$SEED" "$N" "$W/corpus"
for k in bsp vtf; do ls "$W"/corpus/$k/* | xargs -n 500 "$W/driver" $k; done
ls "$W"/corpus/mdl/*.mdl | sed 's/\.mdl$//' | awk '{print $0".mdl "$0".vvd "$0".vtx"}' | xargs -n 600 "$W/driver" mdl
# the hierarchy code (builders, flatten, compact / quad layouts, host refit, reinsertion) on hostile geometry and hostile trees
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer -fopenmp -ffp-contract=off \
    -I"$ROOT/vistrace_b200/csrc" -I"$ROOT/include" -I/usr/local/cuda/include \
    "$HERE/hierarchy_driver.cpp" "$ROOT/vistrace_b200/csrc/vt_bvh_build.cpp" "$ROOT/vistrace_b200/csrc/vt_bvh_ploc.cpp" \
    "$ROOT/vistrace_b200/csrc/vt_bvh_collapse.cpp" "$ROOT/vistrace_b200/csrc/vt_bvh_reinsert.cpp" -o "$W/hdriver"
timeout 1200 "$W/hdriver" "$SEED" $(( N / 60 + 5 ))
rm -rf "$W"
echo "fuzz: seed $SEED, $N files per parser, $(( N / 60 + 5 )) scenes x 13 trees: no sanitizer finding"
