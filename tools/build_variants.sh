#!/bin/bash
# A/B builds of the library with different -D knobs: tools/build_variants.sh name1:"-DX=1 -DY=0" name2:"..."
# Output: vistrace_b200/variants/lib_<name>.so (git-ignored, travels with gpurun; select with VT_LIB=<path>).
set -e
cd "$(dirname "$0")/.."
mkdir -p vistrace_b200/variants
for spec in "$@"; do
    name="${spec%%:*}"; flags="${spec#*:}"
    ( make -C vistrace_b200/csrc OUT="$PWD/vistrace_b200/variants/lib_${name}.so" EXTRA="$flags" 2>&1 | grep -E "error|k_traverse_compactILb0ELb0ELb0ELb1E" -A2 | grep -E "error|registers|spill" | sed "s/^/[$name] /" ) &
done
wait
ls -la vistrace_b200/variants/
