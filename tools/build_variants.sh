#!/bin/bash
# Developer tool: build library variants (dfgnn_b200/variants/lib<name>.so) from "name:flags" specs.
#   tools/build_variants.sh "c2:-DDFGNN_SPMM_C=2" "nw4:-DDFGNN_KNW=4 -DDFGNN_STAGE_CAP=1024"
cd "$(dirname "$0")/../dfgnn_b200/csrc" || exit 1
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  ( make -s BUILD=_build_$name LIB=../variants/lib$name.so EXTRA="$flags" -j4 > /tmp/variant_$name.log 2>&1 \
      && echo "built $name" || { echo "FAILED $name"; tail -5 /tmp/variant_$name.log; } ) &
done
wait
