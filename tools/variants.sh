#!/bin/bash
# development aid: time build variants / run-time knobs of the table path on the C3 batch
cd "$(dirname "$0")/.."
T=${T:-200}
run() { echo "=== $*"; timeout "$T" env "$@" 2>&1 | grep -E "cold-L2|evals/s|max \||first call|Error|error|tables:"; }
run GWIN_B200_LIB=$PWD/gwin_b200/libgwin_b200.so python tools/quick_time.py 16384 256 table,direct
for v in "$@"; do
  run GWIN_B200_LIB=$PWD/gwin_b200/libgwin_$v.so python tools/quick_time.py 16384 256 table
done
