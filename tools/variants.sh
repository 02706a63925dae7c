#!/bin/bash
# development aid: time build variants / run-time knobs of the table path on the C3 batch
cd "$(dirname "$0")/.."
T=${T:-200}
run() { echo "=== $*"; timeout "$T" env "$@" 2>&1 | grep -E "cold-L2|evals/s|max \||first call|Error|error|tables:"; }
run GWIN_DEBUG=1 GWIN_B200_LIB=$PWD/gwin_b200/libgwin_b200.so python tools/quick_time.py 16384 256 table,direct
for v in "$@"; do
  run GWIN_DEBUG=1 GWIN_B200_LIB=$PWD/gwin_b200/libgwin_$v.so python tools/quick_time.py 16384 256 table,direct
done
run GWIN_EXACT_SORT=1 python tools/quick_time.py 16384 256 table
run GWIN_FINE_SLOTS=3 python tools/quick_time.py 16384 256 table
run GWIN_TABLE_FINE=0 python tools/quick_time.py 16384 256 table
