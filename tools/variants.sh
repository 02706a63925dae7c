#!/bin/bash
# development aid: time build variants / run-time knobs of the table path on the C3 batch (bench.py's own split)
cd "$(dirname "$0")/.."
T=${T:-300}
run() { echo "=== $*"; timeout "$T" env "$@" python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], [(l['kernel'][:22], round(l['ms'],4)) for l in d['roofline']['launches']])"; }
run GWIN_B200_LIB=$PWD/gwin_b200/libgwin_b200.so
for v in "$@"; do
  run GWIN_B200_LIB=$PWD/gwin_b200/libgwin_$v.so
done
