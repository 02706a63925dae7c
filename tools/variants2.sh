#!/bin/bash
# development aid: time run-time knobs / build variants given as "ENV=val ... " strings, one per argument
cd "$(dirname "$0")/.."
T=${T:-300}
for spec in "$@"; do
  echo "=== $spec"
  timeout "$T" env $spec python bench.py --steps 50 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], [(l['kernel'][:22], round(l['ms'],4)) for l in d['roofline']['launches']])"
done
