#!/bin/bash
# development aid: runs the stages of a GPU check one by one, each under its own timeout, logging as it goes
cd "$(dirname "$0")/.."
export GWIN_DEBUG=1
run() { echo "=== $*"; timeout "$T" "$@"; echo "=== rc=$?"; }
T=${T:-150}
for spec in "256 64 recurrence" "256 64 table" "2048 64 table,direct" "1024 256 table,direct" "16384 256 table,recurrence"; do
  run python tools/quick_time.py $spec
done
