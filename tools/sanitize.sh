#!/bin/bash
# compute-sanitizer over the small-shape pass (tools/sanitize.py); logs -> gpurun_out/r02_sanitizer_<tool>.log
# usage (GPU box): bash tools/sanitize.sh [memcheck racecheck synccheck]
set -u
mkdir -p gpurun_out
TOOLS=${@:-memcheck racecheck synccheck}
for t in $TOOLS; do
  echo "== compute-sanitizer --tool $t" | tee gpurun_out/r02_sanitizer_$t.log
  timeout 900 compute-sanitizer --tool $t --print-limit 20 python tools/sanitize.py >> gpurun_out/r02_sanitizer_$t.log 2>&1
  echo "exit code $?" >> gpurun_out/r02_sanitizer_$t.log
  tail -4 gpurun_out/r02_sanitizer_$t.log
done
