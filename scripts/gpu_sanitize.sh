#!/bin/bash
# compute-sanitizer over every kernel family (race / sync checking of the kernels that
# hand-roll inter-warp hand-offs: K2's progress flags, the bit-vector kernels' shared-memory
# CAS tables).  Logs under gpurun_out/sanitize/; summaries are copied into profiles/.
#   usage: scripts/gpu_sanitize.sh [tools] [families]      default: "memcheck racecheck synccheck" all
TOOLS=${1:-"memcheck racecheck synccheck"}
FAMS=${2:-""}
OUT=gpurun_out/sanitize
mkdir -p $OUT
for t in $TOOLS; do
  timeout 1500 compute-sanitizer --tool $t --print-limit 20 --error-exitcode 9 \
      python scripts/sanitize_driver.py $FAMS > $OUT/$t.log 2>&1
  echo "$t exit $?" | tee -a $OUT/$t.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|driver done" $OUT/$t.log | tail -3
done
