#!/usr/bin/env bash
# compute-sanitizer pass over every kernel family: memcheck, racecheck, synccheck, initcheck.
#   tools/sanitize.sh TAG      logs -> gpurun_out/TAG_sanitize_<tool>.log (+ a one-line summary per tool)
set -u
TAG="${1:-r02}"
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
for TOOL in memcheck racecheck synccheck initcheck; do
  LOG=gpurun_out/${TAG}_sanitize_${TOOL}.log
  timeout 1500 $SAN --tool $TOOL --print-limit 20 --error-exitcode 7 python tools/sanitize_case.py > $LOG 2>&1
  echo "$TOOL rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $LOG | tail -1)" | tee -a gpurun_out/${TAG}_sanitize_summary.log
done
