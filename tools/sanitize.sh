#!/bin/bash
# compute-sanitizer over the logN = 12 parity workloads (memcheck / racecheck for every kernel).
#   bash tools/sanitize.sh [outdir]     (run on a GPU box: gpurun -- 'bash tools/sanitize.sh gpurun_out/sanitize')
# Summaries go to <outdir>/{memcheck,racecheck}_<workload>.txt; the round's copy lives under profiles/.
OUT=${1:-gpurun_out/sanitize}
mkdir -p "$OUT"
CS=$(command -v compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
rc=0
for tool in memcheck racecheck; do
  for wl in ckks ckks15 bfv wide keygen; do
    extra=""
    [ "$tool" = racecheck ] && extra="--racecheck-report all"
    timeout 900 "$CS" --tool $tool $extra --error-exitcode 9 --print-limit 20 python tools/sanitize_workload.py $wl > "$OUT/${tool}_${wl}.txt" 2>&1
    r=$?
    echo "$tool $wl rc=$r: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload' "$OUT/${tool}_${wl}.txt" | tr '\n' ' ')"
    [ $r -ne 0 ] && rc=$r
  done
done
exit $rc
