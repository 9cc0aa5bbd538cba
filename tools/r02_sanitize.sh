#!/bin/bash
# compute-sanitizer passes over one small run of every launch class (tools/sanitize_step.py); run under gpurun.
# NOTE: the round-2 GPU pool refuses compute-sanitizer ("closed on this pool"): there the same properties are covered by
# tests/test_gpu_memory_safety.py (workspace poison, guard bands, >2^31-element tensors).  Kept for pools that allow it.
# usage: tools/r02_sanitize.sh <tag> [tool ...]      (default tools: memcheck synccheck initcheck)
tag=${1:-r02}; shift
tools=${@:-memcheck synccheck initcheck}
mkdir -p gpurun_out
timeout 120 python tools/sanitize_step.py > gpurun_out/${tag}_sanitize_plain.log 2>&1; echo "plain rc=$?"
for t in $tools; do
  timeout ${SAN_TIMEOUT:-240} compute-sanitizer --tool $t --print-limit 40 --error-exitcode 9 \
      --log-file gpurun_out/${tag}_sanitize_${t}.txt python tools/sanitize_step.py > gpurun_out/${tag}_sanitize_${t}.log 2>&1
  echo "$t rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${tag}_sanitize_${t}.txt | tail -1)"
  grep -c "^ok" gpurun_out/${tag}_sanitize_${t}.log
done
