#!/bin/bash
# compute-sanitizer over every kernel family on small reads; logs under gpurun_out/sanitize_*.log (copied to profiles/)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_driver.py quick > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done" gpurun_out/sanitize_$tool.log | tail -3
done
