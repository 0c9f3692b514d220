#!/bin/bash
# compute-sanitizer passes over scripts/sanitize_probe.py (run under gpurun)
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout ${SANITIZE_TIMEOUT:-420} compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer_$tool.log \
      python scripts/sanitize_probe.py > gpurun_out/sanitizer_$tool.out 2>&1
  echo "$tool exit $?"; tail -2 gpurun_out/sanitizer_$tool.out; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid" gpurun_out/sanitizer_$tool.log | head -8
done
