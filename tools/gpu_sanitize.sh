#!/bin/bash
# one compute-sanitizer tool per gpurun call: bash tools/gpu_sanitize.sh memcheck|racecheck [extra args of sanitize_case.py]
TOOL=${1:-memcheck}; shift
mkdir -p gpurun_out
python tools/sanitize_case.py "$@" > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
tail -1 gpurun_out/sanitize_plain.log | cut -c1-400
timeout 900 compute-sanitizer --tool $TOOL --print-limit 20 python tools/sanitize_case.py "$@" > gpurun_out/sanitize_$TOOL.log 2>&1; echo "sanitizer exit $?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_CASE_OK|Error|error|hazard" gpurun_out/sanitize_$TOOL.log | head -20
