#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python tools/sanitize_target.py ${1:-160} > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|done|diag" gpurun_out/sanitize_memcheck.log | head -20
tail -3 gpurun_out/sanitize_memcheck.log
