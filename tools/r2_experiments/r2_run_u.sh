#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_smoke.py > gpurun_out/r2u_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -12 gpurun_out/r2u_memcheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python tools/sanitize_smoke.py > gpurun_out/r2u_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -4 gpurun_out/r2u_synccheck.log
