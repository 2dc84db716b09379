#!/bin/bash
# One compute-sanitizer tool per GPU call (B200_PROFILING.md): memcheck on the smallest end-to-end case.
set -u
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_plain.log 2>&1 && \
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python __graft_entry__.py smoke > gpurun_out/memcheck_smoke.log 2>&1
echo "memcheck exit $?"; tail -6 gpurun_out/memcheck_smoke.log
