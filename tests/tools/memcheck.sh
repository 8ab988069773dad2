#!/bin/bash
# compute-sanitizer memcheck over a small pass through every kernel family; summary -> gpurun_out/memcheck.log
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python tests/tools/memcheck_case.py > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?"
tail -6 gpurun_out/memcheck.log
