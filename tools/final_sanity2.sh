#!/bin/bash
# One GPU, final tree: memcheck over every kernel family (incl. the push path), then the GPU test-suite.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_case.py > gpurun_out/sanitize_memcheck_b.log 2>&1
echo "exit code memcheck: $?" >> gpurun_out/sanitize_memcheck_b.log
tail -n 9 gpurun_out/sanitize_memcheck_b.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest_gpu.log 2>&1; tail -n 4 gpurun_out/final_pytest_gpu.log
