#!/bin/bash
# One GPU: the slot + push store path of the direct Fourier stage, forced on a single rank (ECT_FFT_PUSH=2).
mkdir -p gpurun_out
{
  echo "== new test"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "push_slots" 2>&1 | tail -n 4
  echo "== whole parity file with ECT_FFT_PUSH=2"; ECT_FFT_PUSH=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -n 4
  echo "== bench default"; timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-parity --stage-timings 2>&1 | grep "^{"
  echo "== bench ECT_FFT_PUSH=2"; ECT_FFT_PUSH=2 timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-parity --stage-timings 2>&1 | grep "^{"
} > gpurun_out/push_check_1gpu.log 2>&1
tail -n 30 gpurun_out/push_check_1gpu.log
