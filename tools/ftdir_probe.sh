#!/bin/bash
# Timing experiments on k_fourier<direct> (results are wrong with the switches on): what the chirp loads from L2 and the
# record stores cost.
mkdir -p gpurun_out
L=gpurun_out/ftdir_probe.log; : > $L
for dbg in 0 8 16 24; do
  echo "== ECT_FFT_DBG=$dbg" >> $L
  ECT_FFT_DBG=$dbg timeout 300 python bench.py --steps 4 --warmup 2 --no-e2e --no-cpu --no-parity --stage-timings 2>&1 | grep "^{" | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print(d['value'], d['stages_ms'], d.get('stages_ms_per_rank'))" >> $L
done
cat $L
