#!/bin/bash
# 8 GPUs: default step (record push, cost-balanced bands) as the full bench line, then two other per-latitude constants
# of the band cost model, then the direct remote stores for comparison.
N=8
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --stage-timings "$@" 2>&1 | grep "^{"; }
L=gpurun_out/push_bench_n8.log
echo "== default" > $L;               run >> $L
echo "== ECT_BAND_PAD=0.08" >> $L;    ECT_BAND_PAD=0.08 run --no-parity --no-e2e >> $L
echo "== ECT_BAND_PAD=0.26" >> $L;    ECT_BAND_PAD=0.26 run --no-parity --no-e2e >> $L
echo "== ECT_FFT_PUSH=0" >> $L;       ECT_FFT_PUSH=0 run --no-parity --no-e2e >> $L
python - <<PY
import json
for line in open("$L"):
    if line.startswith("{"):
        d = json.loads(line)
        print(d["value"], d["stages_ms"], d.get("parity", {}).get("ok"), d.get("e2e", {}).get("value"))
        for k, v in d.get("stages_ms_per_rank", {}).items(): print("  ", k, v)
    else: print(line.strip())
PY
