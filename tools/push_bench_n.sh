#!/bin/bash
# N GPUs: the step with the direct stage's record push (default in peer mode) and with the direct 16-byte remote stores.
N=${1:-4}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu --stage-timings "$@" 2>&1 | grep "^{"; }
echo "== push (default), parity gate on" > gpurun_out/push_bench_n$N.log
run >> gpurun_out/push_bench_n$N.log
echo "== ECT_FFT_PUSH=0" >> gpurun_out/push_bench_n$N.log
ECT_FFT_PUSH=0 run --no-parity >> gpurun_out/push_bench_n$N.log
python - <<PY
import json
for line in open("gpurun_out/push_bench_n$N.log"):
    if line.startswith("{"):
        d = json.loads(line)
        print(d["value"], d["stages_ms"], d.get("parity", {}).get("ok"))
        for k, v in d.get("stages_ms_per_rank", {}).items(): print("  ", k, v)
    else: print(line.strip())
PY
