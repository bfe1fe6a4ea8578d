#!/bin/bash
# 2 GPUs: multi-rank tests (peer mode with the record push, NCCL mode) and the bench line.
mkdir -p gpurun_out
L=gpurun_out/push_check_2gpu.log
{
  echo "== tests/test_gpu_multi.py, world 2"
  timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "distributed_equals_single and 2-" 2>&1 | tail -n 4
  echo "== bench N=2"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --stage-timings 2>&1 | grep "^{"
} > $L 2>&1
grep -v "^{" $L
python - <<PY
import json
for line in open("$L"):
    if line.startswith("{"):
        d = json.loads(line)
        print(d["value"], d["stages_ms"], d.get("parity", {}).get("ok"), d.get("e2e", {}).get("value"))
        for k, v in d.get("stages_ms_per_rank", {}).items(): print("  ", k, v)
PY
