"""Run under torchrun (or alone on one GPU): the oracle-free parity self-check of ectrans_b200/selfcheck.py
(reference golden vectors through the N-rank transform, N ranks == one rank bit for bit, same-direction transforms back
to back with one rank delayed, chunked host path).  ECT_P2P_ENTRY_BARRIER=0 reproduces the round-1 hazard."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ectrans_b200 as eb
from ectrans_b200 import selfcheck

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)


def fresh_uid():
    buf = torch.zeros(eb.ECT_NCCL_UID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(eb.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy().tobytes())


reps = int(os.environ.get("ECT_SELFCHECK_REPS", "1"))
allok = True
for i in range(reps):
    res = selfcheck.reduce(selfcheck.run(eb, world, rank, local, fresh_uid, stress_T=int(os.environ.get("ECT_SELFCHECK_T", "399"))), world, dev)
    allok = allok and res["ok"]
    if rank == 0:
        print(json.dumps({"n_gpus": world, "rep": i, "parity": res}), flush=True)
if rank == 0:
    print("SELFCHECK_OK" if allok else "SELFCHECK_FAIL", flush=True)
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if allok else 1)
