"""Host <-> device copy bandwidth per GPU, one rank alone and all ranks at once (run under torchrun).  Explains the
end-to-end (host-pointer) numbers at N > 1: the eight GPUs of a box share the host's memory system and PCIe uplinks."""
import os, sys, json, time
import torch, torch.distributed as dist

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device=dev)
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n, dtype=torch.uint8, device=dev)
s2 = torch.cuda.Stream()


def bw(kind, reps=4):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if kind in ("h2d", "both"):
            d.copy_(h, non_blocking=True)
        if kind in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return reps * n * (2 if kind == "both" else 1) / dt / 1e9


res = {}
for kind in ("h2d", "d2h", "both"):
    bw(kind, 1)
    # one rank alone
    alone = None
    for r in range(min(world, 2)):
        if world > 1:
            dist.barrier()
        if rank == r:
            v = bw_alone = None
        if rank == r:
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(4):
                if kind in ("h2d", "both"): d.copy_(h, non_blocking=True)
                if kind in ("d2h", "both"):
                    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
            torch.cuda.synchronize()
            alone = 4 * n * (2 if kind == "both" else 1) / (time.perf_counter() - t0) / 1e9
        if world > 1:
            dist.barrier()
    allv = bw(kind)
    t = torch.tensor([allv, alone if alone is not None else -1.0], dtype=torch.float64, device=dev)
    if world > 1:
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
    else:
        g = [t]
    res[kind] = {"all_ranks_at_once_GBps_per_gpu": [round(float(x[0]), 1) for x in g],
                 "one_rank_alone_GBps": [round(float(x[1]), 1) for x in g if float(x[1]) > 0]}
if rank == 0:
    print(json.dumps({"n_gpus": world, "pcie_probe": res}))
if world > 1:
    dist.destroy_process_group()
