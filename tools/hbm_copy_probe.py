import torch, time
a=torch.empty(1<<30, dtype=torch.uint8, device='cuda'); b=torch.empty_like(a)
for _ in range(3): b.copy_(a)
torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): b.copy_(a)
e1.record(); torch.cuda.synchronize()
print("d2d copy GB/s (r+w)", 2*20*(1<<30)/1e9/(e0.elapsed_time(e1)/1e3))
