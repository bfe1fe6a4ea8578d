"""Where the host-API (e2e) time goes: raw PCIe rates next to the chunked / single-shot INV_TRANS and DIR_TRANS."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ectrans_b200 as eb

def t_ms(fn, n=3):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3

nbytes = 3 << 30 if not os.environ.get("PROBE_SKIP_RAW") else 1 << 20
hp = eb.PinnedArray((nbytes // 8,)); hp.array[...] = 1.0
hp2 = eb.PinnedArray((nbytes // 8,)); hp2.array[...] = 2.0
th, th2 = torch.from_numpy(hp.array), torch.from_numpy(hp2.array)
dv, dv2 = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda"), torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
print("H2D GB/s", nbytes / 1e6 / t_ms(lambda: dv.copy_(th, non_blocking=True)))
print("D2H GB/s", nbytes / 1e6 / t_ms(lambda: th.copy_(dv, non_blocking=True)))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def duplex():
    with torch.cuda.stream(s1): dv.copy_(th, non_blocking=True)
    with torch.cuda.stream(s2): th2.copy_(dv2, non_blocking=True)
print("duplex GB/s per direction", nbytes / 1e6 / t_ms(duplex))
del dv, dv2, hp, hp2, th, th2
torch.cuda.empty_cache()

T, N, nlev = 1279, 1280, 137
nuv, nsc = nlev, nlev + 1
nf = 2 * nuv + nsc
tr = eb.Transform(T, eb.octahedral_nloen(N))
h_in = [eb.PinnedArray((tr.nspec2, n)) for n in (nuv, nuv, nsc)]
rng = np.random.default_rng(0)
for x in h_in: x.array[...] = rng.uniform(-0.1, 0.1, x.array.shape)
h_gp = eb.PinnedArray((1, nf, tr.ngptot))
h_out = [eb.PinnedArray((tr.nspec2, n)) for n in (nuv, nuv, nsc)]
for chunk in os.environ.get("PROBE_CHUNKS", "0,48,24,96").split(","):
    os.environ["ECT_HOST_CHUNK_FIELDS"] = chunk
    inv = lambda: tr.inv_trans(h_in[0].array, h_in[1].array, h_in[2].array, out=h_gp.array)
    dr = lambda: tr.dir_trans(h_gp.array, nuv, nsc, out=tuple(x.array for x in h_out))
    print("chunk", chunk, "inv ms %.1f" % t_ms(inv, 2), "dir ms %.1f" % t_ms(dr, 2), flush=True)
