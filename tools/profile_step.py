"""One warm-up + N device-resident inverse+direct steps, for ncu (launch list / --set full captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import argparse
import torch
import ectrans_b200 as eb
from bench import CONFIGS
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="TCo1279_O1280_L137")
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--precision", default="dp", choices=["dp", "sp"])
a = ap.parse_args()
tdt = torch.float64 if a.precision == "dp" else torch.float32
T, N, nlev, nfld = CONFIGS[a.config]
nuv, nsc = nlev, nlev * nfld + 1
dev = torch.device("cuda", 0)
tr = eb.Transform(T, eb.octahedral_nloen(N), stream=torch.cuda.current_stream().cuda_stream, precision=a.precision)
g = torch.Generator(device=dev); g.manual_seed(1)
mk = lambda n: (torch.rand((tr.nspec2, n), generator=g, device=dev, dtype=tdt) - 0.5) * 0.2
v, d, s = mk(nuv), mk(nuv), mk(nsc)
gp = torch.empty((1, 2 * nuv + nsc, tr.ngptot), dtype=tdt, device=dev)
out = (torch.empty_like(v), torch.empty_like(d), torch.empty_like(s))
hist = {"inv": [], "dir": []}
for i in range(a.warmup + a.steps):
    tr.inv_trans(v, d, s, out=gp)
    torch.cuda.synchronize(); hist["inv"].append(tr.timings())
    tr.dir_trans(gp, nuv, nsc, out=out)
    torch.cuda.synchronize(); hist["dir"].append(tr.timings())
for k in ("inv", "dir"):
    h = hist[k][a.warmup:]
    print("timings(%s) min over %d steps:" % (k, len(h)),
          {n: round(min(t[n] for t in h), 3) for n in ("prologue", "legendre", "fourier", "epilogue", "total")},
          "max total %.3f" % max(t["total"] for t in h))
