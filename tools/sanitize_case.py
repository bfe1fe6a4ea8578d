"""Small transforms through every kernel family, for compute-sanitizer (tools/sanitize.sh): smooth and chirp-z rows on
the undivided kernel (direct record stores and the slot + push path), chirp-z rows on the CTA-pair kernel (clusters +
distributed shared memory), the FP64 DMMA
contraction, the tcgen05 contraction of sp handles (TMA + TMEM), adjoints, SPECNORM, GPNORM_TRANS."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ectrans_b200 as eb

rng = np.random.default_rng(0)
T, N = 31, 32
nloen = eb.octahedral_nloen(N)
# last two cases: the direct stage's slot + push record stores, forced on one rank (ECT_FFT_PUSH=2)
for prec, cz, push in (("dp", "0", "0"), ("dp", "1", "0"), ("sp", "0", "0"), ("dp", "0", "2"), ("sp", "0", "2")):
    os.environ["ECT_FFT_CZ"] = cz
    os.environ["ECT_FFT_PUSH"] = push
    tr = eb.Transform(T, nloen, precision=prec)
    dt = np.float64 if prec == "dp" else np.float32
    mk = lambda n: rng.uniform(-0.1, 0.1, (tr.nspec2, n)).astype(dt)
    vor, div, sc = mk(2), mk(2), mk(3)
    for a in (vor, div, sc):
        a[1:2 * (T + 1):2] = 0
    gp = tr.inv_trans(vor, div, sc, scders=True, uvder=True, vorgp=True, divgp=True, nproma=97)
    gp2 = tr.inv_trans(vor, div, sc)
    out = tr.dir_trans(gp2, 2, 3)
    err = float(np.abs(out[2] - sc).max())          # scalars: the round trip is exact to rounding
    tr.specnorm(sc); tr.gpnorm_trans(gp2)
    tr.inv_transad(gp2, 2, 3); tr.dir_transad(vor, div, sc)
    print(f"{prec} ECT_FFT_CZ={cz} ECT_FFT_PUSH={push}: round trip max abs {err:.2e}", flush=True)
    assert err < (1e-9 if prec == "dp" else 1e-4)          # reduced grid: the white-spectrum round trip is exact to ~1e-10 only
    tr.release()
print("SANITIZE_CASE_OK")
