"""sp Legendre contraction on tcgen05 (csrc/legendre_tc.cu) against the FP64 DMMA path (ECT_SP_TC=0) and the oracle.
Run under `timeout`: every mbarrier wait in the kernel is bounded (trap), so a wrong descriptor cannot hang the GPU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import ectrans_b200 as eb
import ectrans_oracle as eo

rel = lambda a, b: float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))
ok = True
for T, N, nuv, nsc in ((47, 48, 2, 3), (159, 160, 5, 70), (399, 400, 20, 21)):
    nloen = eb.octahedral_nloen(N)
    s = eo.setup(T, 2 * N, nloen)
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    vor, div, sc = f32(eo.random_spectral(s, nuv, 1, zero00=True)), f32(eo.random_spectral(s, nuv, 2, zero00=True)), f32(eo.random_spectral(s, nsc, 3))
    ref = eo.inv_trans(s, vor, div, sc)
    nf = 2 * nuv + nsc
    rv, rd, rs = eo.dir_trans(s, f32(ref[:nf]), nuv, nsc)
    T_ = lambda a: np.ascontiguousarray(a.T).astype(np.float32)
    res = {}
    for mode in ("0", "1"):
        os.environ["ECT_SP_TC"] = mode
        tr = eb.Transform(T, nloen, precision="sp")
        t0 = time.time()
        gp = tr.inv_trans(T_(vor), T_(div), T_(sc))
        e_inv = max(rel(gp[0, i].astype(np.float64), ref[i]) for i in range(nf))
        ov, od, os_ = tr.dir_trans(f32(ref[:nf])[None].astype(np.float32), nuv, nsc)
        e_dir = max(rel(a.T.astype(np.float64), b) for a, b in ((ov, rv), (od, rd), (os_, rs)))
        res[mode] = (e_inv, e_dir, tr.timings()["legendre"])
        tr.release()
    print(f"T{T} fields {nf}: DMMA inv {res['0'][0]:.2e} dir {res['0'][1]:.2e} | tcgen05 3xTF32 inv {res['1'][0]:.2e} dir {res['1'][1]:.2e}", flush=True)
    ok = ok and res["1"][0] < 1e-5 and res["1"][1] < 1e-5
print("TC_CHECK_OK" if ok else "TC_CHECK_FAIL")
sys.exit(0 if ok else 1)
