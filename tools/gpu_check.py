"""First-contact GPU diagnostic: table, inverse, direct against the oracle with verbose errors."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import ectrans_b200 as eb
import ectrans_oracle as eo

def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))

def check(T, nloen, nuv, nsc, opts, label, nproma=0, tables=True):
    nloen = np.asarray(nloen)
    t0 = time.time(); tr = eb.Transform(T, nloen); t1 = time.time()
    s = eo.setup(T, nloen.size, nloen)
    print(f"[{label}] T={T} ndgl={nloen.size} setup gpu {t1-t0:.2f}s table {tr.info.table_bytes/1e6:.1f} MB", flush=True)
    if tables:
        worst = 0.0
        for ml in range(0, tr.nump, max(1, tr.nump // 7)):
            m = int(tr.myms[ml])
            if s.ndglu[m] == 0: continue
            worst = max(worst, np.abs(tr.legendre_table(ml, 0).T - s.ps[m]).max(), np.abs(tr.legendre_table(ml, 1).T - s.pa[m]).max() if s.pa[m].size else 0)
        print(f"   table max abs diff {worst:.3e}")
    vor = eo.random_spectral(s, nuv, 1, zero00=True) if nuv else None
    div = eo.random_spectral(s, nuv, 2, zero00=True) if nuv else None
    sc = eo.random_spectral(s, nsc, 3) if nsc else None
    ref = eo.inv_trans(s, vor, div, sc, **opts)
    tt = lambda a: None if a is None else np.ascontiguousarray(a.T)
    gp = tr.inv_trans(tt(vor), tt(div), tt(sc), nproma=nproma, **opts)
    nblk, nf, npr = gp.shape
    gpf = gp.transpose(1, 0, 2).reshape(nf, nblk * npr)[:, :tr.ngptot]
    print(f"   INV rel L2 {rel(gpf, ref):.3e}  maxabs {np.abs(gpf-ref).max():.3e}  timings {tr.timings()}")
    for i in range(nf):
        r = rel(gpf[i], ref[i])
        if r > 1e-11: print(f"     field {i}: rel {r:.3e}")
    # direct on oracle grid data (u, v, scalars)
    iu = (nuv if opts.get('vorgp') else 0) + (nuv if opts.get('divgp') else 0)
    gin = ref[iu:iu + 2 * nuv + nsc]
    rv, rd, rs = eo.dir_trans(s, gin, nuv, nsc)
    g3 = np.zeros((nblk, 2 * nuv + nsc, npr)); flat = np.zeros((2 * nuv + nsc, nblk * npr)); flat[:, :tr.ngptot] = gin
    g3[:] = flat.reshape(2 * nuv + nsc, nblk, npr).transpose(1, 0, 2)
    ov, od, os_ = tr.dir_trans(g3, nuv, nsc, nproma=nproma)
    for nm, a, b in [("vor", ov, rv), ("div", od, rd), ("sc", os_, rs)]:
        if a is not None: print(f"   DIR {nm} rel L2 {rel(a.T, b):.3e} maxabs {np.abs(a.T-b).max():.3e}")
    print(f"   timings {tr.timings()}")
    if nsc: print("   specnorm rel", rel(tr.specnorm(tt(sc)), eo.specnorm(s, sc)))
    tr.release()

if __name__ == "__main__":
    print("fp64 peak TF/s: dmma", eb.measure_fp64_peak(0), "dfma", eb.measure_fp64_peak(1), flush=True)
    check(79, eb.octahedral_nloen(80), 0, 1, {}, "scalar T79")
    check(79, eb.octahedral_nloen(80), 2, 3, dict(scders=True, vorgp=True, divgp=True, uvder=True), "full T79")
    nl = np.load(os.path.join(ROOT, "tests/golden/lon_number_by_lat.npy"))
    check(148, nl, 0, 1, {}, "golden grid")
    check(79, eb.octahedral_nloen(80), 3, 2, dict(scders=True), "nproma T79", nproma=1000)
    check(159, eb.octahedral_nloen(160), 10, 11, dict(scders=True, uvder=True), "T159")
    check(399, eb.octahedral_nloen(400), 4, 5, {}, "TCo399", tables=False)
