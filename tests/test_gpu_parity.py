"""Parity of the CUDA path (through the C ABI) against the oracle and the reference's golden vectors.

Tolerances (BASELINE.json north_star): relative grid-point / spectral L2 difference <= 1e-12 in dp;
golden vectors abs 1e-10 (tests/test_ectrans4py/test_ectrans4py.py:16); round trip <= 100 eps
(src/programs/ectrans-benchmark.F90:847-871)."""
import numpy as np
import pytest

import ectrans_oracle as eo

pytestmark = pytest.mark.gpu
TOL = 1e-12
T_ = lambda a: None if a is None else np.ascontiguousarray(a.T)


def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


@pytest.fixture(scope="module")
def eb(built):
    import ectrans_b200
    return ectrans_b200


def unblock(gp, ngptot):
    nblk, nf, npr = gp.shape
    return gp.transpose(1, 0, 2).reshape(nf, nblk * npr)[:, :ngptot]


def block(flat, nproma):
    nf, n = flat.shape
    nblk = (n + nproma - 1) // nproma
    pad = np.zeros((nf, nblk * nproma))
    pad[:, :n] = flat
    return np.ascontiguousarray(pad.reshape(nf, nblk, nproma).transpose(1, 0, 2))


# ---------------------------------------------------------------------------------------------
def test_golden_vectors(eb, golden):
    tr = eb.Transform(148, golden["nloen"])
    nl = golden["nloen"]
    gpref = np.concatenate([golden["gp_latlon"][i, :nl[i]] for i in range(150)])
    gp = tr.inv_trans(spscalar=golden["sp"][:, None])
    d = gp[0, 0] - gpref
    assert abs(d.max()) < 1e-10 and abs(d.min()) < 1e-10
    _, _, sp = tr.dir_trans(gpref[None, None, :], 0, 1)
    d = sp[:, 0] - golden["sp"]
    assert abs(d.max()) < 1e-10 and abs(d.min()) < 1e-10
    assert (tr.ngptot, tr.nspec2 // 2) == (33052, 11175)
    np.testing.assert_array_equal(tr.nmen, golden["nmen"])
    tr.release()


def test_legendre_table_matches_oracle(eb):
    T, N = 159, 160
    nloen = eb.octahedral_nloen(N)
    tr = eb.Transform(T, nloen)
    s = eo.setup(T, 2 * N, nloen)
    for ml in range(0, tr.nump, 7):
        m = int(tr.myms[ml])
        for par, ref in ((0, s.ps[m]), (1, s.pa[m])):
            if ref.size == 0:
                continue
            got = tr.legendre_table(ml, par).T
            # device FMA contraction only: a few ulp of the largest entry
            assert np.abs(got - ref).max() <= 2e-13 * max(np.abs(ref).max(), 1.0)
    tr.release()


CASES = [
    # T, N, nuv, nsc, options, nproma
    (79, 80, 10, 11, {}, 0),                                                       # BASELINE config 0 (10 levels, 1 scalar field)
    (79, 80, 0, 1, {}, 0),
    (79, 80, 3, 0, dict(vorgp=True, divgp=True, uvder=True), 0),
    (79, 80, 2, 3, dict(scders=True, vorgp=True, divgp=True, uvder=True), 1000),
    (79, 80, 1, 1, dict(scders=True), 17),                                         # ragged last block, odd field counts
    (159, 160, 5, 16, dict(scders=True, uvder=True), 0),                           # BASELINE config 1 shape, fewer levels
    (47, 48, 130, 131, {}, 0),                                                     # many fields: several field tiles
]


@pytest.mark.parametrize("T,N,nuv,nsc,opts,nproma", CASES)
def test_inv_dir_against_oracle(eb, T, N, nuv, nsc, opts, nproma):
    nloen = eb.octahedral_nloen(N)
    tr = eb.Transform(T, nloen)
    s = eo.setup(T, 2 * N, nloen)
    vor = eo.random_spectral(s, nuv, 1, zero00=True) if nuv else None
    div = eo.random_spectral(s, nuv, 2, zero00=True) if nuv else None
    sc = eo.random_spectral(s, nsc, 3) if nsc else None
    ref = eo.inv_trans(s, vor, div, sc, **opts)
    gp = tr.inv_trans(T_(vor), T_(div), T_(sc), nproma=nproma, **opts)
    got = unblock(gp, tr.ngptot)
    assert got.shape == ref.shape
    for i in range(ref.shape[0]):
        assert rel(got[i], ref[i]) < TOL, (i, rel(got[i], ref[i]))
    iu = (nuv if opts.get("vorgp") else 0) + (nuv if opts.get("divgp") else 0)
    gin = ref[iu:iu + 2 * nuv + nsc]
    rv, rd, rs = eo.dir_trans(s, gin, nuv, nsc)
    npr = tr.ngptot if nproma <= 0 else nproma
    ov, od, os_ = tr.dir_trans(block(gin, npr), nuv, nsc, nproma=nproma)
    for a, b in ((ov, rv), (od, rd), (os_, rs)):
        if b is not None:
            assert rel(a.T, b) < TOL
            # Im(m=0) = 0 and, for vor/div, coefficient (0,0) = 0  (updsp_mod.F90:120-133, updspb_mod.F90:98-117)
            assert np.all(a[1:2 * (T + 1):2] == 0.0)
    if nuv:
        assert np.all(ov[0] == 0.0) and np.all(od[0] == 0.0)
    if nsc:
        assert rel(tr.specnorm(T_(sc)), eo.specnorm(s, sc)) < 1e-13
        pmet = 1.0 / (1.0 + np.arange(T + 1.0)) ** 2         # SPECNORM's optional metric PMET(0:NSMAX)
        assert rel(tr.specnorm(T_(sc), pmet=pmet), eo.specnorm(s, sc, pmet)) < 1e-13
    tr.release()


def _grid(kind, golden):
    if kind == "F24":                   # full (regular) Gaussian grid
        return np.full(48, 96, dtype=np.int32)
    if kind == "golden150":             # the irregular reduced grid of the reference's ectrans4py test data
        return np.asarray(golden["nloen"], dtype=np.int32)
    if kind == "N32classic":            # classic reduced Gaussian grid: odd row lengths (27, 45, 75) -> chirp-z with sign-flipped reflections
        return np.asarray(N32_CLASSIC + N32_CLASSIC[::-1], dtype=np.int32)
    return None


N32_CLASSIC = [20, 27, 36, 40, 45, 50, 60, 64, 72, 75, 80, 90, 90, 96, 100, 108, 108, 120, 120, 120, 128, 128, 128, 128,
               128, 128, 128, 128, 128, 128, 128, 128]


# NMEN rules of setup_geom_mod.F90:44-78 other than the cubic octahedral one: linear (T >= NDGL-1), quadratic
# (T >= 2 NDGL / 3 - 1), a full grid, an irregular reduced grid; vor/div + scalars with derivatives each
GRID_CASES = [("O48", 95), ("O48", 63), ("F24", 47), ("F24", 30), ("golden150", 99), ("golden150", 60),
              ("N32classic", 63), ("N32classic", 42), ("N32classic", 21)]


@pytest.mark.parametrize("kind,T", GRID_CASES)
def test_other_grids_and_truncation_rules(eb, golden, kind, T):
    import ectrans_b200
    nloen = ectrans_b200.octahedral_nloen(48) if kind == "O48" else _grid(kind, golden)
    ndgl = int(nloen.size)
    tr = eb.Transform(T, nloen)
    s = eo.setup(T, ndgl, nloen)
    np.testing.assert_array_equal(tr.nmen, s.nmen)
    np.testing.assert_array_equal(tr.ndglu, s.ndglu)
    vor = eo.random_spectral(s, 2, 1, zero00=True); div = eo.random_spectral(s, 2, 2, zero00=True)
    sc = eo.random_spectral(s, 3, 3)
    opts = dict(scders=True, uvder=True, vorgp=True, divgp=True)
    ref = eo.inv_trans(s, vor, div, sc, **opts)
    got = unblock(tr.inv_trans(T_(vor), T_(div), T_(sc), **opts), tr.ngptot)
    for i in range(ref.shape[0]):
        assert rel(got[i], ref[i]) < TOL, (i, rel(got[i], ref[i]))
    gin = ref[4:4 + 4 + 3]
    rv, rd, rs = eo.dir_trans(s, gin, 2, 3)
    ov, od, os_ = tr.dir_trans(gin[None], 2, 3)
    for a, b in ((ov, rv), (od, rd), (os_, rs)):
        assert rel(a.T, b) < TOL
    tr.release()


def test_benchmark_input_roundtrip(eb):
    """ectrans-benchmark's own check: Re psi(4,19) = 1 everywhere, inverse + direct, norm error <= 100 eps."""
    T, N, nlev = 159, 160, 9
    tr = eb.Transform(T, eb.octahedral_nloen(N))
    s = eo.setup(T, 2 * N, eb.octahedral_nloen(N), tables=False)
    sp = eo.benchmark_spectral(s, nlev)
    z = T_(sp)
    for it in range(2):
        gp = tr.inv_trans(z, z, z)
        ov, od, os_ = tr.dir_trans(gp[:, 2 * nlev:] if False else gp, nlev, nlev)
        z = os_
    n0 = eo.specnorm(s, sp)
    n1 = tr.specnorm(os_)
    assert np.abs(n1 / n0 - 1).max() <= 100 * np.finfo(float).eps
    tr.release()


def test_constant_field(eb):
    # tests/transi/transi_test_program.c:76-81,150-164
    tr = eb.Transform(47, eb.octahedral_nloen(48))
    gp = np.stack([np.full(tr.ngptot, c) for c in (1.0, 2.0, 3.0, 4.0)])[None]
    _, _, sp = tr.dir_trans(gp, 0, 4)
    for i, c in enumerate((1.0, 2.0, 3.0, 4.0)):
        assert abs(sp[0, i] - c) < 1e-13
        assert np.abs(sp[1:, i]).max() < 1e-13
    tr.release()


def test_call_mode_2(eb):
    """PGPUV / PGP3A / PGP2 layouts (inv_trans.h:84-104, trltog_mod.F90:579-731)."""
    T, N, nlev, n3, n2 = 47, 48, 4, 2, 3
    nloen = eb.octahedral_nloen(N)
    tr = eb.Transform(T, nloen)
    s = eo.setup(T, 2 * N, nloen)
    vor = eo.random_spectral(s, nlev, 1, zero00=True)
    div = eo.random_spectral(s, nlev, 2, zero00=True)
    sc2 = eo.random_spectral(s, n2, 3)
    sc3 = eo.random_spectral(s, nlev * n3, 4)            # field index = j3 * nlev + lev
    ref = eo.inv_trans(s, vor, div, np.concatenate([sc2, sc3]), scders=True, uvder=True, vorgp=True)
    nproma = 500
    nblk = (tr.ngptot + nproma - 1) // nproma
    sp3 = np.ascontiguousarray(sc3.reshape(n3, nlev, s.nspec2).transpose(0, 2, 1))      # (fld, nspec2, lev)
    gpuv = np.zeros((nblk, 5, nlev, nproma)); gp2 = np.zeros((nblk, 3 * n2, nproma)); gp3 = np.zeros((nblk, 3 * n3, nlev, nproma))
    tr.inv_trans_raw(memspace=0, nproma=nproma, scders=1, vorgp=1, divgp=0, uvder=1,
                     spvor=T_(vor), spdiv=T_(div), nuv=nlev, spsc2=T_(sc2), nsc2=n2,
                     spsc3a=sp3, nsc3a_lev=nlev, nsc3a_fld=n3, gpuv=gpuv, gp2=gp2, gp3a=gp3)
    flat = lambda a: a.reshape(nblk, -1, nproma).transpose(1, 0, 2).reshape(a.size // (nblk * nproma), -1)[:, :tr.ngptot]
    # reference order: vor u v | sc2 sc3 | nsd(sc2) nsd(sc3) | du dv | ewd(sc2) ewd(sc3)
    nsc = n2 + nlev * n3
    o = 0
    r_vor, r_u, r_v = ref[o:o + nlev], ref[o + nlev:o + 2 * nlev], ref[o + 2 * nlev:o + 3 * nlev]; o += 3 * nlev
    r_sc = ref[o:o + nsc]; o += nsc
    r_ns = ref[o:o + nsc]; o += nsc
    r_du, r_dv = ref[o:o + nlev], ref[o + nlev:o + 2 * nlev]; o += 2 * nlev
    r_ew = ref[o:o + nsc]
    assert rel(flat(gpuv), np.concatenate([r_vor, r_u, r_v, r_du, r_dv])) < TOL
    assert rel(flat(gp2), np.concatenate([r_sc[:n2], r_ns[:n2], r_ew[:n2]])) < TOL
    assert rel(flat(gp3), np.concatenate([r_sc[n2:], r_ns[n2:], r_ew[n2:]])) < TOL
    # direct, call mode 2
    guv = np.ascontiguousarray(gpuv[:, 1:3]); g2 = np.ascontiguousarray(gp2[:, :n2]); g3 = np.ascontiguousarray(gp3[:, :n3])
    ovor = np.zeros((s.nspec2, nlev)); odiv = np.zeros_like(ovor); o2 = np.zeros((s.nspec2, n2)); o3 = np.zeros((n3, s.nspec2, nlev))
    tr.dir_trans_raw(memspace=0, nproma=nproma, nuv=nlev, gpuv=guv, gp2=g2, nsc2=n2, gp3a=g3, nsc3a_lev=nlev, nsc3a_fld=n3,
                     spvor=ovor, spdiv=odiv, spsc2=o2, spsc3a=o3)
    rv, rd, rs = eo.dir_trans(s, np.concatenate([r_u, r_v, r_sc]), nlev, nsc)
    assert rel(ovor.T, rv) < TOL and rel(odiv.T, rd) < TOL and rel(o2.T, rs[:n2]) < TOL
    assert rel(o3.transpose(0, 2, 1).reshape(n3 * nlev, -1), rs[n2:]) < TOL
    tr.release()


@pytest.fixture
def forced_chunks(monkeypatch):
    """Host-pointer calls split into many small pipelined field chunks (normally only multi-GB calls are)."""
    monkeypatch.setenv("ECT_HOST_CHUNK_FIELDS", "3")
    monkeypatch.setenv("ECT_HOST_CHUNK_MIN_BYTES", "0")


def test_call_mode_2_chunked(eb, forced_chunks):
    test_call_mode_2(eb)


@pytest.mark.parametrize("nproma", [0, 777])
def test_host_chunked_pipeline(eb, forced_chunks, nproma):
    """The pipelined host path (field chunks, copy streams) against the oracle and against the single-shot path:
    every option that changes the field lists, blocked and unblocked."""
    T, N, nuv, nsc = 79, 80, 7, 10
    nloen = eb.octahedral_nloen(N)
    tr = eb.Transform(T, nloen)
    s = eo.setup(T, 2 * N, nloen)
    vor = eo.random_spectral(s, nuv, 1, zero00=True); div = eo.random_spectral(s, nuv, 2, zero00=True)
    sc = eo.random_spectral(s, nsc, 3)
    opts = dict(scders=True, uvder=True, vorgp=True, divgp=True)
    ref = eo.inv_trans(s, vor, div, sc, **opts)
    gp = tr.inv_trans(T_(vor), T_(div), T_(sc), nproma=nproma, **opts)
    assert tr.timings()["launches"] > 60          # several chunks ran
    flat = gp.transpose(1, 0, 2).reshape(gp.shape[1], -1)[:, :tr.ngptot]
    for i in range(ref.shape[0]):
        assert rel(flat[i], ref[i]) < TOL, i
    iu = 2 * nuv
    gin = np.ascontiguousarray(gp[:, iu:iu + 2 * nuv + nsc])
    ov, od, os_ = tr.dir_trans(gin, nuv, nsc, nproma=nproma)
    rv, rd, rs = eo.dir_trans(s, ref[iu:iu + 2 * nuv + nsc], nuv, nsc)
    assert rel(ov.T, rv) < TOL and rel(od.T, rd) < TOL and rel(os_.T, rs) < TOL
    import os
    os.environ["ECT_HOST_CHUNK_FIELDS"] = "0"      # single shot
    gp1 = tr.inv_trans(T_(vor), T_(div), T_(sc), nproma=nproma, **opts)
    flat1 = gp1.transpose(1, 0, 2).reshape(gp1.shape[1], -1)[:, :tr.ngptot]
    assert rel(flat, flat1) < 1e-13
    tr.release()


def test_adjoints_against_matrix_oracle(eb):
    """INV_TRANSAD / DIR_TRANSAD against the oracle's transposed forward matrices (small truncation), the
    accumulate semantics of INV_TRANSAD (prfi1bad_mod.F90:91-108) and the not-implemented options."""
    T, N, nuv, nsc = 10, 12, 2, 3
    nloen = eb.octahedral_nloen(N)
    tr = eb.Transform(T, nloen)
    s = eo.setup(T, 2 * N, nloen)
    rng = np.random.default_rng(5)
    y = rng.uniform(-1, 1, (2 * nuv + nsc, s.ngptot))
    rv, rd, rs = eo.inv_transad(s, y, nuv, nsc)
    ov, od, os_ = tr.inv_transad(y[None], nuv, nsc)
    assert rel(ov.T, rv) < TOL and rel(od.T, rd) < TOL and rel(os_.T, rs) < TOL
    assert np.all(ov[1:2 * (T + 1):2] == 0) and np.all(os_[1:2 * (T + 1):2] == 0)      # Im(m = 0)
    pre = (np.ones_like(ov), np.full_like(od, 2.0), np.full_like(os_, -1.0))
    tr.inv_transad(y[None], nuv, nsc, out=pre)
    assert rel(pre[0] - 1.0, ov) < 1e-13 and rel(pre[1] - 2.0, od) < 1e-13 and rel(pre[2] + 1.0, os_) < 1e-13
    vor = eo.random_spectral(s, nuv, 1, zero00=True); div = eo.random_spectral(s, nuv, 2, zero00=True)
    sc = eo.random_spectral(s, nsc, 3)
    ref = eo.dir_transad(s, vor, div, sc)
    g = tr.dir_transad(T_(vor), T_(div), T_(sc))
    for i in range(ref.shape[0]):
        assert rel(g[0, i], ref[i]) < TOL, i
    with pytest.raises(eb.EctError, match="not implemented"):
        a = eb._InvArgs(); a.scders = 1; a.nscalar = 1
        eb._check(eb.lib().ect_inv_transad(tr.handle, eb.C.byref(a)), "ect_inv_transad")
    tr.release()


@pytest.mark.parametrize("chunked", [False, True])
def test_adjoint_identity_reference_test(eb, monkeypatch, chunked):
    """The reference's own adjoint tests (tests/trans/test_invtrans_adjoint.F90:152-222, test_dirtrans_adjoint.F90):
    T159, seeded random data, <F x, y> = <x, F* y> to 20000 eps with its inner products; host arrays (also through the
    pipelined chunk path) and device arrays."""
    import torch
    if chunked:
        monkeypatch.setenv("ECT_HOST_CHUNK_FIELDS", "3")
        monkeypatch.setenv("ECT_HOST_CHUNK_MIN_BYTES", "0")
    T, N, nuv, nsc = 159, 160, 4, 5
    nloen = eb.octahedral_nloen(N)
    tr = eb.Transform(T, nloen)
    rng = np.random.default_rng(9)
    mk = lambda n: 0.1 * (1 - 2 * rng.random((tr.nspec2, n)))
    vor, div, sc = mk(nuv), mk(nuv), mk(nsc)
    w = np.zeros(tr.nspec2)
    for m in range(T + 1):
        a0, n = int(tr.nasm0[m]), 2 * (T - m + 1)
        if m == 0:
            w[a0:a0 + n:2] = 1.0
        else:
            w[a0:a0 + n] = 2.0
    w = w[:, None]
    y = 1 - 2 * rng.random((1, 2 * nuv + nsc, tr.ngptot))
    tol = 20000 * np.finfo(float).eps
    fx = tr.inv_trans(vor, div, sc)
    va, da, sa = tr.inv_transad(y, nuv, nsc)
    lhs, rhs = np.sum(fx * y), np.sum(w * vor * va) + np.sum(w * div * da) + np.sum(w * sc * sa)
    assert abs(lhs - rhs) <= tol * abs(lhs)
    dv, dd, ds = tr.dir_trans(y, nuv, nsc)
    ga = tr.dir_transad(vor, div, sc)
    lhs, rhs = np.sum(w * dv * vor) + np.sum(w * dd * div) + np.sum(w * ds * sc), np.sum(y * ga)
    assert abs(lhs - rhs) <= tol * abs(lhs)
    if not chunked:      # device arrays give the same bits as host arrays
        vd, dd2, sd = tr.inv_transad(torch.from_numpy(y).cuda(), nuv, nsc)
        tr.synchronize()               # device-pointer calls are asynchronous on the handle's stream
        assert np.array_equal(vd.cpu().numpy(), va) and np.array_equal(sd.cpu().numpy(), sa)
        gd = tr.dir_transad(torch.from_numpy(vor).cuda(), torch.from_numpy(div).cuda(), torch.from_numpy(sc).cuda())
        tr.synchronize()
        assert np.array_equal(gd.cpu().numpy(), ga)
    tr.release()


def test_device_memspace_matches_host(eb):
    import torch
    T, N = 79, 80
    tr = eb.Transform(T, eb.octahedral_nloen(N), stream=torch.cuda.current_stream().cuda_stream)
    rng = np.random.default_rng(0)
    sc = rng.uniform(-1, 1, (tr.nspec2, 6)); sc[1:2 * (T + 1):2] = 0
    gp_h = tr.inv_trans(spscalar=sc)
    gp_d = tr.inv_trans(spscalar=torch.from_numpy(sc).cuda())
    # asynchronous back-to-back calls must not disturb each other (per-call tables live in a ring)
    outs = [tr.dir_trans(gp_d, 0, 6)[2] for _ in range(6)]
    tr.synchronize()
    assert np.array_equal(gp_d.cpu().numpy(), gp_h)
    sp_h = tr.dir_trans(gp_h, 0, 6)[2]
    for o in outs:
        assert np.array_equal(o.cpu().numpy(), sp_h)
    tr.release()


def test_linearity_and_roundtrip_large(eb):
    """Size-independent properties at a size the oracle is not run on (TCo639-class truncation)."""
    T, N, nf = 639, 640, 4
    tr = eb.Transform(T, eb.octahedral_nloen(N))
    rng = np.random.default_rng(1)
    n = np.concatenate([np.repeat(np.arange(m, T + 1), 2) for m in range(T + 1)]).astype(float)
    a = rng.uniform(-1, 1, (tr.nspec2, nf)) / (1 + n[:, None]) ** 2
    b = rng.uniform(-1, 1, (tr.nspec2, nf)) / (1 + n[:, None]) ** 2
    for x in (a, b):
        x[1:2 * (T + 1):2] = 0
    ga, gb = tr.inv_trans(spscalar=a), tr.inv_trans(spscalar=b)
    gab = tr.inv_trans(spscalar=2.0 * a - 3.0 * b)
    assert rel(gab, 2.0 * ga - 3.0 * gb) < 1e-13
    back = tr.dir_trans(ga, 0, nf)[2]
    # cubic grid: low-order-dominated spectrum survives the round trip to rounding level
    assert rel(back, a) < 1e-11
    assert np.abs(tr.specnorm(back) / tr.specnorm(a) - 1).max() < 1e-12
    tr.release()


def test_errors_are_codes_not_aborts(eb):
    import ctypes
    tr = eb.Transform(47, eb.octahedral_nloen(48))
    a = eb._InvArgs()
    a.nuv = 2          # spvor / spdiv missing
    assert eb.lib().ect_inv_trans(tr.handle, ctypes.byref(a)) == -3
    a = eb._DirArgs(); a.nscalar = 1
    assert eb.lib().ect_dir_trans(tr.handle, ctypes.byref(a)) == -3
    tr.release()
    assert eb.lib().ect_inv_trans(tr.handle, ctypes.byref(eb._InvArgs())) == -8


def test_ectrans4py_face(eb, golden):
    """The reference's own ectrans4py test, verbatim calls (tests/test_ectrans4py/test_ectrans4py.py:109-158)."""
    from ectrans_b200 import ectrans4py
    nl = golden["nloen"]
    gpdata = np.concatenate([golden["gp_latlon"][i, :nl[i]] for i in range(150)])
    sizes = ectrans4py.trans_inq4py(150, 148, len(nl), nl, 10)
    assert sizes[0:2] == (33052, 11175)
    np.testing.assert_array_equal(sizes[2], golden["nmen"])
    gp = ectrans4py.sp2gp_gauss4py(150, 148, 10, int(sum(nl)), len(nl), nl, len(golden["sp"]), False, False, golden["sp"])[0]
    assert np.abs(gp - gpdata).max() < 1e-10
    sp = ectrans4py.gp2sp_gauss4py(11175 * 2, 150, 148, 10, len(nl), nl, len(gpdata), False, gpdata)
    assert np.abs(sp - golden["sp"]).max() < 1e-10
    # LREORDER=True: the same golden field given / returned in the 'model' coefficient order (sp2gp_gauss4py.F90:82-108)
    model = ectrans4py._to_model_order(148, golden["sp"], len(golden["sp"]))
    gp_m = ectrans4py.sp2gp_gauss4py(150, 148, 10, int(sum(nl)), len(nl), nl, len(model), False, True, model)[0]
    assert np.array_equal(gp_m, gp)
    sp_m = ectrans4py.gp2sp_gauss4py(11175 * 2, 150, 148, 10, len(nl), nl, len(gpdata), True, gpdata)
    assert np.array_equal(ectrans4py._from_model_order(148, sp_m), np.where(np.arange(sp.size) < 2 * 149, np.where(np.arange(sp.size) % 2 == 1, 0.0, sp), sp))
    nspec = sum(148 + 2 - im for im in range(149))
    knmeng, weights, polys = ectrans4py.get_legendre_assets(150, 148, len(nl), nspec, nl, 10)
    assert abs(sum(weights) - 1.0) < 1e-10
    assert polys.shape == (75, nspec)


def test_transi_face(eb):
    """transi C API through ctypes: the reference's transi_test_program flow (tests/transi/transi_test_program.c:64-165)."""
    import ctypes as C
    L = eb.lib()

    IP, DP = C.POINTER(C.c_int), C.POINTER(C.c_double)

    class Trans(C.Structure):          # include/transi_b200.h struct Trans_t (= the reference's, transi.h:701-850)
        _fields_ = ([("ndgl", C.c_int), ("nloen", IP), ("nlon", C.c_int), ("nsmax", C.c_int), ("llam", C.c_int),
                     ("lsplit", C.c_int), ("llatlon", C.c_int), ("flt", C.c_int), ("fft", C.c_int),
                     ("readfp", C.c_char_p), ("writefp", C.c_char_p), ("cache", C.c_void_p), ("cachesize", C.c_size_t),
                     ("myproc", C.c_int), ("nproc", C.c_int), ("handle", C.c_int)]
                    + [(n, C.c_int) for n in ("nspec", "nspec2", "nspec2g", "nspec2mx", "nump", "ngptot", "ngptotg", "ngptotmx")]
                    + [("ngptotl", IP), ("nmyms", IP), ("nasm0", IP), ("nprtrw", C.c_int)]
                    + [(n, IP) for n in ("numpp", "npossp", "nptrms", "nallms", "ndim0g", "nvalue")]
                    + [(n, C.c_int) for n in ("n_regions_NS", "n_regions_EW", "my_region_NS", "my_region_EW")]
                    + [("n_regions", IP), ("nfrstlat", IP), ("nlstlat", IP), ("nfrstloff", C.c_int), ("nptrlat", IP),
                       ("nptrfrstlat", IP), ("nptrlstlat", IP), ("nptrfloff", C.c_int), ("nsta", IP), ("nonl", IP),
                       ("ldsplitlat", IP), ("nprtrns", C.c_int), ("nultpp", IP), ("nptrls", IP), ("nnmeng", IP),
                       ("rmu", DP), ("rgw", DP), ("rpnm", DP), ("nlei3", C.c_int), ("nspolegl", C.c_int), ("npms", IP),
                       ("rlapin", DP), ("ndglu", IP), ("pexwn", C.c_double), ("peywn", C.c_double), ("pweight", DP),
                       ("ndgux", C.c_int), ("nmsmax", C.c_int), ("mvalue", IP)])

    class Inv(C.Structure):
        _fields_ = [("rspscalar", C.c_void_p), ("rspvor", C.c_void_p), ("rspdiv", C.c_void_p), ("rmeanu", C.c_void_p),
                    ("rmeanv", C.c_void_p), ("rgp", C.c_void_p), ("nproma", C.c_int), ("nscalar", C.c_int),
                    ("nvordiv", C.c_int), ("lscalarders", C.c_int), ("luvder_EW", C.c_int), ("lvordivgp", C.c_int),
                    ("ngpblks", C.c_int), ("lglobal", C.c_int), ("trans", C.POINTER(Trans)), ("count", C.c_int)]

    class Dir(C.Structure):
        _fields_ = [("rgp", C.c_void_p), ("rspscalar", C.c_void_p), ("rspvor", C.c_void_p), ("rspdiv", C.c_void_p),
                    ("rmeanu", C.c_void_p), ("rmeanv", C.c_void_p), ("nproma", C.c_int), ("nscalar", C.c_int),
                    ("nvordiv", C.c_int), ("ngpblks", C.c_int), ("lglobal", C.c_int), ("trans", C.POINTER(Trans)),
                    ("count", C.c_int)]

    L.new_invtrans.restype = Inv
    L.new_invtrans.argtypes = [C.POINTER(Trans)]
    L.new_dirtrans.restype = Dir
    L.new_dirtrans.argtypes = [C.POINTER(Trans)]
    L.trans_error_msg.restype = C.c_char_p
    t = Trans()
    assert L.trans_new(C.byref(t)) == 0
    nl = eb.octahedral_nloen(24)
    assert L.trans_set_resol(C.byref(t), 48, nl.ctypes.data_as(C.POINTER(C.c_int))) == 0
    assert L.trans_set_trunc(C.byref(t), 23) == 0
    assert L.trans_setup(C.byref(t)) == 0
    assert t.nspec2 == 24 * 25 and t.ngptot == int(nl.sum())
    assert L.trans_inquire(C.byref(t), b"nvalue,nmyms,nasm0,rgw") == 0
    assert t.nvalue[0] == 0 and t.nvalue[2] == 1 and t.nasm0[0] == 1
    assert L.trans_inquire(C.byref(t), b"bogus") == -4
    # every variable list of the reference's transi_test_program.c:51-54
    for vl in (b"numpp,ngptotl,nmyms,nasm0,npossp,nptrms,nallms,ndim0g,nvalue",
               b"nfrstlat,nlstlat,nptrlat,nptrfrstlat,nptrlstlat,nsta,nonl,ldsplitlat", b"nultpp,nptrls,nnmeng",
               b"rmu,rgw,npms,rlapin,ndglu"):
        assert L.trans_inquire(C.byref(t), vl) == 0, vl
    assert (t.npossp[0], t.npossp[1]) == (1, t.nspec2 + 1) and t.ndim0g[0] == 1 and t.ndim0g[1] == 2 * 24 + 1
    assert t.nsta[0] == 1 and t.nonl[0] == int(nl[0]) and t.nonl[47] == int(nl[47]) and t.ldsplitlat[0] == 0
    assert t.nfrstlat[0] == 1 and t.nlstlat[0] == 48 and t.nptrlat[5] == 6 and t.n_regions_NS == 1 and t.nprtrns == 1
    assert t.ndglu[0] == 24 and t.npms[0] == 1 and t.npms[1] == 26 and t.nlei3 == 24 and t.nspolegl == sum(25 - m for m in range(24))
    assert t.rlapin[0] == 0.0 and t.rlapin[1] == 0.0 and abs(t.rlapin[2] + 6371229.0 ** 2 / 2.0) < 1e-3
    assert L.trans_set_radius(C.c_double(6371229.0)) == 0 and L.trans_set_radius(C.c_double(1.0)) == -2
    assert L.trans_set_nprtrv(1) == 0 and L.trans_set_nprtrv(2) == -2 and L.trans_set_leq_regions(1) == 0
    nscalar = 2
    rgp = np.zeros((nscalar, t.ngptot))
    rgp[0] = 1.0
    rgp[1] = 2.0
    rsp = np.zeros((t.nspec2, nscalar))
    d = L.new_dirtrans(C.byref(t))
    d.nscalar = nscalar
    d.rgp = rgp.ctypes.data
    d.rspscalar = rsp.ctypes.data
    assert L.trans_dirtrans(C.byref(d)) == 0
    assert L.trans_dirtrans(C.byref(d)) == -5            # stale argument struct
    assert abs(rsp[0, 0] - 1.0) < 1e-13 and abs(rsp[0, 1] - 2.0) < 1e-13 and np.abs(rsp[1:]).max() < 1e-13
    out = np.zeros_like(rgp)
    i = L.new_invtrans(C.byref(t))
    i.nscalar = nscalar
    i.rspscalar = rsp.ctypes.data
    i.rgp = out.ctypes.data
    assert L.trans_invtrans(C.byref(i)) == 0
    assert np.abs(out - rgp).max() < 1e-13
    i2 = L.new_invtrans(C.byref(t))
    i2.nscalar = 1
    assert L.trans_invtrans(C.byref(i2)) == -3           # missing rspscalar
    assert b"missing" in L.trans_error_msg(-3)

    # adjoints (transi.h:354, 491) and gather / distribute (transi.h:520-616)
    class InvAdj(C.Structure):
        _fields_ = [("rspscalar", C.c_void_p), ("rspvor", C.c_void_p), ("rspdiv", C.c_void_p), ("rmeanu", C.c_void_p),
                    ("rmeanv", C.c_void_p), ("rgp", C.c_void_p), ("nproma", C.c_int), ("nscalar", C.c_int),
                    ("nvordiv", C.c_int), ("lscalarders", C.c_int), ("luvder_EW", C.c_int), ("lvordivgp", C.c_int),
                    ("ngpblks", C.c_int), ("lglobal", C.c_int), ("trans", C.POINTER(Trans)), ("count", C.c_int)]

    class Gath(C.Structure):
        _fields_ = [("rgpg", C.c_void_p), ("rgp", C.c_void_p), ("nto", C.c_void_p), ("nproma", C.c_int), ("nfld", C.c_int),
                    ("ngpblks", C.c_int), ("trans", C.POINTER(Trans)), ("count", C.c_int)]

    class GathSp(C.Structure):
        _fields_ = [("rspecg", C.c_void_p), ("rspec", C.c_void_p), ("nto", C.c_void_p), ("nfld", C.c_int),
                    ("trans", C.POINTER(Trans)), ("count", C.c_int)]

    for fn, ty in (("new_invtrans_adj", InvAdj), ("new_dirtrans_adj", Dir), ("new_gathgrid", Gath), ("new_distgrid", Gath),
                   ("new_gathspec", GathSp), ("new_distspec", GathSp)):
        getattr(L, fn).restype = ty
        getattr(L, fn).argtypes = [C.POINTER(Trans)]
    rng = np.random.default_rng(3)
    x = rng.uniform(-1, 1, (t.nspec2, nscalar)); x[1:48:2] = 0
    y = rng.uniform(-1, 1, (nscalar, t.ngptot))
    fx = np.zeros_like(y)
    i3 = L.new_invtrans(C.byref(t)); i3.nscalar = nscalar; i3.rspscalar = x.ctypes.data; i3.rgp = fx.ctypes.data
    assert L.trans_invtrans(C.byref(i3)) == 0
    ya = np.zeros_like(x)
    ia = L.new_invtrans_adj(C.byref(t)); ia.nscalar = nscalar; ia.rspscalar = ya.ctypes.data; ia.rgp = y.ctypes.data
    assert L.trans_invtrans_adj(C.byref(ia)) == 0
    w = np.full((t.nspec2, 1), 2.0); w[0:48:2] = 1.0; w[1:48:2] = 0.0
    assert abs(np.sum(fx * y) - np.sum(w * x * ya)) <= 20000 * np.finfo(float).eps * abs(np.sum(fx * y))
    dy = np.zeros_like(x)
    d2 = L.new_dirtrans(C.byref(t)); d2.nscalar = nscalar; d2.rgp = y.ctypes.data; d2.rspscalar = dy.ctypes.data
    assert L.trans_dirtrans(C.byref(d2)) == 0
    xa = np.zeros_like(y)
    da = L.new_dirtrans_adj(C.byref(t)); da.nscalar = nscalar; da.rspscalar = x.ctypes.data; da.rgp = xa.ctypes.data
    assert L.trans_dirtrans_adj(C.byref(da)) == 0
    assert abs(np.sum(w * dy * x) - np.sum(y * xa)) <= 20000 * np.finfo(float).eps * abs(np.sum(y * xa))
    ia2 = L.new_invtrans_adj(C.byref(t)); ia2.nscalar = 1; ia2.lscalarders = 1; ia2.rspscalar = ya.ctypes.data; ia2.rgp = y.ctypes.data
    assert L.trans_invtrans_adj(C.byref(ia2)) == -2       # derivative options of the adjoint: not implemented
    nto = np.ones(nscalar, dtype=np.int32)
    gg = np.zeros_like(y)
    g = L.new_gathgrid(C.byref(t)); g.rgp = y.ctypes.data; g.rgpg = gg.ctypes.data; g.nto = nto.ctypes.data; g.nfld = nscalar
    assert L.trans_gathgrid(C.byref(g)) == 0 and np.array_equal(gg, y)
    xs = rng.uniform(-1, 1, (t.nspec2, nscalar)); sg = np.zeros_like(xs)
    gs = L.new_gathspec(C.byref(t)); gs.rspec = xs.ctypes.data; gs.rspecg = sg.ctypes.data; gs.nto = nto.ctypes.data; gs.nfld = nscalar
    assert L.trans_gathspec(C.byref(gs)) == 0
    xz = xs.copy(); xz[1:48:2] = 0
    assert np.array_equal(sg, xz)
    back = np.zeros_like(xs)
    ds = L.new_distspec(C.byref(t)); ds.rspecg = sg.ctypes.data; ds.rspec = back.ctypes.data; ds.nto = nto.ctypes.data; ds.nfld = nscalar   # nto: the slot of DistSpec_t.nfrom
    assert L.trans_distspec(C.byref(ds)) == 0 and np.array_equal(back, xz)
    bad = np.full(nscalar, 2, dtype=np.int32)
    g2 = L.new_gathgrid(C.byref(t)); g2.rgp = y.ctypes.data; g2.rgpg = gg.ctypes.data; g2.nto = bad.ctypes.data; g2.nfld = nscalar
    assert L.trans_gathgrid(C.byref(g2)) == -1            # task 2 of 1
    assert L.trans_delete(C.byref(t)) == 0

    class VdUv(C.Structure):
        _fields_ = [("rspvor", C.c_void_p), ("rspdiv", C.c_void_p), ("rspu", C.c_void_p), ("rspv", C.c_void_p),
                    ("nfld", C.c_int), ("nsmax", C.c_int), ("ncoeff", C.c_int), ("count", C.c_int)]
    L.new_vordiv_to_UV.restype = VdUv
    Tn = 21
    so = eo.setup(Tn, 32, eo.octahedral_nloen(16), tables=False)
    vo = eo.random_spectral(so, 2, 7, zero00=True); di = eo.random_spectral(so, 2, 8, zero00=True)
    ur, vr = eo.vordiv_to_uv(so, vo, di)
    vo_t, di_t = np.ascontiguousarray(vo.T), np.ascontiguousarray(di.T)
    uo, vv = np.zeros_like(vo_t), np.zeros_like(vo_t)
    a = L.new_vordiv_to_UV(); a.rspvor = vo_t.ctypes.data; a.rspdiv = di_t.ctypes.data; a.rspu = uo.ctypes.data; a.rspv = vv.ctypes.data
    a.nfld = 2; a.nsmax = Tn; a.ncoeff = so.nspec2
    assert L.trans_vordiv_to_UV(C.byref(a)) == 0
    assert rel(uo.T, ur) < TOL and rel(vv.T, vr) < TOL
    assert L.trans_vordiv_to_UV(C.byref(a)) == -5          # stale
    b = L.new_vordiv_to_UV(); b.nsmax = Tn
    assert L.trans_vordiv_to_UV(C.byref(b)) == -3          # ncoeff missing


@pytest.mark.parametrize("T,N,nuv,nsc,opts", [(79, 80, 3, 4, dict(scders=True, uvder=True)), (159, 160, 2, 5, {})])
def test_single_precision_face(eb, T, N, nuv, nsc, opts):
    """sp build (BASELINE configs 2 and 4 are sp): float arrays at the boundary, tolerance 1e-5 relative
    (north_star); the reference keeps m = 0 in fp64 for sp (ledir_mod.F90:133-171), here every m is."""
    nloen = eb.octahedral_nloen(N)
    tr = eb.Transform(T, nloen, precision="sp")
    s = eo.setup(T, 2 * N, nloen)
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    vor = f32(eo.random_spectral(s, nuv, 1, zero00=True)); div = f32(eo.random_spectral(s, nuv, 2, zero00=True))
    sc = f32(eo.random_spectral(s, nsc, 3))
    ref = eo.inv_trans(s, vor, div, sc, **opts)
    gp = tr.inv_trans(T_(vor).astype(np.float32), T_(div).astype(np.float32), T_(sc).astype(np.float32), **opts)
    assert gp.dtype == np.float32
    # sp handles compute the Fourier stage in float (as the reference's sp build does); stated tolerance 1e-5
    for i in range(ref.shape[0]):
        assert rel(gp[0, i].astype(np.float64), ref[i]) < 5e-6
    gin = f32(ref[:2 * nuv + nsc])
    rv, rd, rs = eo.dir_trans(s, gin, nuv, nsc)
    ov, od, os_ = tr.dir_trans(gin[None].astype(np.float32), nuv, nsc)
    for a, b in ((ov, rv), (od, rd), (os_, rs)):
        assert a.dtype == np.float32 and rel(a.T.astype(np.float64), b) < 5e-6
    assert rel(tr.specnorm(T_(sc).astype(np.float32)), eo.specnorm(s, sc)) < 1e-6
    tr.release()


def test_tco2559_rows_sp(eb):
    """BASELINE config 4 grid (TCo2559 / O2560, sp; rows up to 10256 points, chirp-z length 16384 in float,
    the longest rows without a staging area) on one GPU with a reduced field count: the benchmark's
    single-harmonic input against its analytic answer, linearity and the round trip, tolerance 1e-5."""
    T, N = 2559, 2560
    nloen = eb.octahedral_nloen(N)
    tr = eb.Transform(T, nloen, precision="sp")
    nf = 4
    rng = np.random.default_rng(11)
    n = np.concatenate([np.repeat(np.arange(m, T + 1), 2) for m in range(T + 1)]).astype(float)
    a = (rng.uniform(-1, 1, (tr.nspec2, nf)) / (1 + n[:, None]) ** 1.5).astype(np.float32)
    a[1:2 * (T + 1):2] = 0
    b = np.zeros_like(a)
    b[int(tr.nasm0[4]) + 2 * (19 - 4)] = 1.0
    ga, gb = tr.inv_trans(spscalar=a), tr.inv_trans(spscalar=b)
    assert ga.dtype == np.float32
    off = np.concatenate([[0], np.cumsum(nloen)])
    for j in (0, 5, 1000, 2047, 2559, 2560, 4000, 5119):
        nlon = int(nloen[j])
        row = gb[0, 0, off[j]:off[j] + nlon].astype(np.float64)
        if tr.nmen[j] >= 4:
            p = eo.supolf(4, 19, tr.rmu[j])[19, 0]
            assert np.abs(row - 2 * p * np.cos(2 * np.pi * 4 * np.arange(nlon) / nlon)).max() < 1e-5
    gab = tr.inv_trans(spscalar=(a - 2.5 * b).astype(np.float32))
    assert rel(gab.astype(np.float64), ga.astype(np.float64) - 2.5 * gb) < 1e-5
    back = tr.dir_trans(ga, 0, nf)[2]
    assert rel(back.astype(np.float64), a.astype(np.float64)) < 1e-5
    assert np.abs(tr.specnorm(back) / tr.specnorm(a) - 1).max() < 1e-5
    tr.release()


def test_full_size_grid_properties(eb):
    """BASELINE headline grid (TCo1279 / O1280, every row length 20..5136 incl. all chirp-z classes) with a
    reduced field count: linearity, round trip, norm preservation, and the benchmark's single-harmonic
    input against its analytic answer 2 P_19^4(mu) cos(4 lambda) (ectrans-benchmark.F90:1389-1415)."""
    T, N = 1279, 1280
    nloen = eb.octahedral_nloen(N)
    tr = eb.Transform(T, nloen)
    nf = 6
    rng = np.random.default_rng(7)
    n = np.concatenate([np.repeat(np.arange(m, T + 1), 2) for m in range(T + 1)]).astype(float)
    a = rng.uniform(-1, 1, (tr.nspec2, nf)) / (1 + n[:, None]) ** 2
    a[1:2 * (T + 1):2] = 0
    b = np.zeros_like(a)
    b[int(tr.nasm0[4]) + 2 * (19 - 4)] = 1.0                      # Re psi(4, 19) = 1 in every field
    ga, gb = tr.inv_trans(spscalar=a), tr.inv_trans(spscalar=b)
    gab = tr.inv_trans(spscalar=a - 2.5 * b)
    assert rel(gab, ga - 2.5 * gb) < 1e-13
    off = np.concatenate([[0], np.cumsum(nloen)])
    for j in (0, 3, 500, 1023, 1279, 1280, 2000, 2559):            # pole, chirp-z and smooth rows, both hemispheres
        nlon = int(nloen[j])
        row = gb[0, 0, off[j]:off[j] + nlon]
        if tr.nmen[j] >= 4:
            p = eo.supolf(4, 19, tr.rmu[j])[19, 0]
            assert np.abs(row - 2 * p * np.cos(2 * np.pi * 4 * np.arange(nlon) / nlon)).max() < 1e-12
        else:
            assert np.abs(row).max() == 0.0
    back = tr.dir_trans(ga, 0, nf)[2]
    assert rel(back, a) < 1e-11
    assert np.abs(tr.specnorm(back) / tr.specnorm(a) - 1).max() < 1e-12
    bb = tr.dir_trans(gb, 0, nf)[2]
    assert np.abs(tr.specnorm(bb) / tr.specnorm(b) - 1).max() <= 100 * np.finfo(float).eps
    tr.release()


# the reference's own benchmark-as-test matrix (tests/CMakeLists.txt:219-326): T47 / O48, --niter 2 --check 100
BENCH_MATRIX = [
    dict(nlev=1, nfld=0),
    dict(nlev=20, nfld=10),
    dict(nlev=20, nfld=10, scders=True, uvder=True),
    dict(nlev=20, nfld=10, scders=True, uvder=True, vordiv=True),
    dict(nlev=20, nfld=10, nproma=16),
]


@pytest.mark.parametrize("cfg", BENCH_MATRIX)
def test_reference_benchmark_matrix(eb, cfg):
    T, N = 47, 48
    nlev, nfld = cfg["nlev"], cfg["nfld"]
    nuv, nsc = nlev, nlev * nfld + 1                     # ectrans-benchmark.F90:450-478
    nloen = eb.octahedral_nloen(N)
    tr = eb.Transform(T, nloen)
    s = eo.setup(T, 2 * N, nloen, tables=False)
    vor = T_(eo.benchmark_spectral(s, nuv)); div = vor.copy(); sc = T_(eo.benchmark_spectral(s, nsc))
    n0v, n0s = tr.specnorm(vor), tr.specnorm(sc)
    opts = dict(scders=cfg.get("scders", False), uvder=cfg.get("uvder", False),
                vorgp=cfg.get("vordiv", False), divgp=cfg.get("vordiv", False))
    nproma = cfg.get("nproma", 0)
    for it in range(2):                                   # --niter 2
        gp = tr.inv_trans(vor, div, sc, nproma=nproma, **opts)
        iu = (nuv if opts["vorgp"] else 0) + (nuv if opts["divgp"] else 0)
        gin = np.ascontiguousarray(gp[:, iu:iu + 2 * nuv + nsc])
        vor, div, sc = tr.dir_trans(gin, nuv, nsc, nproma=nproma)
    eps = np.finfo(float).eps
    # criterion of ectrans-benchmark.F90:847-871 (--check 100): relative spectral-norm error <= 100 eps
    assert np.abs(tr.specnorm(vor) / n0v - 1).max() <= 100 * eps
    assert np.abs(tr.specnorm(div) / n0v - 1).max() <= 100 * eps
    assert np.abs(tr.specnorm(sc) / n0s - 1).max() <= 100 * eps
    tr.release()


def test_benchmark_driver_config0(built):
    """BASELINE config 0 through the ectrans-benchmark mirror: T79 / O80, 10 levels, 1 scalar field,
    inverse + direct, --norms --check 100 (the reference's own pass criterion)."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "tools", "ectrans_benchmark.py"), "-t", "79", "-g", "O80", "-l", "10",
           "-f", "1", "-n", "3", "--niter-warmup", "1", "--norms", "--check", "100"]
    for extra in ([], ["--vordiv", "--scders", "--uvders", "--nproma", "16"], ["--device-resident"]):
        out = subprocess.run(cmd + extra, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        assert "Inverse-direct transforms" in out.stdout and "Correctness test failed" not in out.stdout


def test_benchmark_checksums_reproducible(built, tmp_path):
    """--dump-checksums (ectrans-benchmark.F90:1455-1638 harness): host arrays and device-resident runs, blocked and
    unblocked, give identical per-field checksums at every iteration."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    base = [sys.executable, os.path.join(root, "tools", "ectrans_benchmark.py"), "-t", "47", "-g", "O48", "-l", "3",
            "-f", "2", "-n", "2", "--niter-warmup", "0"]
    dumps = []
    for i, extra in enumerate(([], ["--device-resident"], ["--nproma", "33"])):
        f = tmp_path / f"crc{i}.txt"
        out = subprocess.run(base + extra + ["--dump-checksums", str(f)], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        dumps.append(f.read_text())
    assert dumps[0].count("zgp (") == 2 * (2 * 3 + 3 * 2 + 1) and "zspscalar (7)" in dumps[0]
    assert dumps[0] == dumps[1] == dumps[2]


# ---------------------------------------------------------------------------------------------
# GPNORM_TRANS, VORDIV_TO_UV, TRANS_INQ(PRPNM) / TRANS_PNM
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nproma", [0, 333])
def test_gpnorm_trans(eb, nproma):
    """gpnorm_trans_ctl_mod.F90: weighted per-latitude sums added in latitude order; min / max exact."""
    T, N = 63, 64
    nloen = eb.octahedral_nloen(N)
    tr = eb.Transform(T, nloen)
    s = eo.setup(T, 2 * N, nloen, tables=False)
    rng = np.random.default_rng(11)
    flat = rng.normal(size=(5, tr.ngptot)) + np.arange(5)[:, None]
    ave, mn, mx = eo.gpnorm_trans(s, flat)
    gp = block(flat, nproma) if nproma else flat[None]
    a, lo, hi = tr.gpnorm_trans(gp, nproma=nproma)
    np.testing.assert_allclose(a, ave, rtol=1e-13, atol=1e-14)
    np.testing.assert_array_equal(lo, mn)
    np.testing.assert_array_equal(hi, mx)
    # a constant field: average = the constant (sum of weights = 1)
    a, lo, hi = tr.gpnorm_trans(np.full((1, 2, tr.ngptot), 3.25))
    np.testing.assert_allclose(a, 3.25, rtol=1e-14)
    assert (lo == 3.25).all() and (hi == 3.25).all()
    # LDAVE_ONLY: the extrema come from the caller
    a2, lo2, hi2 = tr.gpnorm_trans(gp, nproma=nproma, ave_only=True, pmin=mn - 1.0, pmax=mx + 2.0)
    np.testing.assert_allclose(a2, ave, rtol=1e-13, atol=1e-14)
    np.testing.assert_array_equal(lo2, mn - 1.0)
    np.testing.assert_array_equal(hi2, mx + 2.0)
    # device pointers
    import torch
    a3, lo3, hi3 = tr.gpnorm_trans(torch.from_numpy(np.ascontiguousarray(gp)).cuda(), nproma=nproma)
    np.testing.assert_array_equal(a3, a2)
    tr.release()


def test_vordiv_to_uv(eb):
    """vd2uv_mod.F90: against the oracle, with and without a resolution handle, and through the transform:
    with no vorticity / divergence at n = T the synthesis of (U, V) cos(theta) as scalars is u, v times cos(theta)."""
    T, N = 79, 80
    nloen = eb.octahedral_nloen(N)
    tr = eb.Transform(T, nloen)
    s = eo.setup(T, 2 * N, nloen)
    vor = eo.random_spectral(s, 3, 5, zero00=True); div = eo.random_spectral(s, 3, 6, zero00=True)
    for m in range(T + 1):      # drop n = T so that the discarded n = T+1 row of U, V is zero
        o = int(s.nasm0[m]) + 2 * (T - m)
        vor[:, o:o + 2] = 0.0; div[:, o:o + 2] = 0.0
    ur, vr = eo.vordiv_to_uv(s, vor, div)
    u, v = tr.vordiv_to_uv(T_(vor), T_(div))
    assert rel(u.T, ur) < TOL and rel(v.T, vr) < TOL
    u0, v0 = eb.vordiv_to_uv(T, T_(vor), T_(div))
    np.testing.assert_array_equal(u0, u); np.testing.assert_array_equal(v0, v)
    gp = tr.inv_trans(T_(vor), T_(div))[0]                  # u, v
    gs = tr.inv_trans(spscalar=np.ascontiguousarray(np.concatenate([u, v], axis=1)))[0]
    cos = np.repeat(np.sqrt(s.r1mu2), nloen)
    assert rel(gs, gp * cos) < 1e-11
    import torch
    ud, vd = tr.vordiv_to_uv(torch.from_numpy(T_(vor)).cuda(), torch.from_numpy(T_(div)).cuda())
    tr.synchronize()
    np.testing.assert_array_equal(ud.cpu().numpy(), u)
    assert eb.lib().ect_vordiv_to_uv(tr.handle, T + 1, 1, 1, 1, 1, 1, 0) == -4     # truncation differs from the handle's
    tr.release()


def test_legendre_polynomials_reference_layout(eb):
    """TRANS_INQ(PRPNM) (trans_inq.F90:444-464) and TRANS_PNM (trans_pnm.F90:127-177) layouts, values = the table
    of the transforms (bit for bit the oracle's SUPOLF restatement up to its last-digit differences)."""
    T, N = 47, 48
    nloen = eb.octahedral_nloen(N)
    tr = eb.Transform(T, nloen)
    s = eo.setup(T, 2 * N, nloen)
    ref = eo.rpnm_reference_layout(s)                       # (ndgnh, nspolegl)
    rpnm, npms = tr.legendre_polynomials()
    assert rpnm.shape == (ref.shape[1], ref.shape[0])
    np.testing.assert_array_equal(npms, np.concatenate([[0], np.cumsum([T + 2 - m for m in range(T)])]))
    assert np.abs(rpnm.T - ref).max() <= 2e-13 * max(np.abs(ref).max(), 1.0)      # FMA contraction on the device: a few ulp of the largest entry
    for m in (0, 1, 17, T):
        pm = tr.trans_pnm(m)                                # (T-m+3, ndgnh)
        np.testing.assert_array_equal(pm[:T + 2 - m], rpnm[npms[m]:npms[m] + T + 2 - m])
        assert not pm[T + 2 - m:].any()
    # ectrans4py.get_legendre_assets: KNMENG, PGW, PRPNM
    from ectrans_b200 import ectrans4py as e4
    nmeng, gw, prpnm = e4.get_legendre_assets(2 * N, T, 2 * N, ref.shape[1], nloen, 10)
    np.testing.assert_array_equal(nmeng, s.nmen)
    np.testing.assert_allclose(gw, s.rw, rtol=1e-14)
    np.testing.assert_array_equal(prpnm, rpnm.T)
    tr.release()


def test_legendre_cache_file_reference_format(eb, tmp_path):
    """CDIO_LEGPOL='writef' / 'readf': the file follows write_legpol_mod.F90:57-170 byte for byte (label, NSMAX,
    NDGNH, NLOEN/NMEN pairs, per m RPNMA then RPNMS column major with n descending, EOF label); a handle set up from
    the file transforms bit-identically; the reader's consistency checks (read_legpol_mod.F90:75-103) fire."""
    T, N = 63, 64
    nloen = eb.octahedral_nloen(N)
    path = str(tmp_path / "legpol.bin")
    tr = eb.Transform(T, nloen, legpol_write=path)
    s = eo.setup(T, 2 * N, nloen)
    raw = open(path, "rb").read()
    assert raw[:8] == b"LEGPOL  " and raw[-16:] == b"LEGPOL---EOF-EOF"
    hdr = np.frombuffer(raw, dtype=np.int32, count=2, offset=8)
    assert tuple(hdr) == (T, N)
    geo = np.frombuffer(raw, dtype=np.int32, count=2 * N, offset=16).reshape(N, 2)
    np.testing.assert_array_equal(geo[:, 0], nloen[:N]); np.testing.assert_array_equal(geo[:, 1], s.nmen[:N])
    off = 16 + 8 * N
    for m in range(T + 1):
        nd, ila, ils = int(s.ndglu[m]), (T - m + 2) // 2, (T - m + 3) // 2
        a = np.frombuffer(raw, dtype=np.float64, count=nd * ila, offset=off).reshape(ila, nd); off += 8 * nd * ila
        b = np.frombuffer(raw, dtype=np.float64, count=nd * ils, offset=off).reshape(ils, nd); off += 8 * nd * ils
        # column J (1-based) <-> n = m + 2 (ILA - J) + 1 resp. m + 2 (ILS - J); the oracle tables are n ascending
        assert a.size == 0 or np.abs(a[::-1].T - s.pa[m]).max() <= 2e-13 * max(np.abs(s.pa[m]).max(), 1.0)
        assert np.abs(b[::-1].T - s.ps[m]).max() <= 2e-13 * max(np.abs(s.ps[m]).max(), 1.0)
    assert off == len(raw) - 16
    sp = T_(eo.random_spectral(s, 3, 4))
    g0 = tr.inv_trans(spscalar=sp)
    tr2 = eb.Transform(T, nloen, legpol_read=path)
    np.testing.assert_array_equal(tr2.inv_trans(spscalar=sp), g0)
    np.testing.assert_array_equal(tr2.dir_trans(g0, 0, 3)[2], tr.dir_trans(g0, 0, 3)[2])
    tr2.release(); tr.release()
    with pytest.raises(eb.EctError, match="WRONG SPECTRAL TRUNCATION"):
        eb.Transform(T - 1, nloen, legpol_read=path)
    with pytest.raises(eb.EctError, match="WRONG NLOEN"):
        nl2 = nloen.copy(); nl2[0] += 4; nl2[-1] += 4
        eb.Transform(T, nl2, legpol_read=path)
    bad = str(tmp_path / "bad.bin"); open(bad, "wb").write(b"LEGPOLBF" + raw[8:])
    with pytest.raises(eb.EctError, match="WRONG LABEL"):
        eb.Transform(T, nloen, legpol_read=bad)


def test_gpnorm_vordiv_single_precision(eb):
    """sp handles: float arrays at the boundary of GPNORM_TRANS / VORDIV_TO_UV (tolerance 1e-5, north_star)."""
    T, N = 63, 64
    nloen = eb.octahedral_nloen(N)
    tr = eb.Transform(T, nloen, precision="sp")
    s = eo.setup(T, 2 * N, nloen, tables=False)
    rng = np.random.default_rng(5)
    flat = (rng.normal(size=(3, tr.ngptot)) + 2.0).astype(np.float32)
    ave, mn, mx = eo.gpnorm_trans(s, flat.astype(np.float64))
    a, lo, hi = tr.gpnorm_trans(flat[None])
    np.testing.assert_allclose(a, ave, rtol=1e-12)          # sums run in fp64 on the float data
    np.testing.assert_array_equal(lo, mn); np.testing.assert_array_equal(hi, mx)
    f32 = lambda x: x.astype(np.float32).astype(np.float64)
    vor = f32(eo.random_spectral(s, 2, 1, zero00=True)); div = f32(eo.random_spectral(s, 2, 2, zero00=True))
    ur, vr = eo.vordiv_to_uv(s, vor, div)
    u, v = tr.vordiv_to_uv(T_(vor).astype(np.float32), T_(div).astype(np.float32))
    assert u.dtype == np.float32 and rel(u.T, ur) < 1e-6 and rel(v.T, vr) < 1e-6
    tr.release()


# ---- chirp-z rows on CTA pairs (csrc/fourier_cz.h, k_fourier_cz) ----
@pytest.fixture
def cz_all(monkeypatch):
    """Route every chirp-z row through the pair kernel (by default only rows whose undivided work array does not fit)."""
    monkeypatch.setenv("ECT_FFT_CZ", "1")


@pytest.mark.parametrize("T,N,prec,tol", [(159, 160, "dp", 1e-12), (399, 400, "dp", 1e-12), (399, 400, "sp", 5e-6)])
def test_cz_pair_kernel_against_oracle(eb, cz_all, T, N, prec, tol):
    nloen = eb.octahedral_nloen(N)
    tr = eb.Transform(T, nloen, precision=prec)
    s = eo.setup(T, 2 * N, nloen)
    nuv, nsc = 3, 4
    dt = np.float64 if prec == "dp" else np.float32
    f = (lambda a: a.astype(np.float32).astype(np.float64)) if prec == "sp" else (lambda a: a)
    vor = f(eo.random_spectral(s, nuv, 1, zero00=True)); div = f(eo.random_spectral(s, nuv, 2, zero00=True))
    sc = f(eo.random_spectral(s, nsc, 3))
    opts = dict(scders=True, uvder=True)
    ref = eo.inv_trans(s, vor, div, sc, **opts)
    gp = tr.inv_trans(T_(vor).astype(dt), T_(div).astype(dt), T_(sc).astype(dt), nproma=1001, **opts)
    got = unblock(gp, tr.ngptot).astype(np.float64)
    for i in range(ref.shape[0]):
        assert rel(got[i], ref[i]) < tol, i
    nf = 2 * nuv + nsc
    gin = f(ref[:nf])
    rv, rd, rs = eo.dir_trans(s, gin, nuv, nsc)
    ov, od, os_ = tr.dir_trans(gin[None].astype(dt), nuv, nsc)
    for a, b in ((ov, rv), (od, rd), (os_, rs)):
        assert rel(a.T.astype(np.float64), b) < tol
    tr.release()


def test_cz_pair_kernel_odd_rows(eb, cz_all):
    """Classic reduced grid (odd row lengths 27, 45, 75) through the pair kernel."""
    nloen = np.asarray(N32_CLASSIC + N32_CLASSIC[::-1], dtype=np.int32)
    T = 42
    tr = eb.Transform(T, nloen)
    s = eo.setup(T, nloen.size, nloen)
    sc = eo.random_spectral(s, 5, 3)
    ref = eo.inv_trans(s, None, None, sc)
    gp = tr.inv_trans(spscalar=T_(sc))
    assert rel(gp[0], ref) < 1e-12
    back = tr.dir_trans(gp, 0, 5)[2]
    rs = eo.dir_trans(s, ref, 0, 5)[2]
    assert rel(back.T, rs) < 1e-12
    tr.release()


def test_tco2559_rows_dp(eb):
    """Rows longer than ~5400 points in double precision (TCo2559 / O2560: up to 10256 points, convolution length
    16384): the undivided chirp-z work array does not fit an SM, the pair kernel's halves (139 KB) do.  Round 1 returned
    ECT_ERR_NOTIMPL here.  Reduced field count; analytic harmonic, linearity and round trip at the dp tolerance."""
    T, N = 2559, 2560
    nloen = eb.octahedral_nloen(N)
    tr = eb.Transform(T, nloen)
    nf = 2
    rng = np.random.default_rng(12)
    n = np.concatenate([np.repeat(np.arange(m, T + 1), 2) for m in range(T + 1)]).astype(float)
    a = rng.uniform(-1, 1, (tr.nspec2, nf)) / (1 + n[:, None]) ** 2
    a[1:2 * (T + 1):2] = 0
    b = np.zeros_like(a)
    b[int(tr.nasm0[4]) + 2 * (19 - 4)] = 1.0
    ga, gb = tr.inv_trans(spscalar=a), tr.inv_trans(spscalar=b)
    off = np.concatenate([[0], np.cumsum(nloen)])
    for j in (0, 5, 1000, 2047, 2559, 2560, 4000, 5119):
        nlon = int(nloen[j])
        row = gb[0, 0, off[j]:off[j] + nlon]
        if tr.nmen[j] >= 4:
            p = eo.supolf(4, 19, tr.rmu[j])[19, 0]
            assert np.abs(row - 2 * p * np.cos(2 * np.pi * 4 * np.arange(nlon) / nlon)).max() < 1e-12
    gab = tr.inv_trans(spscalar=a - 2.5 * b)
    assert rel(gab, ga - 2.5 * gb) < 1e-13
    back = tr.dir_trans(ga, 0, nf)[2]
    assert rel(back, a) < 2e-11
    assert np.abs(tr.specnorm(back) / tr.specnorm(a) - 1).max() < 1e-12
    tr.release()


# ---- sp Legendre contraction on tcgen05 (csrc/legendre_tc.cu) ----
@pytest.mark.parametrize("T,N,nuv,nsc", [(47, 48, 2, 3), (159, 160, 5, 70), (399, 400, 20, 21)])
def test_sp_tcgen05_contraction(eb, monkeypatch, T, N, nuv, nsc):
    """sp handles: m > 0 as 3xTF32 on the tensor cores (TMA-staged MN-major operands, TMEM accumulators), m = 0 on the
    FP64 kernel.  Against the oracle at the north_star sp tolerance (1e-5), and against the all-FP64 contraction of round 1
    (ECT_SP_TC=0), which it must stay close to."""
    nloen = eb.octahedral_nloen(N)
    s = eo.setup(T, 2 * N, nloen)
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    vor, div, sc = f32(eo.random_spectral(s, nuv, 1, zero00=True)), f32(eo.random_spectral(s, nuv, 2, zero00=True)), f32(eo.random_spectral(s, nsc, 3))
    ref = eo.inv_trans(s, vor, div, sc, scders=True)
    nf = 2 * nuv + nsc
    rv, rd, rs = eo.dir_trans(s, f32(ref[:nf]), nuv, nsc)
    S_ = lambda a: np.ascontiguousarray(a.T).astype(np.float32)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("ECT_SP_TC", mode)
        tr = eb.Transform(T, nloen, precision="sp")
        gp = tr.inv_trans(S_(vor), S_(div), S_(sc), scders=True)
        sp3 = tr.dir_trans(f32(ref[:nf])[None].astype(np.float32), nuv, nsc)
        out[mode] = (gp[0].astype(np.float64), [a.T.astype(np.float64) for a in sp3], tr.timings()["launches"])
        tr.release()
    for i in range(ref.shape[0]):
        assert rel(out["1"][0][i], ref[i]) < 1e-5
        assert rel(out["1"][0][i], out["0"][0][i]) < 1e-5
    for a, b in zip(out["1"][1], (rv, rd, rs)):
        assert rel(a, b) < 1e-5
    assert out["1"][2] > out["0"][2]          # the tensor-core path really ran (its split / contraction launches)


@pytest.mark.parametrize("T,N,prec,tol", [(159, 160, "dp", 1e-12), (399, 400, "dp", 1e-12), (399, 400, "sp", 5e-6)])
def test_cz2_two_group_kernel_against_oracle(eb, monkeypatch, T, N, prec, tol):
    """Chirp-z rows with both halves of the split convolution in one CTA (two warp groups, k_fourier_cz2)."""
    monkeypatch.setenv("ECT_FFT_CZ", "2")
    nloen = eb.octahedral_nloen(N)
    tr = eb.Transform(T, nloen, precision=prec)
    s = eo.setup(T, 2 * N, nloen)
    nuv, nsc = 3, 4
    dt = np.float64 if prec == "dp" else np.float32
    f = (lambda a: a.astype(np.float32).astype(np.float64)) if prec == "sp" else (lambda a: a)
    vor = f(eo.random_spectral(s, nuv, 1, zero00=True)); div = f(eo.random_spectral(s, nuv, 2, zero00=True))
    sc = f(eo.random_spectral(s, nsc, 3))
    opts = dict(scders=True, uvder=True)
    ref = eo.inv_trans(s, vor, div, sc, **opts)
    gp = tr.inv_trans(T_(vor).astype(dt), T_(div).astype(dt), T_(sc).astype(dt), nproma=1001, **opts)
    got = unblock(gp, tr.ngptot).astype(np.float64)
    for i in range(ref.shape[0]):
        assert rel(got[i], ref[i]) < tol, i
    nf = 2 * nuv + nsc
    gin = f(ref[:nf])
    rv, rd, rs = eo.dir_trans(s, gin, nuv, nsc)
    ov, od, os_ = tr.dir_trans(gin[None].astype(dt), nuv, nsc)
    for a, b in ((ov, rv), (od, rd), (os_, rs)):
        assert rel(a.T.astype(np.float64), b) < tol
    tr.release()


# ---- direct Fourier stage: records collected in a local slot and pushed as long runs (peer mode; forced here) ----
@pytest.mark.parametrize("T,N,prec,nuv,nsc", [(159, 160, "dp", 3, 4), (399, 400, "dp", 9, 23), (399, 400, "sp", 5, 14),
                                              (95, 96, "dp", 0, 1)])
def test_direct_push_slots_bit_identical(eb, monkeypatch, T, N, prec, nuv, nsc):
    """ECT_FFT_PUSH=2 runs the slot + push store path of k_fourier<direct> on one rank (the consumer's buffer is then the
    local one): same arithmetic, so the spectra must equal the direct-store path bit for bit -- whole chunks of 16
    fields, a ragged last chunk, single-field pairs at group ends -- and match the oracle."""
    nloen = eb.octahedral_nloen(N)
    s = eo.setup(T, 2 * N, nloen)
    dt = np.float64 if prec == "dp" else np.float32
    rng = np.random.default_rng(7)
    nf = 2 * nuv + nsc
    gin = rng.standard_normal((nf, s.ngptot)).astype(dt)
    out = {}
    for mode in ("0", "2"):
        monkeypatch.setenv("ECT_FFT_PUSH", mode)
        tr = eb.Transform(T, nloen, precision=prec)
        for rep in range(2):                       # twice: slots are released and claimed again
            res = tr.dir_trans(gin[None], nuv, nsc)
        import ctypes, torch
        info = (ctypes.c_longlong * 4)()
        assert eb.lib().ect_debug_push_info(ctypes.c_int(tr.handle), info) == 0
        assert bool(info[0]) == (mode == "2")
        if mode == "2":
            nsm = torch.cuda.get_device_properties(0).multi_processor_count
            assert info[1] >= nsm and info[2] > 0 and 2 <= info[3] <= 32, list(info)
            print("push info (enabled, sm id range, scratch bytes, slots per SM):", list(info))
        out[mode] = [np.array(x) for x in res if x is not None]
        tr.release()
    for a, b in zip(out["0"], out["2"]):
        assert a.shape == b.shape and np.array_equal(a, b)
    if T <= 159:
        ref = eo.dir_trans(s, gin.astype(np.float64), nuv, nsc)
        got = out["2"]
        refs = [r for r in ref if r is not None and r.size]
        for a, b in zip(got, refs):
            assert rel(a.T.astype(np.float64), b) < (1e-12 if prec == "dp" else 5e-6)


def test_winds_and_derivatives_known_answers_gpu(eb):
    """The CUDA path against the closed forms of tests/test_oracle_golden.py::test_winds_and_derivatives_known_answers
    (solid-body rotation, purely divergent flow, N-S / E-W derivatives of Pbar_1^0 and Pbar_1^1): pins the wind and
    derivative conventions independently of the oracle."""
    T, N = 79, 80
    nloen = eb.octahedral_nloen(N)
    s = eo.setup(T, 2 * N, nloen, tables=False)
    tr = eb.Transform(T, nloen)
    ra, c = 6371229.0, 3.0e-5

    def one(m, n):
        sp = np.zeros((1, s.nspec2))
        sp[0, int(s.nasm0[m]) + 2 * (n - m)] = c
        return T_(sp)

    zero = T_(np.zeros((1, s.nspec2)))
    lat_of = np.repeat(np.arange(s.ndgl), s.nloen)
    cost = np.sqrt(1.0 - s.rmu ** 2)[lat_of]
    mu = s.rmu[lat_of]
    lam = np.concatenate([2 * np.pi * np.arange(n) / n for n in s.nloen])
    omega = c * np.sqrt(3.0) / 2
    tol = 1e-11
    u, v = unblock(tr.inv_trans(one(0, 1), zero), tr.ngptot)[:2]
    assert np.abs(u - omega * ra * cost).max() < tol * omega * ra and np.abs(v).max() < tol * omega * ra
    u, v = unblock(tr.inv_trans(zero, one(0, 1)), tr.ngptot)[:2]
    assert np.abs(v + 0.5 * ra * c * np.sqrt(3.0) * cost).max() < tol * omega * ra and np.abs(u).max() < tol * omega * ra
    f, ns, ew = unblock(tr.inv_trans(None, None, one(0, 1), scders=True), tr.ngptot)
    assert np.abs(f - c * np.sqrt(3.0) * mu).max() < tol * c
    assert np.abs(ns - c * np.sqrt(3.0) * cost / ra).max() < tol * c / ra and np.abs(ew).max() < tol * c / ra
    f, ns, ew = unblock(tr.inv_trans(None, None, one(1, 1), scders=True), tr.ngptot)
    k = 2 * c * np.sqrt(1.5)
    assert np.abs(f - k * cost * np.cos(lam)).max() < tol * k
    assert np.abs(ns + k * mu * np.cos(lam) / ra).max() < tol * k / ra
    assert np.abs(ew + k * np.sin(lam) / ra).max() < tol * k / ra
    tr.release()
