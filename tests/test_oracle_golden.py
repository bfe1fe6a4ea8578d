"""The oracle against the reference's own golden vectors and known answers.
Reference tests mirrored: tests/test_ectrans4py/test_ectrans4py.py:94-158,
tests/transi/transi_test_program.c:76-81,150-164, src/programs/ectrans-benchmark.F90:1389-1415,847-871."""
import numpy as np
import pytest

import ectrans_oracle as eo

EPSILON = 1e-10   # tests/test_ectrans4py/test_ectrans4py.py:16


@pytest.fixture(scope="module")
def s148(golden):
    return eo.setup(148, 150, golden["nloen"])


def _pack(golden):
    nl = golden["nloen"]
    return np.concatenate([golden["gp_latlon"][i, :nl[i]] for i in range(150)])


def test_sizes_and_nmen(s148, golden):
    # test_trans_inq4py: (33052, 11175) and the zonal wavenumber array
    assert (s148.ngptot, s148.nspec2 // 2) == (33052, 11175)
    np.testing.assert_array_equal(s148.nmen, golden["nmen"])


def test_weights_sum(s148):
    # test_get_legendre_assets
    assert abs(s148.rw.sum() - 1.0) < EPSILON


def test_sp2gp_golden(s148, golden):
    gp = eo.inv_trans(s148, spscalar=golden["sp"][None, :])[0]
    d = gp - _pack(golden)
    assert abs(d.max()) < EPSILON and abs(d.min()) < EPSILON


def test_gp2sp_golden(s148, golden):
    _, _, sp = eo.dir_trans(s148, _pack(golden)[None, :], 0, 1)
    d = sp[0] - golden["sp"]
    assert abs(d.max()) < EPSILON and abs(d.min()) < EPSILON


def test_gauss_nodes_against_numpy():
    for n in (48, 160, 320):
        mu, w = eo.gauss_latitudes(n)
        x, ww = np.polynomial.legendre.leggauss(n)
        assert np.abs(mu - x[::-1]).max() < 1e-14
        assert np.abs(w - ww[::-1] / 2).max() < 1e-14


@pytest.fixture(scope="module")
def s79():
    return eo.setup(79, 160, eo.octahedral_nloen(80))


def test_constant_field(s79):
    # transi_test_program.c: constant grid-point field c => only psi(0,0) = c
    gp = np.stack([np.full(s79.ngptot, c) for c in (1.0, 2.0, 3.0, 4.0)])
    _, _, sp = eo.dir_trans(s79, gp, 0, 4)
    for i, c in enumerate((1.0, 2.0, 3.0, 4.0)):
        assert abs(sp[i, 0] - c) < 1e-13
        assert np.abs(sp[i, 1:]).max() < 1e-13


def test_benchmark_harmonic(s79):
    # ectrans-benchmark input Re psi(4,19) = 1  =>  gp = 2 P_19^4(mu) cos(4 lambda)
    sp = eo.benchmark_spectral(s79, 1)
    gp = eo.inv_trans(s79, spscalar=sp)[0]
    for j in (0, 17, 79, 80, 159):
        nlon = int(s79.nloen[j])
        p = eo.supolf(4, 19, s79.rmu[j])[19, 0]
        lam = 2 * np.pi * np.arange(nlon) / nlon
        row = gp[s79.latoff[j]:s79.latoff[j] + nlon]
        if s79.nmen[j] >= 4:
            assert np.abs(row - 2 * p * np.cos(4 * lam)).max() < 1e-13
        else:
            assert np.abs(row).max() == 0.0
    # round trip: relative spectral-norm error <= 100 eps (ectrans-benchmark.F90:847-871)
    _, _, back = eo.dir_trans(s79, gp[None, :], 0, 1)
    n0, n1 = eo.specnorm(s79, sp)[0], eo.specnorm(s79, back)[0]
    assert abs(n1 - n0) / n0 <= 100 * np.finfo(float).eps


def test_vordiv_roundtrip_and_vorgp(s79):
    vor = eo.random_spectral(s79, 2, 1, zero00=True, decay=True)
    div = eo.random_spectral(s79, 2, 2, zero00=True, decay=True)
    gp = eo.inv_trans(s79, vor, div, vorgp=True, divgp=True)
    # grid-point vorticity equals the scalar transform of the vorticity coefficients
    assert np.abs(gp[0:2] - eo.inv_trans(s79, spscalar=vor)).max() < 1e-12
    v2, d2, _ = eo.dir_trans(s79, gp[4:8], 2, 0)
    for a, b in ((vor, v2), (div, d2)):
        assert np.abs(eo.specnorm(s79, a) - eo.specnorm(s79, b)).max() / eo.specnorm(s79, a).max() < 1e-10


def test_ew_derivative_is_im(s79):
    # E-W derivative of a scalar = i m / (a cos theta) F_m  (fsc_mod.F90:163-187): check against finite spectrum
    sc = eo.random_spectral(s79, 1, 5, decay=True)
    gp = eo.inv_trans(s79, spscalar=sc, scders=True)
    j = 80
    nlon = int(s79.nloen[j])
    row = gp[0, s79.latoff[j]:s79.latoff[j] + nlon]
    f = np.fft.rfft(row)
    d = np.fft.irfft(1j * np.arange(f.size) * f, n=nlon) * s79.racthe[j]
    assert np.abs(d - gp[2, s79.latoff[j]:s79.latoff[j] + nlon]).max() < 1e-12 * np.abs(d).max() + 1e-18


def test_adjoint_identity():
    # tests/trans/test_invtrans_adjoint.F90:180-232: <F x, y>_w = <x, F* y>; on a full Gaussian grid the
    # weighted direct transform is the adjoint of the inverse: <inv(x), y>_gridweights = <x, dir(y)>_spectral
    T, ndgl = 21, 32
    nloen = np.full(ndgl, 64)
    s = eo.setup(T, ndgl, nloen)
    x = eo.random_spectral(s, 1, 11)
    rng = np.random.default_rng(3)
    y = rng.uniform(-1, 1, size=(1, s.ngptot))
    gx = eo.inv_trans(s, spscalar=x)
    _, _, sy = eo.dir_trans(s, y, 0, 1)
    w = np.repeat(s.rw / nloen, nloen)
    lhs = float(np.sum(gx[0] * y[0] * w))
    # spectral inner product with the m>0 double weight (spnormd_mod.F90:36-51)
    wt = np.full(s.nspec2, 2.0)
    wt[:2 * (T + 1)] = 1.0
    rhs = float(np.sum(x[0] * sy[0] * wt))
    assert abs(lhs - rhs) <= 20000 * np.finfo(float).eps * max(abs(lhs), 1.0)


def test_oracle_adjoints_satisfy_reference_identity():
    """The adjoint identity of tests/trans/test_invtrans_adjoint.F90:192-222 and test_dirtrans_adjoint.F90 with the
    reference's inner products, tolerance 20000 eps, for the oracle's matrix-transpose adjoints."""
    T, N = 10, 12
    nloen = eo.octahedral_nloen(N)
    s = eo.setup(T, 2 * N, nloen)
    rng = np.random.default_rng(1)
    nuv, nsc = 2, 2
    vor = eo.random_spectral(s, nuv, 1, zero00=True); div = eo.random_spectral(s, nuv, 2, zero00=True)
    sc = eo.random_spectral(s, nsc, 3)
    w = eo.spectral_weights(s)
    y = rng.uniform(-1, 1, (2 * nuv + nsc, s.ngptot))
    fx = eo.inv_trans(s, vor, div, sc)
    va, da, sa = eo.inv_transad(s, y, nuv, nsc)
    lhs = np.sum(fx * y)
    rhs = np.sum(w * vor * va) + np.sum(w * div * da) + np.sum(w * sc * sa)
    assert abs(lhs - rhs) <= 20000 * np.finfo(float).eps * abs(lhs)
    dv, dd, ds = eo.dir_trans(s, y, nuv, nsc)
    ga = eo.dir_transad(s, vor, div, sc)
    lhs = np.sum(w * dv * vor) + np.sum(w * dd * div) + np.sum(w * ds * sc)
    rhs = np.sum(y * ga)
    assert abs(lhs - rhs) <= 20000 * np.finfo(float).eps * abs(lhs)


def test_oracle_gpnorm_vordiv_rpnm_known_answers():
    """Known answers for the restatements of GPNORM_TRANS, VORDIV_TO_UV and the TRANS_INQ(PRPNM) layout: a constant
    field averages to itself (sum of Gaussian weights = 1, test_ectrans4py.py:119-121); (U, V) cos(theta) synthesised
    as scalars equal u, v cos(theta) of the vor/div inverse transform when nothing sits at n = T; column NPMS(m) + p
    of PRPNM is n = T + 2 - p (trans_inq.F90:450-462), e.g. the last column of the m = 0 block is P_0^0 = 1."""
    T, N = 31, 32
    nloen = eo.octahedral_nloen(N)
    s = eo.setup(T, 2 * N, nloen)
    ave, mn, mx = eo.gpnorm_trans(s, np.full((2, s.ngptot), 1.75))
    assert np.allclose(ave, 1.75, rtol=1e-14) and (mn == 1.75).all() and (mx == 1.75).all()
    vor = eo.random_spectral(s, 2, 1, zero00=True); div = eo.random_spectral(s, 2, 2, zero00=True)
    for m in range(T + 1):
        o = int(s.nasm0[m]) + 2 * (T - m)
        vor[:, o:o + 2] = 0.0; div[:, o:o + 2] = 0.0
    u, v = eo.vordiv_to_uv(s, vor, div)
    guv = eo.inv_trans(s, vor, div, None)
    gs = eo.inv_trans(s, None, None, np.concatenate([u, v]))
    cos = np.repeat(np.sqrt(s.r1mu2), nloen)
    assert np.linalg.norm(gs - guv * cos) / np.linalg.norm(gs) < 1e-12
    r = eo.rpnm_reference_layout(s)
    assert r.shape == (N, sum(T + 2 - m for m in range(T + 1)))
    assert np.allclose(r[:, T + 1], 1.0, rtol=1e-13)                  # m = 0, p = T + 2: n = 0
    assert np.allclose(r[:, T], np.sqrt(3.0) * s.rmu[:N], rtol=1e-13)  # n = 1: sqrt(3) mu


def _one_coef(s, m, n, val):
    sp = np.zeros((1, s.nspec2))
    sp[0, int(s.nasm0[m]) + 2 * (n - m)] = val
    return sp


def test_winds_and_derivatives_known_answers(s79):
    """Closed-form answers that pin the conventions no golden array of the reference covers (vor/div -> u, v and
    the derivatives): with Pbar_1^0 = sqrt(3) mu and Pbar_1^1 = sqrt(3/2) cos(theta) (supolf_mod.F90 normalisation),
      vorticity   c Pbar_1^0 = 2 Omega mu   ->  solid-body rotation u = Omega a cos(theta), v = 0   (vdtuv_mod.F90:121-139,
                                                 fsc_mod.F90:138-160);
      divergence  c Pbar_1^0                ->  chi = -a^2 D / 2, v = (1/a) d chi / d theta = -(a/2) c sqrt(3) cos(theta), u = 0;
      scalar      c Pbar_1^0                ->  N-S derivative (1/a) d f / d theta = c sqrt(3) cos(theta) / a  (spnsde_mod.F90:95-114);
      scalar      c Pbar_1^1 e^{i lambda}   ->  f = 2 c sqrt(3/2) cos(theta) cos(lambda), N-S derivative -2 c sqrt(3/2) mu cos(lambda) / a,
                                                 E-W derivative (1/(a cos theta)) d f / d lambda = -2 c sqrt(3/2) sin(lambda) / a."""
    s = s79
    ra, c = 6371229.0, 3.0e-5
    zero = np.zeros((1, s.nspec2))
    lat_of = np.repeat(np.arange(s.ndgl), s.nloen)
    cost = np.sqrt(1.0 - s.rmu ** 2)[lat_of]
    mu = s.rmu[lat_of]
    lam = np.concatenate([2 * np.pi * np.arange(n) / n for n in s.nloen])
    omega = c * np.sqrt(3.0) / 2
    u, v = eo.inv_trans(s, _one_coef(s, 0, 1, c), zero)[:2]
    assert np.abs(u - omega * ra * cost).max() < 1e-11 * omega * ra and np.abs(v).max() < 1e-11 * omega * ra
    u, v = eo.inv_trans(s, zero, _one_coef(s, 0, 1, c))[:2]
    assert np.abs(v + 0.5 * ra * c * np.sqrt(3.0) * cost).max() < 1e-11 * omega * ra and np.abs(u).max() < 1e-11 * omega * ra
    f, ns, ew = eo.inv_trans(s, spscalar=_one_coef(s, 0, 1, c), scders=True)
    assert np.abs(f - c * np.sqrt(3.0) * mu).max() < 1e-11 * c
    assert np.abs(ns - c * np.sqrt(3.0) * cost / ra).max() < 1e-11 * c / ra and np.abs(ew).max() < 1e-11 * c / ra
    f, ns, ew = eo.inv_trans(s, spscalar=_one_coef(s, 1, 1, c), scders=True)
    k = 2 * c * np.sqrt(1.5)
    assert np.abs(f - k * cost * np.cos(lam)).max() < 1e-11 * k
    assert np.abs(ns + k * mu * np.cos(lam) / ra).max() < 1e-11 * k / ra
    assert np.abs(ew + k * np.sin(lam) / ra).max() < 1e-11 * k / ra
