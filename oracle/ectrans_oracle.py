"""CPU restatement (NumPy) of the ecTrans global spectral transform hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product path (``ectrans_b200``) never
routes through this module and has no CPU fallback.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks this module
against the reference's own golden vectors
(``tests/test_ectrans4py/data/tl149-c24-s1t@sp{,2gp}.npy``, abs tol 1e-10,
``tests/test_ectrans4py/test_ectrans4py.py:16,133-158``), the NMEN array
(``zonal_wavenumbers.npy``), the (33052, 11175) sizes and sum(weights) = 1.

The reference cannot be compiled in this image (no Fortran compiler, no fiat,
no FFTW, no BLAS -- SURVEY.md 8(c)), so the third-party arithmetic is restated:
FFTW's unnormalised c2r / r2c  ->  numpy.fft (pocketfft), BLAS xGEMM -> numpy
matmul (OpenBLAS).  Both are mathematically fixed operations, results differ
by rounding order only.

Every function cites the reference file:line it follows (paths relative to
``/root/reference/src/trans``).  Nothing is read from ``/root/reference`` at
run time.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

RA = 6371229.0  # common/external/setup_trans0.F90:129


# --------------------------------------------------------------------------
#  Gaussian latitudes and weights
# --------------------------------------------------------------------------
def _zfn_row(n: int) -> np.ndarray:
    """Fourier (cosine-series) coefficients of the normalised ordinary Legendre
    polynomial of degree n.  cpu/internal/suleg_mod.F90:249-263 (only row KDGL
    is consumed by SUGAW's LLOLD branch, sugaw_mod.F90:160-164)."""
    zfn = np.zeros(n + 1)
    zfnn = 2.0
    for jgl in range(1, n + 1):
        zfnn = zfnn * math.sqrt(1.0 - 0.25 / float(jgl * jgl))
    iodd = n % 2
    zfn[n] = zfnn
    for jgl in range(2, n - iodd + 1, 2):
        zfn[n - jgl] = zfn[n - jgl + 2] * float((jgl - 1) * (2 * n - jgl + 2)) / float(jgl * (2 * n - jgl + 1))
    return zfn


def gauss_latitudes(ndgl: int):
    """Gaussian latitudes mu (north->south) and weights (sum = 1).

    SUGAW LLOLD branch common/internal/sugaw_mod.F90:157-190, Newton iteration
    GAWL common/internal/gawl_mod.F90:91-111, one step CPLEDN
    common/internal/cpledn_mod.F90:94-129.
    """
    kn = ndgl
    row = _zfn_row(ndgl)
    iodd = ndgl % 2
    zfn = row[iodd::2].copy()  # ZFN(IK)=PFN(KDGL,JGL), JGL=IODD,KDGL,2
    ins2 = ndgl // 2
    jn = np.arange(2 - iodd, kn + 1, 2, dtype=np.float64)
    coef = zfn[1:1 + jn.size]
    eps = np.finfo(np.float64).eps
    mu = np.zeros(ndgl)
    w = np.zeros(ndgl)
    for jgl in range(1, ins2 + 1):
        z = float(4 * jgl - 1) * math.pi / float(4 * kn + 2)
        zx = z + 1.0 / (math.tan(z) * float(8 * kn * kn))
        iflag = 0
        zxn = zx
        zw = 0.0
        for _ in range(21):  # ITEMAX+1
            if iflag == 0:
                zdlk = 0.5 * zfn[0] if iodd == 0 else 0.0
                zdlk = zdlk + float(np.dot(coef, np.cos(jn * zx)))
                zdlldn = -float(np.dot(coef * jn, np.sin(jn * zx)))
                zmod = -zdlk / zdlldn
                zxn = zx + zmod
                zx = zxn
                if abs(zmod) <= eps * 1000.0:
                    iflag = 1
            else:
                zdlldn = -float(np.dot(coef * jn, np.sin(jn * zx)))
                zw = float(2 * kn + 1) / zdlldn ** 2
                break
        mu[jgl - 1] = math.cos(zxn)
        w[jgl - 1] = zw
    for jgl in range(ins2):
        mu[ndgl - 1 - jgl] = -mu[jgl]
        w[ndgl - 1 - jgl] = w[jgl]
    return mu, w


# --------------------------------------------------------------------------
#  Associated Legendre polynomials (SUPOLF)
# --------------------------------------------------------------------------
def supolf(km: int, knsmax: int, mu, kcheap: int = 1) -> np.ndarray:
    """Normalised associated Legendre polynomials P_n^m(mu), n = 0..knsmax,
    vectorised over an array of mu.  common/internal/supolf_mod.F90:85-247
    with the INI_POL constants of common/internal/tpm_pol.F90:74-81.

    kcheap: 1 all n, 2 only n-m even, 3 only n-m odd (entries of the other
    parity beyond n = m+3 are left at 0, as in the reference they are unset).
    Returns array [knsmax+1, nlat].
    """
    x = np.atleast_1d(np.asarray(mu, dtype=np.float64)).copy()
    nlat = x.size
    pol = np.zeros((knsmax + 1, nlat))
    eps = np.finfo(np.float64).eps
    cos2 = 1.0 - x * x
    cost = np.sqrt(cos2)
    polar = np.abs(cost) <= eps
    x = np.where(polar, 1.0, x)
    cost = np.where(polar, 0.0, cost)
    cos2 = np.where(polar, 0.0, cos2)
    with np.errstate(divide="ignore"):
        cost_r = np.where(polar, 0.0, 1.0 / np.where(polar, 1.0, cost))

    def dfa(n):
        return 1.0 / math.sqrt(float(n * (n + 1)))

    def dfb(n):
        return math.sqrt(float(2 * n + 1) / float(n * (n + 1)))

    if km == 0:
        dlkm2 = np.ones(nlat)
        dlkm1 = x.copy()
        pol[0] = dlkm2
        if knsmax >= 1:
            pol[1] = dlkm1 * dfb(1) / dfa(1)
        for n in range(2, knsmax + 1):
            dlk = (float(2 * n - 1) / float(n)) * x * dlkm1 - (float(n - 1) / float(n)) * dlkm2
            pol[n] = dlk * dfb(n) / dfa(n)
            dlkm2 = dlkm1
            dlkm1 = dlk
        return pol
    if km == 1:
        dlkm2 = np.ones(nlat)
        dlkm1 = x.copy()
        pol[0] = 0.0
        if knsmax >= 1:
            pol[1] = cost * dfb(1)
        for n in range(2, knsmax + 1):
            dlk = (float(2 * n - 1) / float(n)) * x * dlkm1 - (float(n - 1) / float(n)) * dlkm2
            dl1 = float(n) * (dlkm1 - x * dlk) * cost_r
            pol[n] = dl1 * dfb(n)
            dlkm2 = dlkm1
            dlkm1 = dlk
        return pol

    zscale = 1.0e100
    ziscale = 1.0e-100
    zlsita = np.ones(nlat)
    icorr3 = np.zeros(nlat, dtype=np.int64)
    for _ in range(km // 2):
        zlsita = zlsita * cos2
        small = np.abs(zlsita) < ziscale
        zlsita = np.where(small, zlsita * zscale, zlsita)
        icorr3 += small
    if km % 2 == 1:
        zlsita = zlsita * cost

    zfac = 1.0
    for n in range(1, km):
        zfac = zfac * math.sqrt(float(2 * n - 1))
        zfac = zfac / math.sqrt(float(2 * n))
    zfac = zfac * math.sqrt(float(2 * km - 1))

    zfac0 = 1.0
    for ic in range(0, min(knsmax - km, 3) + 1):
        zfac0 = zfac0 * float(2 * km + ic)
        if ic == 0:
            zfac1 = 1.0
            zmult = zfac * np.ones(nlat)
        elif ic == 1:
            zfac1 = 1.0
            zfac = zfac * float(2 * km + 1)
            zmult = zfac * x
        elif ic == 2:
            zfac1 = 2.0
            zmult = 0.5 * zfac * (float(2 * km + 3) * x * x - 1.0)
        else:
            zfac1 = 6.0
            zfac = zfac * float(2 * km + 3)
            zmult = (1.0 / 6.0) * x * zfac * (float(2 * km + 5) * x * x - 3.0)
        pol[km + ic] = zlsita * zmult * math.sqrt(2.0 * (float(km + ic) + 0.5) * zfac1 / zfac0)

    # ICORR(n) = ICORR3 - (number of rescalings whose JN-4 <= n): keep the deltas
    dcorr = np.zeros((knsmax + 2, nlat), dtype=np.int64)
    istart = 1 if kcheap == 3 else 0
    iinc = 2 if kcheap in (2, 3) else 1

    def dcl(k):
        return math.sqrt((float(k - km + 1) * float(k - km + 2) * float(k + km + 1) * float(k + km + 2))
                         / (float(2 * k + 1) * float(2 * k + 3) * float(2 * k + 3) * float(2 * k + 5)))

    def ddl(k):
        return (2.0 * float(k) * float(k + 1) - 2.0 * float(km * km) - 1.0) / (float(2 * k - 1) * float(2 * k + 3))

    x2 = x * x
    for n in range(km + istart + 4, knsmax + 1, iinc):
        big = np.abs(pol[n - 4]) > zscale
        if big.any():
            pol[n - 4:n, big] = pol[n - 4:n, big] / zscale
            dcorr[n - 4, big] -= 1
        pol[n] = ((x2 - ddl(n - 2)) * pol[n - 2] - dcl(n - 4) * pol[n - 4]) / dcl(n - 2)

    icorr = icorr3[None, :] + np.cumsum(dcorr[:knsmax + 1], axis=0)
    sel = np.arange(km + istart, knsmax + 1, iinc)
    if sel.size:
        sub = pol[sel]
        ic = icorr[sel].copy()
        while True:
            act = ic > 0
            if not act.any():
                break
            sub = np.where(act, sub / zscale, sub)
            sub = np.where(act & (sub < eps), eps, sub)
            ic = ic - act
        pol[sel] = sub
    return pol


# --------------------------------------------------------------------------
#  Decomposition helpers
# --------------------------------------------------------------------------
def suwavedi(nsmax: int, nprtrw: int):
    """Zig-zag distribution of zonal wavenumbers over the W-set.
    common/internal/suwavedi_mod.F90:118-137.  Returns nprocm[m] (0-based
    rank) and per-rank list of m in local order (MYMS)."""
    nprocm = np.zeros(nsmax + 1, dtype=np.int64)
    myms = [[] for _ in range(nprtrw)]
    ind = 1
    ik = 0
    for jm in range(nsmax + 1):
        ik += ind
        if ik > nprtrw:
            ik = nprtrw
            ind = -1
        elif ik < 1:
            ik = 1
            ind = 1
        nprocm[jm] = ik - 1
        myms[ik - 1].append(jm)
    return nprocm, myms


def sumplatb_fourier(nloen, nproca: int):
    """Latitude -> W-set rank distribution in Fourier space (LDSPLIT=.F.,
    LDFOURIER=.T.).  common/internal/sumplatb_mod.F90:171-216 and
    sumplatf_mod.F90:110-139.  Returns (first, count) per rank, 0-based."""
    ndgl = len(nloen)
    icost = [int(v) for v in nloen]
    imedia = sum(icost)
    kmediap = imedia // nproca
    krestm = imedia - kmediap * nproca
    if krestm > 0:
        kmediap += 1
    klast = [0] * (nproca + 1)  # 1-based
    itot_top = 0
    itot_bot = 0
    igl_top = 1
    igl_bot = ndgl
    for ja in range(1, (nproca - 1) // 2 + 2):
        if ja != nproca // 2 + 1:
            while True:
                if igl_top <= ndgl and itot_top + icost[igl_top - 1] < kmediap:
                    klast[ja] = igl_top
                    itot_top += icost[igl_top - 1]
                    igl_top += 1
                else:
                    itot_top -= kmediap
                    break
            klast[nproca - ja + 1] = igl_bot
            while True:
                if igl_bot >= 1 and itot_bot + icost[igl_bot - 1] < kmediap:
                    itot_bot += icost[igl_bot - 1]
                    igl_bot -= 1
                else:
                    itot_bot -= kmediap
                    break
        else:
            klast[ja] = igl_bot
    if any(klast[ja] == 0 for ja in range(1, nproca + 1)):
        ilats = [0] * (nproca + 1)
        ia = 0
        for _ in range(ndgl):
            ia += 1
            ilats[ia] += 1
            if ia == nproca:
                ia = 0
        klast[1] = ilats[1]
        for ja in range(2, nproca + 1):
            klast[ja] = klast[ja - 1] + ilats[ja]
    first, count = [], []
    prev = 0
    for ja in range(1, nproca + 1):
        cnt = klast[ja] - prev if klast[ja] != 0 else 0
        first.append(prev)
        count.append(cnt)
        prev += cnt
    return first, count


# --------------------------------------------------------------------------
#  Setup
# --------------------------------------------------------------------------
def octahedral_nloen(n: int) -> np.ndarray:
    """O<N> grid: nloen(i) = 20 + 4(i-1), mirrored.
    src/programs/ectrans-benchmark.F90:1043-1047."""
    half = 20 + 4 * np.arange(n, dtype=np.int64)
    return np.concatenate([half, half[::-1]])


@dataclass
class Setup:
    nsmax: int
    ndgl: int
    nloen: np.ndarray
    ndgnh: int = 0
    rmu: np.ndarray = None
    rw: np.ndarray = None
    r1mu2: np.ndarray = None
    racthe: np.ndarray = None
    nmen: np.ndarray = None
    ndglu: np.ndarray = None
    nasm0: np.ndarray = None  # 0-based offset of (m, n=m, re) in single-rank layout
    nspec2: int = 0
    ngptot: int = 0
    latoff: np.ndarray = None
    pa: dict = field(default_factory=dict)  # m -> [ndglu, ILA]  n = m+1, m+3, ... (ascending)
    ps: dict = field(default_factory=dict)  # m -> [ndglu, ILS]  n = m, m+2, ...
    ms: list = None


def compute_nmen(nsmax: int, ndgl: int, nloen, r1mu2, reduced: bool = True) -> np.ndarray:
    """common/internal/setup_geom_mod.F90:44-78."""
    ndgnh = (ndgl + 1) // 2
    nmen = np.zeros(ndgl, dtype=np.int64)
    nsmaxlin = ndgl - 1
    if nsmax >= nsmaxlin or not reduced:
        for j in range(ndgl):
            nmen[j] = min(nsmax, (int(nloen[j]) - 1) // 2)
        return nmen
    if nsmax >= ndgl * 2 // 3 - 1:
        fac = float(3 * (nsmaxlin - nsmax) // ndgl)
        zsqm2 = fac * r1mu2
        sub = 0
    else:
        zsqm2 = r1mu2
        sub = 1

    def val(j):
        return int((float(int(nloen[j]) - 1)) / (2.0 + zsqm2[j])) - sub

    nmen[0] = min(nsmax, val(0))
    for j in range(1, ndgnh):
        nmen[j] = min(nsmax, max(nmen[j - 1], val(j)))
    nmen[ndgl - 1] = min(nsmax, val(ndgl - 1))
    for j in range(ndgl - 2, ndgnh - 1, -1):
        nmen[j] = min(nsmax, max(nmen[j + 1], val(j)))
    return nmen


def setup(nsmax: int, ndgl: int, nloen, ms=None, tables: bool = True) -> Setup:
    """SETUP_TRANS for one task.  cpu/external/setup_trans.F90:169-428,
    SULEG cpu/internal/suleg_mod.F90 (Gaussian latitudes :249-293, R1MU2 and
    RACTHE :386-394, table fill via SUPOLF :597-760 and :890-960),
    SETUP_GEOM, SUWAVEDI (NASM0)."""
    nloen = np.asarray(nloen, dtype=np.int64)
    assert nloen.size == ndgl and ndgl % 2 == 0
    s = Setup(nsmax=nsmax, ndgl=ndgl, nloen=nloen)
    s.ndgnh = (ndgl + 1) // 2
    s.rmu, s.rw = gauss_latitudes(ndgl)
    theta = np.arcsin(s.rmu)
    zcos = np.cos(theta)
    s.r1mu2 = zcos ** 2
    s.racthe = 1.0 / zcos / RA
    reduced = not np.all(nloen == nloen[0])
    s.nmen = compute_nmen(nsmax, ndgl, nloen, s.r1mu2, reduced)
    s.ndglu = np.zeros(nsmax + 1, dtype=np.int64)
    for m in range(nsmax + 1):
        s.ndglu[m] = int(np.sum(s.nmen[:s.ndgnh] >= m))
    s.nasm0 = np.zeros(nsmax + 1, dtype=np.int64)
    pos = 0
    for m in range(nsmax + 1):
        s.nasm0[m] = pos
        pos += 2 * (nsmax - m + 1)
    s.nspec2 = pos
    s.ngptot = int(nloen.sum())
    s.latoff = np.concatenate([[0], np.cumsum(nloen)[:-1]]).astype(np.int64)
    s.ms = list(range(nsmax + 1)) if ms is None else list(ms)
    if tables:
        imaxn = nsmax + 1
        for m in s.ms:
            idglu = min(s.ndgnh, int(s.ndglu[m]))
            isl = max(s.ndgnh - int(s.ndglu[m]) + 1, 1)
            ila = (nsmax - m + 2) // 2
            ils = (nsmax - m + 3) // 2
            lat = s.rmu[isl - 1:isl - 1 + idglu]
            inmax_a = imaxn + 1 if (imaxn - m) % 2 == 0 else imaxn
            inmax_s = imaxn if (imaxn - m) % 2 == 0 else imaxn + 1
            if idglu == 0:
                s.pa[m] = np.zeros((0, ila))
                s.ps[m] = np.zeros((0, ils))
                continue
            pol = supolf(m, inmax_a, lat, kcheap=3)
            s.pa[m] = np.ascontiguousarray(pol[m + 1:m + 2 * ila:2].T)
            pol = supolf(m, inmax_s, lat, kcheap=2)
            s.ps[m] = np.ascontiguousarray(pol[m:m + 2 * ils - 1:2].T)
    return s


def epsnm(m: int, n) -> np.ndarray:
    """common/internal/pre_suleg_mod.F90:46-65."""
    n = np.asarray(n, dtype=np.float64)
    return np.sqrt((n * n - float(m * m)) / (4.0 * n * n - 1.0))


def rlapin(n) -> np.ndarray:
    n = np.asarray(n, dtype=np.float64)
    out = np.zeros_like(n)
    pos = n >= 1
    out[pos] = -(RA * RA / (n[pos] * (n[pos] + 1.0)))
    return out


# --------------------------------------------------------------------------
#  Spectral helpers
# --------------------------------------------------------------------------
def spec_m(s: Setup, sp: np.ndarray, m: int) -> np.ndarray:
    """Complex coefficients [nfld, T-m+1] of wavenumber m from sp[nfld, nspec2]
    (single-task layout: m-major, n ascending, re/im interleaved,
    suwavedi_mod.F90:130-135)."""
    o = int(s.nasm0[m])
    cnt = s.nsmax - m + 1
    blk = sp[:, o:o + 2 * cnt]
    return blk[:, 0::2] + 1j * blk[:, 1::2]


def specnorm(s: Setup, sp: np.ndarray, pmet=None) -> np.ndarray:
    """cpu/internal/spnormd_mod.F90:36-51 + spnorm_ctl_mod.F90:56-57; pmet: optional metric (0:nsmax)."""
    out = np.zeros(sp.shape[0])
    for m in range(s.nsmax + 1):
        c = spec_m(s, sp, m)
        w = np.ones(s.nsmax + 1 - m) if pmet is None else np.asarray(pmet, dtype=np.float64)[m:]
        if m == 0:
            out += np.sum(w * c.real ** 2, axis=1)
        else:
            out += 2.0 * np.sum(w * np.abs(c) ** 2, axis=1)
    return np.sqrt(out)


def _vdtuv(s: Setup, m: int, vor: np.ndarray, div: np.ndarray):
    """(vor, div)[nfld, n=m..T] -> (U, V)[nfld, n=m..T+1].
    cpu/internal/vdtuv_mod.F90:97-143."""
    T = s.nsmax
    n = np.arange(m, T + 2)
    nf = vor.shape[0]

    def ext(a):  # index by n in m-1 .. T+2
        e = np.zeros((nf, T + 4 - m + 0), dtype=np.complex128)  # n = m-1 .. T+2
        e[:, 1:1 + (T - m + 1)] = a
        return e

    ve, de = ext(vor), ext(div)
    idx = n - (m - 1)
    lap = rlapin(np.arange(-1, T + 4))  # index by n+1

    def lp(nn):
        return lap[np.asarray(nn) + 1]

    eps_n = epsnm(m, n)
    eps_np1 = epsnm(m, n + 1)
    c1 = (n - 1.0) * eps_n * lp(n - 1)
    c2 = (n + 2.0) * eps_np1 * lp(n + 1)
    zkm = float(m)
    u = 1j * zkm * lp(n) * de[:, idx] + c1 * ve[:, idx - 1] - c2 * ve[:, idx + 1]
    v = 1j * zkm * lp(n) * ve[:, idx] - c1 * de[:, idx - 1] + c2 * de[:, idx + 1]
    return u, v


def _spnsde(s: Setup, m: int, f: np.ndarray):
    """N-S derivative, n = m..T+1.  cpu/internal/spnsde_mod.F90:95-114."""
    T = s.nsmax
    n = np.arange(m, T + 2)
    nf = f.shape[0]
    e = np.zeros((nf, T + 4 - m), dtype=np.complex128)
    e[:, 1:1 + (T - m + 1)] = f
    idx = n - (m - 1)
    return -(n - 1.0) * epsnm(m, n) * e[:, idx - 1] + (n + 2.0) * epsnm(m, n + 1) * e[:, idx + 1]



def ltinv_m(s: Setup, m: int, spvor, spdiv, spscalar, scders=False, vorgp=False, divgp=False):
    """LTINV for one zonal wavenumber: returns (north, south) Fourier coefficients
    [ndglu(m), kf_out_lt] complex.  cpu/internal/ltinv_mod.F90:139-320."""
    kf_uv = 0 if spvor is None else spvor.shape[0]
    kf_sc = 0 if spscalar is None else spscalar.shape[0]
    cols = []
    if kf_uv:
        vor = spec_m(s, spvor, m)
        div = spec_m(s, spdiv, m)
        u, v = _vdtuv(s, m, vor, div)
        pad = np.zeros((kf_uv, 1), dtype=np.complex128)
        if vorgp:
            cols.append(np.concatenate([vor, pad], axis=1))
        if divgp:
            cols.append(np.concatenate([div, pad], axis=1))
        cols += [u, v]
    if kf_sc:
        sc = spec_m(s, spscalar, m)
        cols.append(np.concatenate([sc, np.zeros((kf_sc, 1), dtype=np.complex128)], axis=1))
        if scders:
            cols.append(_spnsde(s, m, sc))
    x = np.concatenate(cols, axis=0).T  # [n=m..T+1, fld]
    if m == 0:
        x = x.real.astype(np.complex128)  # KM=0: imaginary columns skipped (leinv_mod.F90:103-110)
    xs, xa = x[0::2], x[1::2]  # n-m even (symmetric), odd (antisymmetric)
    pa, ps = s.pa[m], s.ps[m]
    za = (pa @ xa.real) + 1j * (pa @ xa.imag)
    zs = (ps @ xs.real) + 1j * (ps @ xs.imag)
    return zs + za, zs - za


def ledir_m(s: Setup, m: int, fn: np.ndarray, fs: np.ndarray, kf_uv: int):
    """PRFI2B + LDFOU2 + LEDIR for one m: north/south Fourier coefficients [ndglu, nfld] ->
    spectral coefficients [n = m..T+1, nfld].  prfi2b_mod.F90:84-94, ldfou2_mod.F90:90-96,
    ledir_mod.F90:118-261."""
    T = s.nsmax
    ndglu = int(s.ndglu[m])
    isl = s.ndgnh - ndglu
    ila = (T - m + 2) // 2
    ils = (T - m + 3) // 2
    sym = fn + fs
    asym = fn - fs
    ract = s.racthe[isl:isl + ndglu][:, None]
    sym[:, :2 * kf_uv] *= ract
    asym[:, :2 * kf_uv] *= ract
    wgt = s.rw[isl:isl + ndglu][:, None]
    zb_a = asym * wgt
    zb_s = sym * wgt
    if m == 0:
        zb_a = zb_a.real.astype(np.complex128)
        zb_s = zb_s.real.astype(np.complex128)
    pa, ps = s.pa[m], s.ps[m]
    ca = (pa.T @ zb_a.real) + 1j * (pa.T @ zb_a.imag)  # n = m+1, m+3, ...
    cs = (ps.T @ zb_s.real) + 1j * (ps.T @ zb_s.imag)  # n = m, m+2, ...
    oa = np.zeros((T + 2 - m, fn.shape[1]), dtype=np.complex128)  # n = m..T+1
    oa[0::2] = cs[:ils]
    oa[1::2] = ca[:ila]
    return oa

# --------------------------------------------------------------------------
#  Inverse transform
# --------------------------------------------------------------------------
def inv_trans(s: Setup, spvor=None, spdiv=None, spscalar=None, scders=False,
              vorgp=False, divgp=False, uvder=False, return_fourier=False):
    """INV_TRANS, one task.  Returns gp[nfld_gp, ngptot] with the field order of
    include/ectrans/inv_trans.h:66-76:  [vor][div] u v scalars [NS-ders]
    [EW du, dv] [EW-ders].

    LTINV cpu/internal/ltinv_mod.F90:139-320 (PRFI1B, VDTUV, SPNSDE, LEINV
    leinv_mod.F90:116-186, ASRE1B asre1b_mod.F90:88-102), FOURIER_IN, FSC
    fsc_mod.F90:138-187, FTINV ftinv_mod.F90:65-84 with FFTW's unnormalised c2r
    (tpm_fftw.F90:163,286-303).
    """
    T = s.nsmax
    kf_uv = 0 if spvor is None else spvor.shape[0]
    kf_sc = 0 if spscalar is None else spscalar.shape[0]
    n_vor = kf_uv if vorgp else 0
    n_div = kf_uv if divgp else 0
    n_nsd = kf_sc if scders else 0
    kf_out_lt = n_vor + n_div + 2 * kf_uv + kf_sc + n_nsd
    four = [np.zeros((int(s.nmen[j]) + 1, kf_out_lt), dtype=np.complex128) for j in range(s.ndgl)]
    for m in s.ms:
        ndglu = int(s.ndglu[m])
        if ndglu == 0:
            continue
        north, south = ltinv_m(s, m, spvor, spdiv, spscalar, scders, vorgp, divgp)
        isl = s.ndgnh - ndglu  # 0-based first northern latitude
        for i in range(ndglu):
            jn = isl + i
            js = s.ndgl - 1 - jn
            four[jn][m] = north[i]
            four[js][m] = south[i]
    if return_fourier:
        return four
    n_uvd = 2 * kf_uv if uvder else 0
    kf_fs = kf_out_lt + n_uvd + n_nsd
    gp = np.zeros((kf_fs, s.ngptot))
    i_uv = n_vor + n_div
    i_sc = i_uv + 2 * kf_uv
    i_nsd = i_sc + kf_sc
    for j in range(s.ndgl):
        nlon = int(s.nloen[j])
        imen = int(s.nmen[j])
        f = four[j].copy()  # [m, fld]
        ract = s.racthe[j]
        f[:, i_uv:i_uv + 2 * kf_uv] *= ract
        if scders:
            f[:, i_nsd:i_nsd + n_nsd] *= ract
        parts = [f]
        mm = np.arange(imen + 1, dtype=np.float64)[:, None] * ract
        if uvder:
            parts.append(1j * mm * f[:, i_uv:i_uv + 2 * kf_uv])
        if scders:
            parts.append(1j * mm * f[:, i_sc:i_sc + kf_sc])
        f = np.concatenate(parts, axis=1)
        half = np.zeros((nlon // 2 + 1, kf_fs), dtype=np.complex128)
        half[:imen + 1] = f
        row = np.fft.irfft(half, n=nlon, axis=0) * nlon
        gp[:, s.latoff[j]:s.latoff[j] + nlon] = row.T
    return gp


# --------------------------------------------------------------------------
#  Direct transform
# --------------------------------------------------------------------------
def dir_trans(s: Setup, gp: np.ndarray, kf_uv: int = 0, kf_sc: int = 0):
    """DIR_TRANS, one task.  gp[2*kf_uv + kf_sc, ngptot] in the order u, v,
    scalars (include/ectrans/dir_trans.h:59-61).  Returns (spvor, spdiv,
    spscalar), each [nfld, nspec2] (None when absent).

    FTDIR cpu/internal/ftdir_mod.F90:67-84 with r2c then /N
    (tpm_fftw.F90:310-321), FOURIER_OUT, PRFI2B prfi2b_mod.F90:84-94, LDFOU2
    ldfou2_mod.F90:90-96, LEDIR ledir_mod.F90:118-261, UVTVD
    uvtvd_mod.F90:91-139, UPDSP updsp_mod.F90:104-161, UPDSPB
    updspb_mod.F90:92-149.
    """
    T = s.nsmax
    kf_fs = 2 * kf_uv + kf_sc
    assert gp.shape[0] == kf_fs
    four = []
    for j in range(s.ndgl):
        nlon = int(s.nloen[j])
        imen = int(s.nmen[j])
        row = gp[:, s.latoff[j]:s.latoff[j] + nlon]
        f = np.fft.rfft(row, axis=1) / nlon
        four.append(f[:, :imen + 1].T.copy())  # [m, fld]
    spvor = np.zeros((kf_uv, s.nspec2)) if kf_uv else None
    spdiv = np.zeros((kf_uv, s.nspec2)) if kf_uv else None
    spsc = np.zeros((kf_sc, s.nspec2)) if kf_sc else None
    for m in s.ms:
        ndglu = int(s.ndglu[m])
        isl = s.ndgnh - ndglu
        ila = (T - m + 2) // 2
        ils = (T - m + 3) // 2
        fn = np.zeros((ndglu, kf_fs), dtype=np.complex128)
        fs = np.zeros((ndglu, kf_fs), dtype=np.complex128)
        for i in range(ndglu):
            jn = isl + i
            js = s.ndgl - 1 - jn
            fn[i] = four[jn][m]
            fs[i] = four[js][m]
        oa = ledir_m(s, m, fn, fs, kf_uv)
        o = int(s.nasm0[m])
        cnt = T - m + 1

        def put(dst, coef):  # coef [nfld, cnt]
            dst[:, o:o + 2 * cnt:2] = coef.real
            dst[:, o + 1:o + 2 * cnt:2] = 0.0 if m == 0 else coef.imag

        if kf_uv:
            u = oa[:, :kf_uv].T  # [fld, n=m..T+1]
            v = oa[:, kf_uv:2 * kf_uv].T
            n = np.arange(m, T + 1)
            ue = np.zeros((kf_uv, T + 3 - m), dtype=np.complex128)  # n = m-1..T+1
            ve = np.zeros((kf_uv, T + 3 - m), dtype=np.complex128)
            ue[:, 1:] = u
            ve[:, 1:] = v
            idx = n - (m - 1)
            c1 = n * epsnm(m, n + 1)
            c2 = (n + 1.0) * epsnm(m, n)
            zkm = float(m)
            vor = 1j * zkm * ve[:, idx] - c1 * ue[:, idx + 1] + c2 * ue[:, idx - 1]
            div = 1j * zkm * ue[:, idx] + c1 * ve[:, idx + 1] - c2 * ve[:, idx - 1]
            put(spvor, vor)
            put(spdiv, div)
            if m == 0:
                spvor[:, o] = 0.0
                spdiv[:, o] = 0.0
        if kf_sc:
            put(spsc, oa[:cnt, 2 * kf_uv:].T)
    return spvor, spdiv, spsc


# --------------------------------------------------------------------------
#  Synthetic inputs
# --------------------------------------------------------------------------
def random_spectral(s: Setup, nfld: int, seed: int, zero00: bool = False, decay: bool = False) -> np.ndarray:
    """Seeded uniform(-0.1, 0.1) spectral coefficients with Im(m=0) = 0 (mirrors
    tests/trans/test_invtrans_adjoint.F90:152-159)."""
    rng = np.random.default_rng(seed)
    sp = rng.uniform(-0.1, 0.1, size=(nfld, s.nspec2))
    cnt0 = s.nsmax + 1
    sp[:, 1:2 * cnt0:2] = 0.0
    if zero00:
        sp[:, 0] = 0.0
    if decay:
        for m in range(s.nsmax + 1):
            o = int(s.nasm0[m])
            n = np.arange(m, s.nsmax + 1, dtype=np.float64)
            sc = np.repeat((n + 1.0) ** -2, 2)
            sp[:, o:o + sc.size] *= sc
    return sp


def benchmark_spectral(s: Setup, nfld: int) -> np.ndarray:
    """ectrans-benchmark input: Re psi(m=4, n=19) = 1, everything else 0.
    src/programs/ectrans-benchmark.F90:1389-1415."""
    sp = np.zeros((nfld, s.nspec2))
    if s.nsmax >= 19:
        sp[:, int(s.nasm0[4]) + 2 * (19 - 4)] = 1.0
    return sp


# ---------------------------------------------------------------------------------------------
# Adjoints (INV_TRANSAD / DIR_TRANSAD).  The reference hand-codes them (cpu/internal/*ad_mod.F90); what pins them
# is the adjoint identity of its tests (tests/trans/test_invtrans_adjoint.F90:192-222, test_dirtrans_adjoint.F90)
# with the inner products <a, b>_gp = sum_j a_j b_j and <x, y>_sp = sum mfact (Re Re + Im Im), mfact = 2 for m > 0
# and 1 for the real parts of m = 0 (SCALPRODSP :276-311).  The oracle therefore builds the matrix of the forward
# transform column by column and transposes it -- exact by construction, affordable only at small truncations.
# ---------------------------------------------------------------------------------------------
def spectral_weights(s: Setup) -> np.ndarray:
    """mfact per local spectral index; 0 for the (ignored) imaginary parts of m = 0."""
    w = np.zeros(s.nspec2)
    for m in range(s.nsmax + 1):
        n = 2 * (s.nsmax - m + 1)
        a = s.nasm0[m]
        if m == 0:
            w[a:a + n:2] = 1.0
        else:
            w[a:a + n] = 2.0
    return w


def _forward_matrices(s: Setup, uv: bool):
    """F (grid x spectral) of INV_TRANS and D (spectral x grid) of DIR_TRANS for one scalar field (uv False) or for
    one (vor, div) -> (u, v) level (uv True; vectors are [vor; div] and [u; v])."""
    key = ("_adj_uv" if uv else "_adj_sc")
    if key in s.__dict__:
        return s.__dict__[key]
    ns, ng = s.nspec2, s.ngptot
    if not uv:
        F = np.zeros((ng, ns)); D = np.zeros((ns, ng))
        for i in range(ns):
            e = np.zeros((1, ns)); e[0, i] = 1.0
            F[:, i] = inv_trans(s, spscalar=e)[0]
        for j in range(ng):
            e = np.zeros((1, ng)); e[0, j] = 1.0
            D[:, j] = dir_trans(s, e, 0, 1)[2][0]
    else:
        F = np.zeros((2 * ng, 2 * ns)); D = np.zeros((2 * ns, 2 * ng))
        z = np.zeros((1, ns))
        for i in range(ns):
            e = np.zeros((1, ns)); e[0, i] = 1.0
            g = inv_trans(s, e, z, None); F[:, i] = np.concatenate([g[0], g[1]])
            g = inv_trans(s, z, e, None); F[:, ns + i] = np.concatenate([g[0], g[1]])
        zg = np.zeros(ng)
        for j in range(ng):
            e = np.zeros(ng); e[j] = 1.0
            v, d, _ = dir_trans(s, np.stack([e, zg]), 1, 0); D[:, j] = np.concatenate([v[0], d[0]])
            v, d, _ = dir_trans(s, np.stack([zg, e]), 1, 0); D[:, ng + j] = np.concatenate([v[0], d[0]])
    s.__dict__[key] = (F, D)
    return F, D


def inv_transad(s: Setup, gp: np.ndarray, kf_uv: int = 0, kf_sc: int = 0):
    """M^-1 F^T gp; gp (2 kf_uv + kf_sc, ngptot) ordered u, v, scalars -> (vor_ad, div_ad, sc_ad)."""
    w = spectral_weights(s)
    winv = np.where(w > 0, 1.0 / np.maximum(w, 1e-300), 0.0)
    ns = s.nspec2
    vor = np.zeros((kf_uv, ns)); div = np.zeros((kf_uv, ns)); sc = np.zeros((kf_sc, ns))
    if kf_uv:
        F, _ = _forward_matrices(s, True)
        for j in range(kf_uv):
            r = F.T @ np.concatenate([gp[j], gp[kf_uv + j]])
            vor[j], div[j] = winv * r[:ns], winv * r[ns:]
    if kf_sc:
        F, _ = _forward_matrices(s, False)
        for j in range(kf_sc):
            sc[j] = winv * (F.T @ gp[2 * kf_uv + j])
    return vor, div, sc


def dir_transad(s: Setup, spvor=None, spdiv=None, spscalar=None) -> np.ndarray:
    """D^T M psi -> gp (2 kf_uv + kf_sc, ngptot) ordered u, v, scalars."""
    w = spectral_weights(s)
    kf_uv = 0 if spvor is None else spvor.shape[0]
    kf_sc = 0 if spscalar is None else spscalar.shape[0]
    ng = s.ngptot
    out = np.zeros((2 * kf_uv + kf_sc, ng))
    if kf_uv:
        _, D = _forward_matrices(s, True)
        for j in range(kf_uv):
            r = D.T @ np.concatenate([w * spvor[j], w * spdiv[j]])
            out[j], out[kf_uv + j] = r[:ng], r[ng:]
    if kf_sc:
        _, D = _forward_matrices(s, False)
        for j in range(kf_sc):
            out[2 * kf_uv + j] = D.T @ (w * spscalar[j])
    return out


# --------------------------------------------------------------------------
#  GPNORM_TRANS, VORDIV_TO_UV, TRANS_INQ(PRPNM)
# --------------------------------------------------------------------------
def gpnorm_trans(s: Setup, gp: np.ndarray):
    """(ave, min, max) per field of gp[nfld, ngptotg]: per-latitude sums weighted by RW(lat)/NLOEN(lat), added in
    latitude order.  cpu/internal/gpnorm_trans_ctl_mod.F90:170-215, :422-428."""
    nf = gp.shape[0]
    ave = np.zeros(nf)
    off = 0
    for j in range(s.ndgl):
        n = int(s.nloen[j])
        row = gp[:, off:off + n]
        ave = ave + row.sum(axis=1) * s.rw[j] / n
        off += n
    return ave, gp.min(axis=1), gp.max(axis=1)


def vordiv_to_uv(s: Setup, spvor: np.ndarray, spdiv: np.ndarray):
    """Spectral (vor, div)[nfld, nspec2] -> spectral (U, V) cos(theta) [nfld, nspec2], rows n <= T, scaled by 1/a.
    cpu/internal/vd2uv_mod.F90:86-112 (VDTUV as in the inverse transform, the n = T+1 row dropped)."""
    u = np.zeros_like(spvor)
    v = np.zeros_like(spvor)
    for m in range(s.nsmax + 1):
        um, vm = _vdtuv(s, m, spec_m(s, spvor, m) if m else spec_m(s, spvor, m).real + 0j,
                        spec_m(s, spdiv, m) if m else spec_m(s, spdiv, m).real + 0j)
        o = int(s.nasm0[m])
        cnt = s.nsmax - m + 1
        for dst, src in ((u, um), (v, vm)):
            dst[:, o:o + 2 * cnt:2] = src[:, :cnt].real / RA
            dst[:, o + 1:o + 2 * cnt:2] = 0.0 if m == 0 else src[:, :cnt].imag / RA
    return u, v


def rpnm_reference_layout(s: Setup) -> np.ndarray:
    """PRPNM(NDGNH, NSPOLEGL) as TRANS_INQ returns it (cpu/external/trans_inq.F90:444-464): the block of wavenumber m
    starts at NPMS(m) = sum_{m' < m} (T + 2 - m'), its column p = 1 .. T+2-m holds n = T + 2 - p; rows above
    ISL = NDGNH - NDGLU(m) + 1 stay zero."""
    T, ndgnh = s.nsmax, s.ndgl // 2
    ncol = sum(T + 2 - m for m in range(T + 1))
    out = np.zeros((ndgnh, ncol))
    c0 = 0
    for m in range(T + 1):
        nl = int(s.ndglu[m])
        for n in range(m, T + 2):
            col = s.ps[m][:, (n - m) // 2] if (n - m) % 2 == 0 else s.pa[m][:, (n - m - 1) // 2]
            out[ndgnh - nl:, c0 + (T + 2 - n) - 1] = col
        c0 += T + 2 - m
    return out


# --------------------------------------------------------------------------
#  Grid-point decomposition: eq_regions + SUMPLAT (LDSPLIT=T) + SUSTAONL
#  (the default of ectrans-benchmark: LDEQ_REGIONS=T, LDSPLIT=T, ectrans-benchmark.F90:385)
# --------------------------------------------------------------------------
def _eq_gamma(x: float) -> float:
    """The reference's own gamma function (eq_regions_mod.F90:282-330), used for the area of the sphere."""
    p = [0.999999999999999990e+00, -0.422784335098466784e+00, -0.233093736421782878e+00, 0.191091101387638410e+00,
         -0.024552490005641278e+00, -0.017645244547851414e+00, 0.008023273027855346e+00, -0.000804329819255744e+00,
         -0.000360837876648255e+00, 0.000145596568617526e+00, -0.000017545539395205e+00, -0.000002591225267689e+00,
         0.000001337767384067e+00, -0.000000199542863674e+00]
    n = int(round(x - 2))
    w = x - (n + 2)
    y = 0.0
    for c in reversed(p):
        y = y * w + c
    if n > 0:
        w = x - 1
        for k in range(2, n + 1):
            w = w * (x - k)
    else:
        w = 1.0
        for k in range(0, -n):
            y = y * (x + k)
    return w / y


def _nint(x: float) -> int:
    """Fortran NINT: half away from zero."""
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def eq_regions(n: int):
    """Number of regions per latitude band of the equal-area partition of the sphere into n regions (Leopardi's
    recursive zonal partition as coded in common/internal/eq_regions_mod.F90:76-230)."""
    if n == 1:
        return [1]
    pi = 2.0 * math.asin(1.0)
    area = (2.0 * pi ** 1.5 / _eq_gamma(1.5)) / n                     # area_of_ideal_region
    cap = lambda s: 4.0 * pi * math.sin(s / 2.0) ** 2                  # area_of_cap
    c_polar = pi / 2.0 if n == 2 else 2.0 * math.asin(math.sqrt(area / pi) / 2.0)
    a_ideal = area ** 0.5
    n_collars = max(1, _nint((pi - 2.0 * c_polar) / a_ideal)) if (n > 2 and a_ideal > 0) else 0
    r = [1.0]
    if n_collars > 0:
        a_fit = (pi - 2.0 * c_polar) / n_collars
        for c in range(1, n_collars + 1):
            r.append((cap(c_polar + c * a_fit) - cap(c_polar + (c - 1) * a_fit)) / area)
    r.append(1.0)
    out, disc = [], 0.0
    for v in r:                                                        # round_to_naturals
        k = _nint(v + disc)
        disc += v - k
        out.append(k)
    assert sum(out) == n
    return out


def sumplat_eq_split(nloen, nproc: int, n_regions):
    """SUMPLATBEQ (LDSPLIT=T, unweighted; sumplatbeq_mod.F90:84-150) + SUMPLAT (sumplat_mod.F90:138-150): for every
    A-set (band) its number of points and first / last latitude (0-based, a split latitude belongs to both)."""
    nloen = [int(x) for x in nloen]
    ndgl = len(nloen)
    total = sum(nloen)
    mediap = total // nproc
    restm = total - mediap * nproc
    if restm > 0:
        mediap += 1
    nprocagp, last, indic = [], [], []
    rest, ilast, ipe = 0, -1, 0
    for ja, nb in enumerate(n_regions):
        comp = 0
        for _ in range(nb):
            ipe += 1
            comp += mediap if (ipe <= restm or restm == 0) else mediap - 1
        nprocagp.append(comp)
        itot = rest
        for jgl in range(ilast + 1, ndgl):
            ilast = jgl
            if itot + nloen[jgl] < comp:
                itot += nloen[jgl]
            elif itot + nloen[jgl] == comp:
                rest = 0
                last.append(jgl); indic.append(-1)
                break
            else:
                rest = nloen[jgl] - (comp - itot)
                last.append(jgl); indic.append(jgl)
                break
    na = len(n_regions)
    frst, lst = [0] * na, [0] * na
    lst[na - 1] = ndgl - 1
    for ja in range(na - 1):
        if indic[ja] < 0:
            frst[ja + 1] = last[ja] + 1
            lst[ja] = last[ja]
        else:
            frst[ja + 1] = indic[ja]
            lst[ja] = indic[ja]
    return nprocagp, frst, lst


def gridpoint_partition(nloen, nproc: int):
    """Grid-point decomposition of the reference for LDEQ_REGIONS=T, LDSPLIT=T (SUSTAONL, sustaonl_mod.F90:118-196):
    inside a band the points go to the B-sets one at a time, always from the latitude whose next point lies furthest
    west (angle in 1/1000 degree, NINT; ties -> the northernmost latitude).  Returns (n_regions, segs) with
    segs[pe] = [(lat, first point, count), ...] (0-based) in the task's local point order; pe = sum(n_regions[:a]) + b."""
    nloen = [int(x) for x in nloen]
    nreg = eq_regions(nproc)
    if nproc == 1:
        return nreg, [[(j, 0, nloen[j]) for j in range(len(nloen))]]
    agp, frst, lst = sumplat_eq_split(nloen, nproc, nreg)
    segs = []
    gpta = 0
    for ja, nb in enumerate(nreg):
        lats = list(range(frst[ja], lst[ja] + 1))
        gptprsets = sum(nloen[:frst[ja]])
        gpts = agp[ja]
        ix = [1] * len(lats)                       # IXPTLAT (1-based next point)
        ilst = [nloen[j] for j in lats]            # ILSTPTLAT
        ix[0] = gpta - gptprsets + 1
        nplat = nloen[lats[0]] - ix[0] + 1 + sum(nloen[j] for j in lats[1:])
        ilst[-1] = nloen[lats[-1]] - nplat + gpts
        div = [360000.0 / nloen[j] for j in lats]
        gptsp, irest = gpts // nb, gpts - nb * (gpts // nb)
        for jb in range(nb):
            npts = gptsp + 1 if jb < irest else gptsp
            sta = [0] * len(lats); onl = [0] * len(lats)
            for _ in range(npts):
                best, inx = 360000, -1
                for k in range(len(lats)):
                    if ix[k] <= ilst[k]:
                        a = _nint((ix[k] - 1) * div[k])
                        if a < best:
                            best, inx = a, k
                if sta[inx] == 0:
                    sta[inx] = ix[inx]
                onl[inx] += 1
                ix[inx] += 1
            segs.append([(lats[k], sta[k] - 1, onl[k]) for k in range(len(lats)) if onl[k] > 0])
        gpta += gpts
    return nreg, segs
