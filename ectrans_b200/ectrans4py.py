"""ectrans4py face (reference src/ectrans4py/__init__.py) on the B200 library.

Same function names, argument order and return values as the reference's global (Gaussian-grid)
entry points: ``trans_inq4py`` (:164), ``sp2gp_gauss4py`` (:305), ``gp2sp_gauss4py`` (:364),
``get_legendre_assets`` (:89).  Single task, double precision, one field per call, handles cached per
(truncation, nlat, nloen) like spec_setup4py.F90:100-159.  LREORDER=True (the ARPEGE 'model' coefficient order,
sp2gp_gauss4py.F90:82-108, gp2sp_gauss4py.F90:92-106) is a permutation on the host.  LAM entry points are out of scope.
"""
from __future__ import annotations

import numpy as np

from . import Transform

_handles = {}


def init_env(*args, **kwargs):
    """Accepted for source compatibility (the reference sets OMP/stack limits here)."""
    return None


def _get(ktrunc, ksizej, kloen, knummaxresol):
    key = (int(ktrunc), int(ksizej), hash(np.asarray(kloen, dtype=np.int64).tobytes()))
    if key not in _handles:
        if len(_handles) >= int(knummaxresol):
            raise RuntimeError("ectrans4py: KNUMMAXRESOL handles exceeded")
        _handles[key] = Transform(int(ktrunc), np.asarray(kloen, dtype=np.int32)[:int(ksizej)])
    return _handles[key]


def _model_order(T):
    """Indices (0-based) of the 'model' coefficient order used with LREORDER=True: for total wavenumber n the value of
    (m, n) sits at NASM0(n) + m (real part) and NASM0(n) - m (imaginary part, m > 0), NASM0(0) = 1, NASM0(n+1) =
    NASM0(n) + 2 n + 2 (sp2gp_gauss4py.F90:84-107).  Returns (idx_re, idx_im) in ecTrans order (m-major, n ascending);
    idx_im is -1 for m = 0."""
    nasm0 = np.zeros(T + 1, dtype=np.int64)
    ji = 1
    for n in range(T + 1):
        nasm0[n] = ji
        ji += 2 * n + 2
    re, im = [], []
    for m in range(T + 1):
        for n in range(m, T + 1):
            re.append(nasm0[n] + m - 1)
            im.append(nasm0[n] - m - 1 if m else -1)
    return np.asarray(re), np.asarray(im)


def _from_model_order(T, pspec):
    re, im = _model_order(T)
    out = np.zeros(2 * re.size)
    out[0::2] = pspec[re]
    out[1::2] = np.where(im >= 0, pspec[np.maximum(im, 0)], 0.0)
    return out


def _to_model_order(T, native, size):
    re, im = _model_order(T)
    out = np.zeros(size)
    out[re] = native[0::2]
    out[im[im >= 0]] = native[1::2][im >= 0]
    return out


def trans_inq4py(KSIZEJ, KTRUNC, KSLOEN, KLOEN, KNUMMAXRESOL):
    """Returns (KGPTOT, KSPEC, KNMENG)."""
    t = _get(KTRUNC, KSIZEJ, KLOEN, KNUMMAXRESOL)
    return t.ngptot, t.nspec2 // 2, t.nmen.astype(np.int64)


def sp2gp_gauss4py(KSIZEJ, KTRUNC, KNUMMAXRESOL, KGPTOT, KSLOEN, KLOEN, KSIZE, LGRADIENT, LREORDER, PSPEC):
    """Returns (PGPT, PGPTM, PGPTL): field, N-S derivative, E-W derivative (zeros unless LGRADIENT)."""
    t = _get(KTRUNC, KSIZEJ, KLOEN, KNUMMAXRESOL)
    PSPEC = np.asarray(PSPEC, dtype=np.float64)
    if LREORDER:
        PSPEC = _from_model_order(int(KTRUNC), PSPEC)
    sp = np.ascontiguousarray(PSPEC, dtype=np.float64).reshape(-1, 1)
    assert sp.shape[0] == t.nspec2 == KSIZE and KGPTOT == t.ngptot
    gp = t.inv_trans(spscalar=sp, scders=bool(LGRADIENT))
    z = np.zeros(t.ngptot)
    if LGRADIENT:
        return gp[0, 0].copy(), gp[0, 1].copy(), gp[0, 2].copy()
    return gp[0, 0].copy(), z, z.copy()


def gp2sp_gauss4py(KSPEC, KSIZEJ, KTRUNC, KNUMMAXRESOL, KSLOEN, KLOEN, KSIZE, LREORDER, PGPT):
    """Returns PSPEC (KSPEC real values: (re, im) pairs, m-major, n ascending)."""
    t = _get(KTRUNC, KSIZEJ, KLOEN, KNUMMAXRESOL)
    assert KSPEC == t.nspec2 and KSIZE == t.ngptot
    gp = np.ascontiguousarray(PGPT, dtype=np.float64).reshape(1, 1, -1)
    sp = t.dir_trans(gp, 0, 1)[2][:, 0].copy()
    return _to_model_order(int(KTRUNC), sp, int(KSPEC)) if LREORDER else sp


def get_legendre_assets(KSIZEJ, KTRUNC, KSLOEN, KSPOLEGL, KLOEN, KNUMMAXRESOL):
    """Returns (KNMENG, PGW, PRPNM[KSLOEN//2, KSPOLEGL]).  PRPNM follows the F%RPNM order
    (per m, n descending from T+1 to m, setup_dims_mod.F90:28-38); entries of latitudes that do not
    carry m on the reduced grid (not resident in HBM) are returned as 0."""
    t = _get(KTRUNC, KSIZEJ, KLOEN, KNUMMAXRESOL)
    rpnm, _ = t.legendre_polynomials()                    # TRANS_INQ(PRPNM): (nspolegl, ndgnh)
    if rpnm.shape[0] != int(KSPOLEGL):
        raise RuntimeError("get_legendre_assets: KSPOLEGL must be sum(T + 2 - m), m = 0 .. T")
    return t.nmen.astype(np.int64), t.rgw.copy(), np.ascontiguousarray(rpnm.T)
