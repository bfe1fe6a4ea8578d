// Normalised associated Legendre polynomials for one (m, mu) column, one parity of n-m.
// Follows SUPOLF (reference common/internal/supolf_mod.F90:85-247) including its 1e+-100
// rescaling bookkeeping and the EPSILON clamp on un-scaling (:236-245), with the INI_POL
// constants of common/internal/tpm_pol.F90:74-81.  __host__ __device__: the table kernel calls it
// per (m, latitude) thread; tests/hostemu runs it on the CPU against the oracle.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#ifndef ECT_HD
#define ECT_HD __host__ __device__ __forceinline__
#endif

struct EctSupolfM {      // per-m constants evaluated on the host in the reference's operation order
    double f0, f1, f3;   // ZFAC after IC = 0, 1, 3
    double sq[4];        // SQRT(2 (m + ic + 1/2) ic! / prod_{j<=ic}(2m + j))
};

inline void ect_supolf_consts(int km, EctSupolfM& c) {
    double zfac = 1.0;
    for (int jn = 1; jn <= km - 1; ++jn) {
        zfac = zfac * sqrt((double)(2 * jn - 1));
        zfac = zfac / sqrt((double)(2 * jn));
    }
    zfac = zfac * sqrt((double)(2 * km - 1));
    c.f0 = zfac;
    c.f1 = c.f0 * (double)(2 * km + 1);
    c.f3 = c.f1 * (double)(2 * km + 3);
    double zfac0 = 1.0;
    const double zfac1[4] = {1.0, 1.0, 2.0, 6.0};
    for (int ic = 0; ic < 4; ++ic) {
        zfac0 = zfac0 * (double)(2 * km + ic);
        c.sq[ic] = sqrt(2.0 * ((double)(km + ic) + 0.5) * zfac1[ic] / zfac0);
    }
}

ECT_HD double ect_sup_undo(double v, int icorr) {
    const double zscale = 1.0e100, zeps = 2.220446049250313e-16;
    for (int j = 0; j < icorr; ++j) {
        v = v / zscale;
        if (v < zeps) v = zeps;
    }
    return v;
}

// Writes out[k * ld] = P_n^m(mu), n = km + par + 2k, k = 0 .. kcount-1.
// knsmax: upper bound of the recurrence loop (INMAX in cpu/internal/suleg_mod.F90:646-650, :928-932).
ECT_HD void ect_supolf_column(int km, int par, int kcount, int knsmax, double mu, const EctSupolfM& cm,
                              double* out, long long ld) {
    const double zeps = 2.220446049250313e-16;
    double dlx = mu;
    double zcos2 = 1.0 - dlx * dlx;
    double zcos = sqrt(zcos2), zcos_r;
    if (fabs(zcos) <= zeps) { dlx = 1.0; zcos = 0.0; zcos_r = 0.0; zcos2 = 0.0; }
    else zcos_r = 1.0 / zcos;
    if (km <= 1) {
        // ordinary Legendre recurrence, supolf_mod.F90:130-151
        double dlkm2 = 1.0, dlkm1 = dlx;
        for (int jn = 0; jn <= km + par + 2 * (kcount - 1); ++jn) {
            double val;
            if (jn == 0) val = (km == 0) ? 1.0 : 0.0;
            else {
                const double dfb = sqrt((double)(2 * jn + 1) / (double)(jn * (jn + 1)));
                if (jn == 1) {
                    val = (km == 0) ? dlkm1 * dfb / (1.0 / sqrt((double)(jn * (jn + 1)))) : zcos * dfb;
                } else {
                    const double dlk = ((double)(2 * jn - 1) / (double)jn) * dlx * dlkm1 -
                                       ((double)(jn - 1) / (double)jn) * dlkm2;
                    if (km == 0) val = dlk * dfb / (1.0 / sqrt((double)(jn * (jn + 1))));
                    else val = ((double)jn * (dlkm1 - dlx * dlk) * zcos_r) * dfb;
                    dlkm2 = dlkm1;
                    dlkm1 = dlk;
                }
            }
            const int r = jn - km - par;
            if (r >= 0 && (r & 1) == 0) out[(long long)(r >> 1) * ld] = val;
        }
        return;
    }
    const double zscale = 1.0e100, ziscale = 1.0e-100;
    double zlsita = 1.0;
    int corr = 0;
    for (int jn = 1; jn <= km / 2; ++jn) {
        zlsita = zlsita * zcos2;
        if (fabs(zlsita) < ziscale) { zlsita = zlsita * zscale; ++corr; }
    }
    if (km & 1) zlsita = zlsita * zcos;
    // explicit first two values of this parity (ic = par, par + 2), supolf_mod.F90:183-212
    double zm0, zm1;
    if (par == 0) {
        zm0 = cm.f0;
        zm1 = 0.5 * cm.f1 * ((double)(2 * km + 3) * dlx * dlx - 1.0);
    } else {
        zm0 = cm.f1 * dlx;
        zm1 = (1.0 / 6.0) * dlx * cm.f3 * ((double)(2 * km + 5) * dlx * dlx - 3.0);
    }
    double a = zlsita * zm0 * cm.sq[par];        // P[n0]
    double b = zlsita * zm1 * cm.sq[par + 2];    // P[n0 + 2]
    const int n0 = km + par;
    const int nlast = n0 + 2 * (kcount - 1);
    const double x2 = dlx * dlx;
    const double dkm2 = (double)km * (double)km;
    int emitted = 0;                              // entries written so far (k index)
    for (int jn = n0 + 4; jn <= knsmax; jn += 2) {
        if (fabs(a) > zscale) { a = a / zscale; b = b / zscale; --corr; }
        if (emitted < kcount) out[(long long)emitted * ld] = ect_sup_undo(a, corr);
        ++emitted;
        // DCL(k), DDL(k): supolf_mod.F90:85-89
        const double k2 = (double)(jn - 2), k4 = (double)(jn - 4);
        const double dcl4 = sqrt(((k4 - km + 1.0) * (k4 - km + 2.0) * (k4 + km + 1.0) * (k4 + km + 2.0)) /
                                 ((2.0 * k4 + 1.0) * (2.0 * k4 + 3.0) * (2.0 * k4 + 3.0) * (2.0 * k4 + 5.0)));
        const double dcl2 = sqrt(((k2 - km + 1.0) * (k2 - km + 2.0) * (k2 + km + 1.0) * (k2 + km + 2.0)) /
                                 ((2.0 * k2 + 1.0) * (2.0 * k2 + 3.0) * (2.0 * k2 + 3.0) * (2.0 * k2 + 5.0)));
        const double ddl2 = (2.0 * k2 * (k2 + 1.0) - 2.0 * dkm2 - 1.0) / ((2.0 * k2 - 1.0) * (2.0 * k2 + 3.0));
        const double c = ((x2 - ddl2) * b - dcl4 * a) / dcl2;
        a = b;
        b = c;
    }
    if (emitted < kcount) out[(long long)emitted * ld] = ect_sup_undo(a, corr);
    ++emitted;
    if (emitted < kcount) out[(long long)emitted * ld] = ect_sup_undo(b, corr);
    (void)nlast;
}
