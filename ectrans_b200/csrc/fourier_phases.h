// Fourier stage, per field-pair phases.  Two real fields are transformed as one complex
// FFT of length nlon (z = f1 + i f2).  Each phase is a cooperative loop over (tid, nthr);
// kernels separate phases with __syncthreads(), the CPU test harness runs tid = 0..nthr-1.
//
// Inverse (FTINV): reference cpu/internal/fourier_in_mod.F90:58-77 (gather m -> row),
//   fsc_mod.F90:132-187 (1/(a cos theta) scaling, E-W derivatives), ftinv_mod.F90:65-84
//   (zero m > NMEN, unnormalised c2r).
// Direct (FTDIR): ftdir_mod.F90:67-84 (r2c, /N tpm_fftw.F90:317-321), fourier_out_mod.F90:58-77.
#pragma once
#include "fft_plan.h"

struct EctFsField {      // one Fourier-space field of the inverse transform
    int src_c;           // column (2 * Legendre field index) in a Fourier-buffer record; -1 = absent
    int pw;              // power of 1/(a cos theta): 0, 1, 2
    int deriv;           // 1: multiply by i * m * 1/(a cos theta) on top (E-W derivative)
};

struct EctPairCtx {
    // geometry of this latitude
    int nlon, km;
    double racthe;
    // plan data
    const uint16_t* perm;         // perm pool + plan.perm_off
    EctTw qt;                     // two-level twiddle table (shared memory in the kernels)
    const double2* roots;
    const double2* chirp;         // Bluestein only
    const double2* bhat;          // Bluestein only (direction specific)
    int bluestein, m;
    // Fourier buffer
    const int* rec;               // record index for m = 0..km of this latitude
    long long cp;                 // record pitch (doubles)
};

ECT_HD double2 ect_load_fs(const double* __restrict__ fb, const EctPairCtx& c, long long recbase,
                           const EctFsField& f, int m, double s1, double s2) {
    if (f.src_c < 0) return make_double2(0.0, 0.0);
    const double2 v = *reinterpret_cast<const double2*>(fb + recbase + f.src_c);
    double s = f.pw == 0 ? 1.0 : (f.pw == 1 ? s1 : s2);
    double re = v.x * s, im = (m == 0) ? 0.0 : v.y * s;
    if (f.deriv) {
        const double z = s1 * (double)m;
        return make_double2(-im * z, re * z);
    }
    return make_double2(re, im);
}

// ---- inverse, phase 1: gather spectrum of the pair into the work array ----
ECT_HD void ftinv_load(double2* data, const double* __restrict__ fb, const EctPairCtx& c,
                       EctFsField fa, EctFsField fbd, int tid, int nthr) {
    const int N = c.nlon, km = c.km;
    const double s1 = c.racthe, s2 = c.racthe * c.racthe;
    if (!c.bluestein) {
        for (int k = km + 1 + tid; k < N - km; k += nthr) data[ECT_PAD((int)c.perm[k])] = make_double2(0.0, 0.0);
        for (int k = tid; k <= km; k += nthr) {
            const long long rb = (long long)c.rec[k] * c.cp;
            const double2 a = ect_load_fs(fb, c, rb, fa, k, s1, s2);
            const double2 b = ect_load_fs(fb, c, rb, fbd, k, s1, s2);
            data[ECT_PAD((int)c.perm[k])] = make_double2(a.x - b.y, a.y + b.x);
            if (k > 0) data[ECT_PAD((int)c.perm[N - k])] = make_double2(a.x + b.y, b.x - a.y);
        }
    } else {
        for (int u = 2 * km + 1 + tid; u < c.m; u += nthr) data[ECT_PAD(u)] = make_double2(0.0, 0.0);
        for (int k = tid; k <= km; k += nthr) {
            const long long rb = (long long)c.rec[k] * c.cp;
            const double2 a = ect_load_fs(fb, c, rb, fa, k, s1, s2);
            const double2 b = ect_load_fs(fb, c, rb, fbd, k, s1, s2);
            const double2 ch = c.chirp[k];
            const double2 xp = c_mul(make_double2(a.x - b.y, a.y + b.x), ch);
            data[ECT_PAD(km + k)] = make_double2(xp.y, xp.x);          // stored swapped for the sign-(-) FFT
            if (k > 0) {
                const double2 xm = c_mul(make_double2(a.x + b.y, b.x - a.y), ch);
                data[ECT_PAD(km - k)] = make_double2(xm.y, xm.x);
            }
        }
    }
}

// ---- inverse, last phase: write the two real rows ----
// rowa/rowb: pointers to element j = 0 of the row in the (blocked) grid-point array; rows are
// addressed through gp_index() by the caller, here they are plain contiguous segments
ECT_HD double2 ftinv_out(const double2* data, const EctPairCtx& c, int j) {
    if (!c.bluestein) return data[ECT_PAD(j)];
    long long jj = j;
    double sg = 1.0;        // c[N - j] = (-1)^N c[j]
    if (jj > c.nlon / 2) { jj = c.nlon - jj; if (c.nlon & 1) sg = -1.0; }
    const double2 ch = c.chirp[jj];
    return c_mul(make_double2(sg * ch.x, sg * ch.y), data[ECT_PAD(j)]);   // t = k - o0 with o0 = 0
}

// ---- direct, phase 1: load two real rows (swapped: sign - transform on the sign + core) ----
ECT_HD void ftdir_put(double2* data, const EctPairCtx& c, int j, double va, double vb) {
    if (!c.bluestein) {
        data[ECT_PAD((int)c.perm[j])] = make_double2(vb, va);
    } else {
        long long jj = j;
        double sg = 1.0;
        if (jj > c.nlon / 2) { jj = c.nlon - jj; if (c.nlon & 1) sg = -1.0; }
        const double2 a = c_mul(make_double2(sg * vb, sg * va), c.chirp[jj]);
        data[ECT_PAD(j)] = make_double2(a.y, a.x);
    }
}
ECT_HD void ftdir_zero_tail(double2* data, const EctPairCtx& c, int tid, int nthr) {
    if (c.bluestein)
        for (int u = c.nlon + tid; u < c.m; u += nthr) data[ECT_PAD(u)] = make_double2(0.0, 0.0);
}

// Z[k] for k in [-km, km] of the sign-(-) DFT of z = fa + i fb
ECT_HD double2 ftdir_z(const double2* data, const EctPairCtx& c, int k) {
    if (!c.bluestein) {
        const int idx = k >= 0 ? k : c.nlon + k;
        const double2 r = data[ECT_PAD(idx)];
        return make_double2(r.y, r.x);
    }
    const int ak = k >= 0 ? k : -k;
    const double2 x = c_mul(c.chirp[ak], data[ECT_PAD(k + c.km)]);
    return make_double2(x.y, x.x);
}

// ---- direct, last phase: split the pair, normalise by 1/nlon, scatter into records ----
ECT_HD void ftdir_store(const double2* data, double* __restrict__ fb, const EctPairCtx& c,
                        int ca, int cb, int tid, int nthr) {
    const double sc = 0.5 / (double)c.nlon;
    for (int k = tid; k <= c.km; k += nthr) {
        const double2 zk = ftdir_z(data, c, k);
        const double2 zn = (k == 0) ? zk : ftdir_z(data, c, -k);
        const long long rb = (long long)c.rec[k] * c.cp;
        // F1 = (Zk + conj Zn)/2, F2 = -i (Zk - conj Zn)/2
        if (ca >= 0)
            *reinterpret_cast<double2*>(fb + rb + ca) = make_double2((zk.x + zn.x) * sc, (zk.y - zn.y) * sc);
        if (cb >= 0)
            *reinterpret_cast<double2*>(fb + rb + cb) = make_double2((zk.y + zn.y) * sc, (zn.x - zk.x) * sc);
    }
}
