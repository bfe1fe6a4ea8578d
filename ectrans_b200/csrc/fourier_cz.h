// Chirp-z rows of the Fourier stage, split over a CTA pair.
//
// Row lengths with a prime factor > 31 (60 % of the rows of an octahedral grid) are transformed as a convolution of
// length M = 2 H (fft_plan.cu).  One radix-2 step of the length-M transforms is taken analytically, which turns the
// convolution into two INDEPENDENT length-H ones:
//     even bins:  e[u] = x[u] + x[u + H]                         E = IFFT_H(FFT_H(e) . Bhat[2r])
//     odd bins :  o[u] = (x[u] - x[u + H]) w^-u                  O = IFFT_H(FFT_H(o) . Bhat[2r+1])
//     y[j] = E[j] + w^j O[j],  y[j + H] = E[j] - w^j O[j],       w = exp(2 pi i / M)
// CTA 0 of a cluster works on E, CTA 1 on O, each in a work array of H elements (half the shared memory, so that
// three CTAs fit one SM for the longest dp rows of TCo1279, where the undivided array allowed one); the halves meet
// only in the output phase, through distributed shared memory.  The zero padding of the convolution makes the
// radix-2 step free: the inverse transform has x[u + H] = 0, the direct one needs y[j] for j < H only.
//
// The phases are __host__ __device__ cooperative loops (tid, nthr) like fourier_phases.h: the kernels separate them
// with __syncthreads() / cluster barriers, tests/hostemu runs them on the CPU.  Global-memory access goes through
// functors so that the same arithmetic serves the kernel (batched loads) and the harness.
//
// Same arithmetic conventions as fourier.cu / fourier_phases.h: the forward (sign -) transform runs on the sign-+
// core on swapped (re <-> im) data; the middle step un-swaps.
#pragma once
#include "fourier_phases.h"

template <typename C>
struct CzCtx {
    int N, km, H;              // row length, NMEN of the row, half convolution length (M = 2 H)
    int half;                  // 0: even bins (E), 1: odd bins (O)
    EctTwT<C> twm;             // two-level table of exp(2 pi i j / M)
    EctTwT<C> twc;             // two-level table of exp(2 pi i t / (2 N)): the chirp c[j] = exp(i pi j^2 / N) is entry j^2 mod 2N
    unsigned n2, magic;        // 2 N and ceil(2^32 / (2 N))
};

template <typename C>
ECT_HD C cz_conj(C a) { return c_make<C>(a.x, -a.y); }
template <typename C>
ECT_HD C cz_swap(C a) { return c_make<C>(a.y, a.x); }

// c[j] = exp(i pi j^2 / N) for 0 <= j < N out of shared memory (no global-memory latency in the load / output
// phases): t = j^2 mod 2N by a multiply-high (j^2 < 2^27 for rows up to 11585 points), then two table entries
template <typename C>
ECT_HD C cz_chirp(const CzCtx<C>& c, int j) {
    const unsigned x = (unsigned)j * (unsigned)j;
    const unsigned q = (unsigned)(((unsigned long long)x * c.magic) >> 32);     // floor(x / 2N) or one more
    int r = (int)x - (int)(q * c.n2);
    if (r < 0) r += (int)c.n2;
    return tw_get(c.twc, r);
}

#define CZ_NB 4     // independent global / distributed-shared loads in flight per thread (x 2 .. 4 arrays)
#define CZ_NB8 8

// ---- inverse, load: FOURIER_IN + FSC (fsc_mod.F90:132-187) + chirp + the analytic radix-2 step ----
// loadrec(k, ra, rb): the two double2 spectral values (fields a, b) of wavenumber k of this latitude
struct CzInvScale {            // per pair
    int pwa, deriva, pwb, derivb, hasb;
    double s1, s2;             // 1 / (a cos theta) and its square
    double rowscale;           // 1, or w / N for DIR_TRANSAD
};
template <typename C, typename LoadRec>
ECT_HD void cz_inv_load(C* data, const CzCtx<C>& c, const CzInvScale& f, LoadRec loadrec, int tid, int nthr) {
    typedef typename EctReal<C>::type R_;
    const int km = c.km;
    for (int u = 2 * km + 1 + tid; u < c.H; u += nthr) data[ECT_PAD(u)] = c_make<C>(0, 0);
    const R_ s1 = (R_)f.s1;
    const R_ sa_ = (R_)((f.pwa == 0 ? 1.0 : (f.pwa == 1 ? f.s1 : f.s2)) * f.rowscale);
    const R_ sb_ = (R_)((f.pwb == 0 ? 1.0 : (f.pwb == 1 ? f.s1 : f.s2)) * f.rowscale);
    for (int k0 = tid; k0 <= km; k0 += CZ_NB8 * nthr) {
        double2 ra[CZ_NB8], rb[CZ_NB8];
#pragma unroll
        for (int i = 0; i < CZ_NB8; ++i) {
            const int k = k0 + i * nthr;
            ra[i] = rb[i] = make_double2(0.0, 0.0);
            if (k <= km) loadrec(k, ra[i], rb[i]);
        }
#pragma unroll
        for (int i = 0; i < CZ_NB8; ++i) {
            const int k = k0 + i * nthr;
            if (k > km) continue;
            const C va = c_cvt<C>(ra[i]), vb = c_cvt<C>(rb[i]);
            C fa_ = c_make<C>(va.x * sa_, k == 0 ? (R_)0 : va.y * sa_);
            C fb_ = c_make<C>(vb.x * sb_, k == 0 ? (R_)0 : vb.y * sb_);
            const R_ z = s1 * (R_)k;
            if (f.deriva) fa_ = c_make<C>(-fa_.y * z, fa_.x * z);
            if (f.derivb) fb_ = c_make<C>(-fb_.y * z, fb_.x * z);
            const C zp = c_make<C>(fa_.x - fb_.y, fa_.y + fb_.x);      // Z[k]  of z = fa + i fb
            const C zm = c_make<C>(fa_.x + fb_.y, fb_.x - fa_.y);      // Z[-k]
            const C chk = cz_chirp(c, k);
            C xp = c_mul(zp, chk);
            if (c.half) xp = c_mul(xp, cz_conj(tw_get(c.twm, km + k)));
            data[ECT_PAD(km + k)] = cz_swap(xp);
            if (k > 0) {
                C xm = c_mul(zm, chk);
                if (c.half) xm = c_mul(xm, cz_conj(tw_get(c.twm, km - k)));
                data[ECT_PAD(km - k)] = cz_swap(xm);
            }
        }
    }
}

// ---- inverse, output: y = E +- w^j O, times the chirp; store(j, y): y.x -> field a, y.y -> field b at longitude j ----
// mine / other: result array of this CTA / of the partner (distributed shared memory in the kernel)
template <typename C, typename Store>
ECT_HD void cz_inv_out(const C* mine, const C* other, const CzCtx<C>& c, Store store, int tid, int nthr) {
    const int Jn = c.H < c.N ? c.H : c.N, Jh = (Jn + 1) >> 1;
    const int j_lo = c.half * Jh, j_hi = (j_lo + Jh < Jn) ? j_lo + Jh : Jn;
    for (int j0 = j_lo + tid; j0 < j_hi; j0 += CZ_NB8 * nthr) {
        C m[CZ_NB8], o[CZ_NB8];
#pragma unroll
        for (int i = 0; i < CZ_NB8; ++i) {
            const int j = j0 + i * nthr;
            if (j < j_hi) {
                o[i] = other[ECT_PAD(j)];
                m[i] = mine[ECT_PAD(j)];
            }
        }
#pragma unroll
        for (int i = 0; i < CZ_NB8; ++i) {
            const int j = j0 + i * nthr;
            if (j >= j_hi) continue;
            const C E = c.half ? o[i] : m[i], O = c.half ? m[i] : o[i];
            const C t = c_mul(tw_get(c.twm, j), O);
            store(j, c_mul(cz_chirp(c, j), c_add(E, t)));
            if (j + c.H < c.N) store(j + c.H, c_mul(cz_chirp(c, j + c.H), c_sub(E, t)));
        }
    }
}

// ---- direct, load: two real rows, chirp, the radix-2 step ----
// loadgp(j, va, vb): grid-point values of fields a, b at longitude j < N
template <typename C, typename LoadGp>
ECT_HD void cz_dir_load(C* data, const CzCtx<C>& c, LoadGp loadgp, int tid, int nthr) {
    typedef typename EctReal<C>::type R_;
    for (int j0 = tid; j0 < c.H; j0 += CZ_NB * nthr) {
        R_ a0[CZ_NB], b0[CZ_NB], a1[CZ_NB], b1[CZ_NB];
#pragma unroll
        for (int i = 0; i < CZ_NB; ++i) {
            const int j = j0 + i * nthr;
            a0[i] = b0[i] = a1[i] = b1[i] = 0;
            if (j < c.H) {
                if (j < c.N) loadgp(j, a0[i], b0[i]);
                if (j + c.H < c.N) loadgp(j + c.H, a1[i], b1[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < CZ_NB; ++i) {
            const int j = j0 + i * nthr;
            if (j >= c.H) continue;
            const C t0 = j < c.N ? c_mul(c_make<C>(b0[i], a0[i]), cz_chirp(c, j)) : c_make<C>(0, 0);
            const C t1 = j + c.H < c.N ? c_mul(c_make<C>(b1[i], a1[i]), cz_chirp(c, j + c.H)) : c_make<C>(0, 0);
            C x;
            if (c.half) x = c_mul(c_sub(t0, t1), cz_conj(tw_get(c.twm, j)));
            else x = c_add(t0, t1);
            data[ECT_PAD(j)] = cz_swap(x);
        }
    }
}

// ---- direct, output: Z[k], Z[-k] of the pair for this CTA's share of k = 0 .. km; storerec(k, Zk, Zn) ----
template <typename C, typename StoreRec>
ECT_HD void cz_dir_out(const C* mine, const C* other, const CzCtx<C>& c, StoreRec storerec, int tid, int nthr) {
    const int K = c.km + 1, Kh = (K + 1) >> 1;
    const int k_lo = c.half * Kh, k_hi = (k_lo + Kh < K) ? k_lo + Kh : K;
    for (int k0 = k_lo + tid; k0 < k_hi; k0 += CZ_NB * nthr) {
        C mp[CZ_NB], mm[CZ_NB], op[CZ_NB], om[CZ_NB];
#pragma unroll
        for (int i = 0; i < CZ_NB; ++i) {
            const int k = k0 + i * nthr;
            if (k < k_hi) {
                op[i] = other[ECT_PAD(c.km + k)]; om[i] = other[ECT_PAD(c.km - k)];
                mp[i] = mine[ECT_PAD(c.km + k)]; mm[i] = mine[ECT_PAD(c.km - k)];
            }
        }
#pragma unroll
        for (int i = 0; i < CZ_NB; ++i) {
            const int k = k0 + i * nthr;
            if (k >= k_hi) continue;
            const C Ep = c.half ? op[i] : mp[i], Op = c.half ? mp[i] : op[i];
            const C Em = c.half ? om[i] : mm[i], Om = c.half ? mm[i] : om[i];
            const C yp = c_add(Ep, c_mul(tw_get(c.twm, c.km + k), Op));
            const C ym = c_add(Em, c_mul(tw_get(c.twm, c.km - k), Om));
            // stored values are swapped (sign - transform on the sign + core): Z = (y, x)
            const C chk = cz_chirp(c, k);
            storerec(k, cz_swap(c_mul(chk, yp)), cz_swap(c_mul(chk, ym)));
        }
    }
}
