// In-shared-memory mixed-radix complex FFT core (sign +, unnormalised) used by the
// Fourier stage (FTINV / FTDIR, reference cpu/internal/ftinv_mod.F90:65-84,
// ftdir_mod.F90:67-84; the reference calls FFTW there, tpm_fftw.F90:251-377).
//
// Everything here is __host__ __device__ so the index logic can be exercised on the
// CPU by tests/ (tests/hostemu) -- the product only ever calls it from CUDA kernels.
//
// Data: double2 array of length n processed in place; element i lives at ECT_PAD(i)
// (one pad element every 16) so that the stride-R accesses of the innermost passes do not
// pile onto the same shared-memory banks.
//   DIT:  input at digit-reversed positions (perm table), output natural order.
//   DIF:  input natural order, output at digit-reversed positions.
// Stage s (0 = innermost) has radix r_s and sub-length L_s = prod_{j<s} r_j.
// Twiddles: one table lookup per butterfly (w = exp(2 pi i k / (R L))), its powers by
// multiplication -- the profile of the first version showed the shared-memory pipe, not the FP64
// pipe, to be the limiter (profiles/r01_ncu_full_summary.txt).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define ECT_HD __host__ __device__ __forceinline__
#define ECT_MAX_STAGES 14
#define ECT_MAX_RADIX 31
#define ECT_PAD(i) ((i) + ((i) >> 4))
#define ECT_PADDED_LEN(n) ((n) + ((n) >> 4) + 1)

struct EctFftPlan {        // POD; one per transform length
    int n;                 // length (even)
    int nst;               // number of stages
    int quarter;           // 1: n % 4 == 0, twiddle table holds exp(2 pi i j/n) for j <= n/4; 0: j < n/2
    int radix[ECT_MAX_STAGES];
    int sublen[ECT_MAX_STAGES];   // L_s
    int lshift[ECT_MAX_STAGES];   // log2(L_s) if L_s is a power of two, else -1
    int perm_off;          // offset (elements) into the uint16 permutation pool: pos(i), i < n
    int tw_off;            // offset (double2) into the twiddle pool
    int tw_len;            // entries of the twiddle table
};

ECT_HD double2 c_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
ECT_HD double2 c_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
ECT_HD double2 c_mul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
ECT_HD double2 c_muli(double2 a) { return make_double2(-a.y, a.x); }   // a * i
// single-precision arithmetic (sp handles): the same core on float2
ECT_HD float2 c_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
ECT_HD float2 c_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
ECT_HD float2 c_mul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
ECT_HD float2 c_muli(float2 a) { return make_float2(-a.y, a.x); }
template <typename C> struct EctReal;
template <> struct EctReal<double2> { typedef double type; };
template <> struct EctReal<float2> { typedef float type; };
template <typename C, typename A, typename B>
ECT_HD C c_make(A x, B y) { C c; c.x = (typename EctReal<C>::type)x; c.y = (typename EctReal<C>::type)y; return c; }
template <typename C, typename S>
ECT_HD C c_cvt(S a) { return c_make<C>(a.x, a.y); }

// exp(+2 pi i j / n) for 0 <= j < n from the quarter table qt[0 .. n/4] (n4 = n/4 > 0)
// or from the half table qt[0 .. n/2) (n4 = -(n/2))
ECT_HD double2 tw_lookup(const double2* __restrict__ qt, int j, int n4) {
    if (n4 < 0) {
        const int nh = -n4;
        if (j >= nh) { const double2 w = qt[j - nh]; return make_double2(-w.x, -w.y); }
        return qt[j];
    }
    int q = 0;
    if (j >= 2 * n4) { j -= 2 * n4; q = 2; }
    if (j >= n4) { j -= n4; q += 1; }
    double2 w = qt[j];
    if (q == 1) return make_double2(-w.y, w.x);
    if (q == 2) return make_double2(-w.x, -w.y);
    if (q == 3) return make_double2(w.y, -w.x);
    return w;
}

// Two-level twiddle table (lives in shared memory in the kernels): exp(2 pi i j / n) = t1[j >> 7] * t2[j & 127]
template <typename C>
struct EctTwT {
    const C* t1;           // exp(2 pi i 128 a / n), a <= n / 128
    const C* t2;           // exp(2 pi i b / n), b < 128
    int sh = 0;            // index shift: a table of length n 2^sh serves transforms of length n (split chirp-z: M = 2 H)
};
typedef EctTwT<double2> EctTw;
#define ECT_TW1_LEN(n) ((n) / 128 + 2)
#define ECT_TW2_LEN 128
template <typename C>
ECT_HD C tw_get(const EctTwT<C>& t, int j) { j <<= t.sh; return c_mul(t.t1[j >> 7], t.t2[j & 127]); }
// fills the tables cooperatively from the per-length table qt (quarter / half wave, see tw_lookup)
template <typename C>
ECT_HD void tw_build(C* t1, C* t2, const double2* __restrict__ qt, int n, int tid, int nthr) {
    const int n4 = (n & 3) ? -(n >> 1) : (n >> 2);
    for (int a = tid; a < ECT_TW1_LEN(n); a += nthr) t1[a] = c_cvt<C>(tw_lookup(qt, (128 * a) % n, n4));
    for (int b = tid; b < ECT_TW2_LEN; b += nthr) t2[b] = c_cvt<C>(tw_lookup(qt, b % n, n4));
}

// ---- butterflies: u[p] = sum_q v[q] exp(+2 pi i p q / R), in place on v[0..R) ----
// C = double2 (dp handles) or float2 (sp handles)
template <typename C>
ECT_HD void bfly2(C* v) {
    C a = v[0], b = v[1];
    v[0] = c_add(a, b);
    v[1] = c_sub(a, b);
}
template <typename C>
ECT_HD void bfly4(C& v0, C& v1, C& v2, C& v3) {
    C t0 = c_add(v0, v2), t1 = c_sub(v0, v2);
    C t2 = c_add(v1, v3), t3 = c_muli(c_sub(v1, v3));
    v0 = c_add(t0, t2);
    v1 = c_add(t1, t3);
    v2 = c_sub(t0, t2);
    v3 = c_sub(t1, t3);
}
#define ECT_SQH 0.70710678118654752440
#define ECT_C8 0.92387953251128675613
#define ECT_S8 0.38268343236508977173
// radix 8 = 4 x 2: q = 2a + b, p = p1 + 4 p2
template <typename C>
ECT_HD void bfly8(C* v) {
    typedef typename EctReal<C>::type R_;
    const R_ sqh = (R_)ECT_SQH;
    bfly4(v[0], v[2], v[4], v[6]);          // b = 0: y[0][p1] in v[0], v[2], v[4], v[6]
    bfly4(v[1], v[3], v[5], v[7]);          // b = 1
    // twiddle y[1][p1] by W8^p1
    v[3] = c_make<C>((v[3].x - v[3].y) * sqh, (v[3].x + v[3].y) * sqh);
    v[5] = c_muli(v[5]);
    v[7] = c_make<C>((-v[7].x - v[7].y) * sqh, (v[7].x - v[7].y) * sqh);
    // 2-point DFTs over b: X[p1] = y0 + y1, X[p1 + 4] = y0 - y1
    C x0 = c_add(v[0], v[1]), x4 = c_sub(v[0], v[1]);
    C x1 = c_add(v[2], v[3]), x5 = c_sub(v[2], v[3]);
    C x2 = c_add(v[4], v[5]), x6 = c_sub(v[4], v[5]);
    C x3 = c_add(v[6], v[7]), x7 = c_sub(v[6], v[7]);
    v[0] = x0; v[1] = x1; v[2] = x2; v[3] = x3; v[4] = x4; v[5] = x5; v[6] = x6; v[7] = x7;
}
// radix 16 = 4 x 4: q = 4a + b, p = p1 + 4 p2
template <typename C>
ECT_HD void bfly16(C* v) {
#pragma unroll
    for (int b = 0; b < 4; ++b) bfly4(v[b], v[4 + b], v[8 + b], v[12 + b]);   // y[b][p1] at v[4 p1 + b]
    // twiddles W16^(b p1)
    const C w1 = c_make<C>(ECT_C8, ECT_S8), w2 = c_make<C>(ECT_SQH, ECT_SQH), w3 = c_make<C>(ECT_S8, ECT_C8);
    v[5] = c_mul(v[5], w1);                                     // b=1,p1=1
    v[6] = c_mul(v[6], w2);                                     // b=2,p1=1
    v[7] = c_mul(v[7], w3);                                     // b=3,p1=1
    v[9] = c_mul(v[9], w2);                                     // b=1,p1=2
    v[10] = c_muli(v[10]);                                      // b=2,p1=2 : W16^4 = i
    v[11] = c_mul(v[11], c_make<C>(-ECT_SQH, ECT_SQH));         // b=3,p1=2 : W16^6
    v[13] = c_mul(v[13], w3);                                   // b=1,p1=3
    v[14] = c_mul(v[14], c_make<C>(-ECT_SQH, ECT_SQH));         // b=2,p1=3 : W16^6
    v[15] = c_mul(v[15], c_make<C>(-ECT_C8, -ECT_S8));          // b=3,p1=3 : W16^9
#pragma unroll
    for (int p1 = 0; p1 < 4; ++p1) bfly4(v[4 * p1], v[4 * p1 + 1], v[4 * p1 + 2], v[4 * p1 + 3]);   // X[p1 + 4 p2] at v[4 p1 + p2]
    // transpose 4x4 to natural order
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) { C t = v[4 * a + b]; v[4 * a + b] = v[4 * b + a]; v[4 * b + a] = t; }
}

// odd radix R: rt[j] = (cos, sin)(2 pi j / R), j < R.  Outputs are handed to emit(p, value) as they are
// produced so that only the R inputs live in registers.
template <int R, typename C, typename Emit>
ECT_HD void bfly_odd(C* v, const C* __restrict__ rt, Emit emit) {
    typedef typename EctReal<C>::type R_;
    constexpr int H = (R - 1) / 2;
    // in place: v[q] <- v[q] + v[R-q] (t_q), v[R-q] <- v[q] - v[R-q] (d_q)
    C s0 = v[0];
#pragma unroll
    for (int q = 1; q <= H; ++q) {
        const C a = v[q], b = v[R - q];
        v[q] = c_add(a, b);
        v[R - q] = c_sub(a, b);
        s0 = c_add(s0, v[q]);
    }
    emit(0, s0);
#pragma unroll
    for (int p = 1; p <= H; ++p) {
        R_ mx = v[0].x, my = v[0].y, nx = 0, ny = 0;
#pragma unroll
        for (int q = 1; q <= H; ++q) {
            const C w = rt[(p * q) % R];
            mx += w.x * v[q].x;
            my += w.x * v[q].y;
            nx += w.y * v[R - q].x;
            ny += w.y * v[R - q].y;
        }
        // n = i * (nx + i ny) = (-ny, nx)
        emit(p, c_make<C>(mx - ny, my + nx));
        emit(R - p, c_make<C>(mx + ny, my - nx));
    }
}

template <int R, typename C>
ECT_HD void bfly_pow2(C* v) {
    if constexpr (R == 2) bfly2(v);
    else if constexpr (R == 4) bfly4(v[0], v[1], v[2], v[3]);
    else if constexpr (R == 8) bfly8(v);
    else bfly16(v);
}

// rt_all: concatenated root tables; the table of odd radix R occupies [R(R-1)/2, R(R+1)/2)
#define ECT_ROOTS_OFF(R) ((R) * ((R) - 1) / 2)
#define ECT_ROOTS_SIZE (ECT_MAX_RADIX * (ECT_MAX_RADIX + 1) / 2)

// Composite radix R = RO * RP (RO odd in {3, 5, 7}, RP in {2, 4}: 6, 10, 12, 14) by the prime-factor map, all in
// registers and without internal twiddles: with q = (RP a + RO b) mod R and p = CRT(p1, p2) (p1 = p mod RO,
// p2 = p mod RP), exp(2 pi i p q / R) = exp(2 pi i p1 a / RO) exp(2 pi i p2 b / RP).  One pass of radix 14 where the
// plain plan needs a radix-7 and a radix-2 pass over shared memory.
template <int RO, int RP, typename C>
ECT_HD void bfly_pfa(C* v, const C* __restrict__ rt_all) {
    constexpr int R = RO * RP;
    const C* rt = rt_all + ECT_ROOTS_OFF(RO);
#pragma unroll
    for (int b = 0; b < RP; ++b) {
        C u[RO];
#pragma unroll
        for (int a = 0; a < RO; ++a) u[a] = v[(RP * a + RO * b) % R];
        bfly_odd<RO>(u, rt, [&](int p1, C val) { v[(RP * p1 + RO * b) % R] = val; });
    }
#pragma unroll
    for (int p1 = 0; p1 < RO; ++p1) {
        if constexpr (RP == 2) {
            const C x = v[(RP * p1) % R], y = v[(RP * p1 + RO) % R];
            v[(RP * p1) % R] = c_add(x, y);
            v[(RP * p1 + RO) % R] = c_sub(x, y);
        } else {
            bfly4(v[(RP * p1) % R], v[(RP * p1 + RO) % R], v[(RP * p1 + 2 * RO) % R], v[(RP * p1 + 3 * RO) % R]);
        }
    }
    C t[R];
#pragma unroll
    for (int p = 0; p < R; ++p) t[p] = v[(RP * (p % RO) + RO * (p % RP)) % R];
#pragma unroll
    for (int p = 0; p < R; ++p) v[p] = t[p];
}

// register-resident butterflies: powers of two and the composite radices
#define ECT_REG_RADIX(R) ((R) == 2 || (R) == 4 || (R) == 8 || (R) == 16 || (R) == 6 || (R) == 10 || (R) == 12 || (R) == 14)
template <int R, typename C>
ECT_HD void bfly_reg(C* v, const C* __restrict__ rt_all) {
    if constexpr (R == 6) bfly_pfa<3, 2>(v, rt_all);
    else if constexpr (R == 10) bfly_pfa<5, 2>(v, rt_all);
    else if constexpr (R == 12) bfly_pfa<3, 4>(v, rt_all);
    else if constexpr (R == 14) bfly_pfa<7, 2>(v, rt_all);
    else bfly_pow2<R>(v);
}

// v[q] *= w1^q, q = 1 .. R-1.  Powers 1..7 come from a shallow product tree, the rest as
// w1^(8c) * w1^(q mod 8), so that only eight twiddles are live at a time.
template <int R, typename C>
ECT_HD void apply_twiddles(C* v, C w1) {
    constexpr int T = R < 8 ? R : 8;
    C w[T];
    w[0] = c_make<C>(1, 0);
    w[1] = w1;
#pragma unroll
    for (int q = 2; q < T; ++q) w[q] = c_mul(w[q >> 1], w[q - (q >> 1)]);
#pragma unroll
    for (int q = 1; q < T; ++q) v[q] = c_mul(v[q], w[q]);
    if constexpr (R > 8) {
        const C w8 = c_mul(w[4], w[4]);
        C base = w8;
#pragma unroll
        for (int c8 = 8; c8 < R; c8 += 8) {
            v[c8] = c_mul(v[c8], base);
#pragma unroll
            for (int q = 1; q < 8; ++q)
                if (c8 + q < R) v[c8 + q] = c_mul(v[c8 + q], c_mul(base, w[q]));
            if (c8 + 8 < R) base = c_mul(base, w8);
        }
    }
}

// One stage over the whole array, executed cooperatively by nthr threads.
// DIF == false: twiddle then butterfly (decimation in time stage B_s)
// DIF == true : butterfly then twiddle (its transpose)
template <int R, typename C>
ECT_HD void stage_addr(int b, int L, int lshift, int& base, int& k) {
    int blk;
    if (lshift >= 0) { blk = b >> lshift; k = b & (L - 1); }
    else { blk = b / L; k = b - blk * L; }
    base = blk * R * L + k;
}
// power-of-two / composite radix: twiddles + butterfly on a register-resident column
template <int R, bool DIF, typename C>
ECT_HD void stage_math_pow2(C* v, bool tw, C w1, const C* __restrict__ rt_all) {
    if (!DIF && tw) apply_twiddles<R>(v, w1);
    bfly_reg<R>(v, rt_all);
    if (DIF && tw) apply_twiddles<R>(v, w1);
}
template <int R, bool DIF, typename C>
ECT_HD void fft_stage_r(C* data, int n, int L, int lshift, const EctTwT<C> qt,
                        const C* __restrict__ rt_all, int tid, int nthr) {
    const C* rt = rt_all + ECT_ROOTS_OFF(R < ECT_MAX_RADIX ? R : ECT_MAX_RADIX);
    const int nb = n / R;
    const int tstride = n / (R * L);   // twiddle index stride: exp(2 pi i q k / (R L))
    for (int b = tid; b < nb; b += nthr) {
        int base, k;
        stage_addr<R, C>(b, L, lshift, base, k);
        C v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = data[ECT_PAD(base + q * L)];
        C w1 = c_make<C>(1, 0);
        const bool tw = (L > 1) && (k > 0);
        if (tw) w1 = tw_get(qt, k * tstride);
        if constexpr (ECT_REG_RADIX(R)) {
            stage_math_pow2<R, DIF>(v, tw, w1, rt_all);
#pragma unroll
            for (int q = 0; q < R; ++q) data[ECT_PAD(base + q * L)] = v[q];
        } else {
            if (!DIF && tw) apply_twiddles<R>(v, w1);
            if (DIF && tw) {
                // outputs arrive as (p, R-p) pairs: w^p by running product, w^(R-p) = w^R * conj(w^p)
                const C wr = tw_get(qt, k * (n / L));      // w^R = exp(2 pi i k / L)
                C wp = c_make<C>(1, 0);
                int last = 0;
                bfly_odd<R>(v, rt, [&](int q, C val) {
                    if (q > 0) {
                        if (q <= (R - 1) / 2) {
                            if (q != last) { wp = c_mul(wp, w1); last = q; }
                            val = c_mul(val, wp);
                        } else {
                            val = c_mul(val, c_mul(wr, c_make<C>(wp.x, -wp.y)));
                        }
                    }
                    data[ECT_PAD(base + q * L)] = val;
                });
            } else {
                bfly_odd<R>(v, rt, [&](int q, C val) { data[ECT_PAD(base + q * L)] = val; });
            }
        }
    }
}

// COMP: the plan may contain the composite radices 6 / 10 / 12 / 14 (chirp-z half plans)
template <bool DIF, int MAXR = ECT_MAX_RADIX, bool COMP = true, typename C>
ECT_HD void fft_stage(C* data, int n, int r, int L, int lshift, const EctTwT<C> qt,
                      const C* __restrict__ rt_all, int tid, int nthr) {
    const C* rt = rt_all;
    switch (r) {
        case 14: if constexpr (COMP) fft_stage_r<14, DIF>(data, n, L, lshift, qt, rt, tid, nthr); break;
        case 12: if constexpr (COMP) fft_stage_r<12, DIF>(data, n, L, lshift, qt, rt, tid, nthr); break;
        case 10: if constexpr (COMP) fft_stage_r<10, DIF>(data, n, L, lshift, qt, rt, tid, nthr); break;
        case 6:  if constexpr (COMP) fft_stage_r<6, DIF>(data, n, L, lshift, qt, rt, tid, nthr); break;
        case 16: fft_stage_r<16, DIF>(data, n, L, lshift, qt, rt, tid, nthr); break;
        case 8:  fft_stage_r<8, DIF>(data, n, L, lshift, qt, rt, tid, nthr); break;
        case 4:  fft_stage_r<4, DIF>(data, n, L, lshift, qt, rt, tid, nthr); break;
        case 2:  fft_stage_r<2, DIF>(data, n, L, lshift, qt, rt, tid, nthr); break;
        case 3:  fft_stage_r<3, DIF>(data, n, L, lshift, qt, rt, tid, nthr); break;
        case 5:  fft_stage_r<5, DIF>(data, n, L, lshift, qt, rt, tid, nthr); break;
        case 7:  fft_stage_r<7, DIF>(data, n, L, lshift, qt, rt, tid, nthr); break;
        case 11: if constexpr (MAXR >= 11) fft_stage_r<11, DIF>(data, n, L, lshift, qt, rt, tid, nthr); break;
        case 13: if constexpr (MAXR >= 13) fft_stage_r<13, DIF>(data, n, L, lshift, qt, rt, tid, nthr); break;
        case 17: if constexpr (MAXR >= 17) fft_stage_r<17, DIF>(data, n, L, lshift, qt, rt, tid, nthr); break;
        case 19: if constexpr (MAXR >= 19) fft_stage_r<19, DIF>(data, n, L, lshift, qt, rt, tid, nthr); break;
        case 23: if constexpr (MAXR >= 23) fft_stage_r<23, DIF>(data, n, L, lshift, qt, rt, tid, nthr); break;
        case 29: if constexpr (MAXR >= 29) fft_stage_r<29, DIF>(data, n, L, lshift, qt, rt, tid, nthr); break;
        case 31: if constexpr (MAXR >= 31) fft_stage_r<31, DIF>(data, n, L, lshift, qt, rt, tid, nthr); break;
        default: break;
    }
}

// Chirp-z middle step: the innermost DIF stage (L = 1, no twiddles), the pointwise product with the
// (digit-reversed, 1/M-scaled) kernel spectrum and the innermost DIT stage fused into one pass.
// The forward transform runs on swapped (re <-> im) data, the product un-swaps it.
template <int R, typename C>
ECT_HD void blue_middle_r(C* data, int n, const C* __restrict__ bhat, int tid, int nthr) {
    const int nb = n / R;
    for (int b = tid; b < nb; b += nthr) {
        C v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = data[ECT_PAD(b * R + q)];
        bfly_pow2<R>(v);
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = c_mul(c_make<C>(v[q].y, v[q].x), bhat[q * nb + b]);   // bhat stored [q][b]: coalesced
        bfly_pow2<R>(v);
#pragma unroll
        for (int q = 0; q < R; ++q) data[ECT_PAD(b * R + q)] = v[q];
    }
}
// The same step with the kernel spectrum fetched BEFORE the work array is read: the global-memory (L2) latency of
// the R spectrum values then overlaps the shared-memory loads and the first butterfly (split chirp-z kernels, where
// no staging phase hides it)
template <int R, typename C>
ECT_HD void blue_middle_early_r(C* data, int n, const C* __restrict__ bhat, int tid, int nthr) {
    const int nb = n / R;
    for (int b = tid; b < nb; b += nthr) {
        C bh[R], v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) bh[q] = bhat[q * nb + b];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = data[ECT_PAD(b * R + q)];
        bfly_pow2<R>(v);
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = c_mul(c_make<C>(v[q].y, v[q].x), bh[q]);
        bfly_pow2<R>(v);
#pragma unroll
        for (int q = 0; q < R; ++q) data[ECT_PAD(b * R + q)] = v[q];
    }
}
template <typename C>
ECT_HD void blue_middle_early(C* data, int n, int r, const C* __restrict__ bhat, int tid, int nthr) {
    switch (r) {
        case 16: blue_middle_early_r<16>(data, n, bhat, tid, nthr); break;
        case 8:  blue_middle_early_r<8>(data, n, bhat, tid, nthr); break;
        case 4:  blue_middle_early_r<4>(data, n, bhat, tid, nthr); break;
        case 2:  blue_middle_early_r<2>(data, n, bhat, tid, nthr); break;
        default: break;
    }
}
template <typename C>
ECT_HD void blue_middle(C* data, int n, int r, const C* __restrict__ bhat, int tid, int nthr) {
    switch (r) {
        case 16: blue_middle_r<16>(data, n, bhat, tid, nthr); break;
        case 8:  blue_middle_r<8>(data, n, bhat, tid, nthr); break;
        case 4:  blue_middle_r<4>(data, n, bhat, tid, nthr); break;
        case 2:  blue_middle_r<2>(data, n, bhat, tid, nthr); break;
        default: break;
    }
}
