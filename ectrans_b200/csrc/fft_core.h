// In-shared-memory mixed-radix complex FFT core (sign +, unnormalised) used by the
// Fourier stage (FTINV / FTDIR, reference cpu/internal/ftinv_mod.F90:65-84,
// ftdir_mod.F90:67-84; the reference calls FFTW there, tpm_fftw.F90:251-377).
//
// Everything here is __host__ __device__ so the index logic can be exercised on the
// CPU by tests/ (tests/hostemu) -- the product only ever calls it from CUDA kernels.
//
// Data: double2 array of length n, processed in place.
//   DIT:  input at digit-reversed positions (perm table), output natural order.
//   DIF:  input natural order, output at digit-reversed positions.
// Stage s (0 = innermost) has radix r_s and sub-length L_s = prod_{j<s} r_j.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define ECT_HD __host__ __device__ __forceinline__
#define ECT_MAX_STAGES 14
#define ECT_MAX_RADIX 31

struct EctFftPlan {        // POD; one per transform length
    int n;                 // length (even)
    int nst;               // number of stages
    int quarter;           // 1: n % 4 == 0, twiddle table holds exp(2 pi i j/n) for j <= n/4; 0: j < n/2
    int radix[ECT_MAX_STAGES];
    int sublen[ECT_MAX_STAGES];   // L_s
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

// ---- butterflies: u[p] = sum_q v[q] exp(+2 pi i p q / R), in place on v[0..R) ----
ECT_HD void bfly2(double2* v) {
    double2 a = v[0], b = v[1];
    v[0] = c_add(a, b);
    v[1] = c_sub(a, b);
}
ECT_HD void bfly4(double2* v) {
    double2 t0 = c_add(v[0], v[2]), t1 = c_sub(v[0], v[2]);
    double2 t2 = c_add(v[1], v[3]), t3 = c_muli(c_sub(v[1], v[3]));
    v[0] = c_add(t0, t2);
    v[1] = c_add(t1, t3);
    v[2] = c_sub(t0, t2);
    v[3] = c_sub(t1, t3);
}
// odd radix R: rt[j] = (cos, sin)(2 pi j / R), j < R.  Outputs are handed to emit(p, value) as they are
// produced so that only the R inputs live in registers.
template <int R, typename Emit>
ECT_HD void bfly_odd(double2* v, const double2* __restrict__ rt, Emit emit) {
    constexpr int H = (R - 1) / 2;
    // in place: v[q] <- v[q] + v[R-q] (t_q), v[R-q] <- v[q] - v[R-q] (d_q)
    double2 s0 = v[0];
#pragma unroll
    for (int q = 1; q <= H; ++q) {
        const double2 a = v[q], b = v[R - q];
        v[q] = c_add(a, b);
        v[R - q] = c_sub(a, b);
        s0 = c_add(s0, v[q]);
    }
    emit(0, s0);
#pragma unroll
    for (int p = 1; p <= H; ++p) {
        double mx = v[0].x, my = v[0].y, nx = 0.0, ny = 0.0;
#pragma unroll
        for (int q = 1; q <= H; ++q) {
            const double2 w = rt[(p * q) % R];
            mx += w.x * v[q].x;
            my += w.x * v[q].y;
            nx += w.y * v[R - q].x;
            ny += w.y * v[R - q].y;
        }
        // n = i * (nx + i ny) = (-ny, nx)
        emit(p, make_double2(mx - ny, my + nx));
        emit(R - p, make_double2(mx + ny, my - nx));
    }
}

// One stage over the whole array, executed cooperatively by nthr threads.
// DIF == false: twiddle then butterfly (decimation in time stage B_s)
// DIF == true : butterfly then twiddle (its transpose)
template <int R, bool DIF>
ECT_HD void fft_stage_r(double2* data, int n, int L, const double2* __restrict__ qt,
                        const double2* __restrict__ rt, int tid, int nthr) {
    const int nb = n / R;
    const int n4 = (n & 3) ? -(n >> 1) : (n >> 2);
    const int tstride = n / (R * L);   // twiddle index stride: exp(2 pi i q k / (R L))
    for (int b = tid; b < nb; b += nthr) {
        const int blk = b / L;
        const int k = b - blk * L;
        double2* p = data + (size_t)blk * R * L + k;
        const int kt = k * tstride;
        double2 v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = p[q * L];
        if (!DIF && L > 1) {
#pragma unroll
            for (int q = 1; q < R; ++q) v[q] = c_mul(v[q], tw_lookup(qt, q * kt, n4));
        }
        if constexpr (R == 2 || R == 4) {
            if constexpr (R == 2) bfly2(v); else bfly4(v);
            if (DIF && L > 1) {
#pragma unroll
                for (int q = 1; q < R; ++q) v[q] = c_mul(v[q], tw_lookup(qt, q * kt, n4));
            }
#pragma unroll
            for (int q = 0; q < R; ++q) p[q * L] = v[q];
        } else {
            bfly_odd<R>(v, rt, [&](int q, double2 val) {
                if (DIF && L > 1 && q > 0) val = c_mul(val, tw_lookup(qt, q * kt, n4));
                p[q * L] = val;
            });
        }
    }
}

// rt_all: concatenated root tables; the table of odd radix R occupies [R(R-1)/2, R(R+1)/2)
#define ECT_ROOTS_OFF(R) ((R) * ((R) - 1) / 2)
#define ECT_ROOTS_SIZE (ECT_MAX_RADIX * (ECT_MAX_RADIX + 1) / 2)
template <bool DIF, int MAXR = ECT_MAX_RADIX>
ECT_HD void fft_stage(double2* data, int n, int r, int L, const double2* __restrict__ qt,
                      const double2* __restrict__ rt_all, int tid, int nthr) {
    const double2* rt = rt_all + ECT_ROOTS_OFF(r);
    switch (r) {
        case 2:  fft_stage_r<2, DIF>(data, n, L, qt, rt, tid, nthr); break;
        case 4:  fft_stage_r<4, DIF>(data, n, L, qt, rt, tid, nthr); break;
        case 3:  fft_stage_r<3, DIF>(data, n, L, qt, rt, tid, nthr); break;
        case 5:  fft_stage_r<5, DIF>(data, n, L, qt, rt, tid, nthr); break;
        case 7:  fft_stage_r<7, DIF>(data, n, L, qt, rt, tid, nthr); break;
        case 11: if constexpr (MAXR >= 11) fft_stage_r<11, DIF>(data, n, L, qt, rt, tid, nthr); break;
        case 13: if constexpr (MAXR >= 13) fft_stage_r<13, DIF>(data, n, L, qt, rt, tid, nthr); break;
        case 17: if constexpr (MAXR >= 17) fft_stage_r<17, DIF>(data, n, L, qt, rt, tid, nthr); break;
        case 19: if constexpr (MAXR >= 19) fft_stage_r<19, DIF>(data, n, L, qt, rt, tid, nthr); break;
        case 23: if constexpr (MAXR >= 23) fft_stage_r<23, DIF>(data, n, L, qt, rt, tid, nthr); break;
        case 29: if constexpr (MAXR >= 29) fft_stage_r<29, DIF>(data, n, L, qt, rt, tid, nthr); break;
        case 31: if constexpr (MAXR >= 31) fft_stage_r<31, DIF>(data, n, L, qt, rt, tid, nthr); break;
        default: break;
    }
}
