// Legendre stage of the spectral transform for sm_100a.
//
//   k_supolf_table   : SETUP_TRANS table fill (reference cpu/internal/suleg_mod.F90:597-760, SUPOLF)
//   k_ltinv_prologue : PRFI1B + VDTUV + SPNSDE (cpu/internal/prfi1b_mod.F90:81-115,
//                      vdtuv_mod.F90:97-143, spnsde_mod.F90:95-114)
//   k_leinv          : LEINV + ASRE1B  (leinv_mod.F90:116-186, asre1b_mod.F90:88-102): both parities
//                      of one (m, latitude tile, field tile) in one CTA on FP64 DMMA tensor cores,
//                      north = S + A / south = S - A written straight into the Fourier buffer
//   k_ledir          : PRFI2B + LDFOU2 + LEDIR (prfi2b_mod.F90:84-94, ldfou2_mod.F90:90-96,
//                      ledir_mod.F90:118-261): N/S split, Gaussian weights and 1/(a cos) folded into
//                      the B-operand loader
//   k_ltdir_epilogue : UVTVD + UPDSP/UPDSPB (uvtvd_mod.F90:91-139, updsp_mod.F90:104-161,
//                      updspb_mod.F90:92-149)
#include "ect_internal.h"
#include "supolf.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>

// ------------------------------------------------------------------------------------------
// small PTX helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& d0, double& d1, const double a, const double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(s), "l"(gmem), "r"(sz));
}
// ---- bulk asynchronous copies (TMA engine, 1-D form: no tensor map) completing on an mbarrier ----
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" :: "r"(addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem), "r"(bytes),
                    "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N)); }

// ------------------------------------------------------------------------------------------
// table fill
// ------------------------------------------------------------------------------------------
__global__ void k_supolf_table(const EctLegM* __restrict__ legm, const EctSupolfM* __restrict__ cm,
                               const double* __restrict__ rmu, double* __restrict__ ptab, int nsmax) {
    const EctLegM lm = legm[blockIdx.y];
    const int par = blockIdx.z;                 // 0: n-m even (symmetric), 1: odd (antisymmetric)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= lm.ndglu) return;
    const int imaxn = nsmax + 1;
    const bool even = ((imaxn - lm.m) % 2) == 0;
    // INMAX: suleg_mod.F90:646-650 (antisymmetric), :928-932 (symmetric)
    const int knsmax = par ? (even ? imaxn + 1 : imaxn) : (even ? imaxn : imaxn + 1);
    const int kcount = par ? lm.ila : lm.ils;
    if (kcount == 0) return;
    double* out = ptab + (par ? lm.pa_off : lm.ps_off) + i;
    ect_supolf_column(lm.m, par, kcount, knsmax, rmu[lm.isl + i], cm[blockIdx.y], out, lm.ldp);
}

// ------------------------------------------------------------------------------------------
// inverse prologue: user spectral arrays -> X[row = (m, n)][c = 2 f + reim]
// ------------------------------------------------------------------------------------------
struct EctSpecField { const double* base; long long stride; };   // value(ispec) = base[ispec * stride]

struct ProArgs {
    const EctLegM* legm; const int* nasm0;
    const EctSpecField* vor; const EctSpecField* div; const EctSpecField* sc;
    double* x; int cp; int nsmax;
    int kf_uv, kf_sc, scders, vorgp, divgp, fp32;
    int adj;      // DIR_TRANSAD: UVTVD^T = VDTUV . diag(n (n + 1) / a^2), the inputs are scaled on load
    // sp handles on the tensor-core contraction (legendre_tc.cu): rows of m > 0 are written as hi = tf32(x) and
    // lo = x - hi float rows in parity-split order (m, parity, k) instead of the double rows (m = 0 stays double)
    float* xh; float* xl;
};

__device__ __forceinline__ double d_eps(int m, int n) {     // pre_suleg_mod.F90:46-65
    const double dn = (double)n, dm = (double)m;
    return sqrt((dn * dn - dm * dm) / (4.0 * dn * dn - 1.0));
}
__device__ __forceinline__ double d_lap(int n) {            // RLAPIN
    return n >= 1 ? -(ECT_RA * ECT_RA / ((double)n * (double)(n + 1))) : 0.0;
}
template <bool FP32>
__device__ __forceinline__ double2 ld_spec(const EctSpecField f, int idx, bool valid, bool m0) {
    if (!valid) return make_double2(0.0, 0.0);
    if (FP32) {
        const float* b = reinterpret_cast<const float*>(f.base);
        const double re = (double)b[(long long)idx * f.stride];
        const double im = m0 ? 0.0 : (double)b[(long long)(idx + 1) * f.stride];
        return make_double2(re, im);
    }
    const double re = f.base[(long long)idx * f.stride];
    const double im = m0 ? 0.0 : f.base[(long long)(idx + 1) * f.stride];
    return make_double2(re, im);
}

#define PRO_ROWS 4
template <bool FP32>
__global__ void k_ltinv_prologue(ProArgs a) {
    const EctLegM lm = a.legm[blockIdx.y];
    const int m = lm.m, T = a.nsmax;
    // a CTA walks PRO_ROWS consecutive rows n of one m (one CTA per row made the launch rate, not HBM, the limit)
    for (int r = blockIdx.x * PRO_ROWS; r < min((int)(blockIdx.x + 1) * PRO_ROWS, T - m + 2); ++r) {
    const int n = m + r;
    const double zl = (double)m * d_lap(n);                                       // zkm * lapin(n)
    const double c1 = ((double)(n - 1) * d_eps(m, n)) * d_lap(n - 1);
    const double c2 = ((double)(n + 2) * d_eps(m, n + 1)) * d_lap(n + 1);
    const double e1 = (double)(n - 1) * d_eps(m, n);                              // spnsde
    const double e2 = (double)(n + 2) * d_eps(m, n + 1);
    // adjoint of UVTVD: -1 / RLAPIN(n') = n' (n' + 1) / a^2 for the rows n, n - 1, n + 1 read below
    auto ilap = [](int k) { return k >= 1 ? (double)k * (double)(k + 1) / (ECT_RA * ECT_RA) : 0.0; };
    const double g0 = ilap(n), gm = ilap(n - 1), gp = ilap(n + 1);
    const int base = a.nasm0[blockIdx.y];
    const int idx = base + 2 * r;
    const bool m0 = (m == 0);
    const bool v0 = (n <= T), vm = (n - 1 >= m), vp = (n + 1 <= T);
    double* row = a.x + (lm.xrow0 + r) * (long long)a.cp;
    const bool tc = a.xh != nullptr && m > 0;
    const long long frow = (lm.xrow0 + ((r & 1) ? lm.ils : 0) + (r >> 1)) * (long long)a.cp;
    auto put = [&](int c, double2 v) {
        if (tc) {
            const float f0 = (float)v.x, f1 = (float)v.y;
            const float h0 = __uint_as_float(__float_as_uint(f0) & 0xffffe000u), h1 = __uint_as_float(__float_as_uint(f1) & 0xffffe000u);
            *reinterpret_cast<float2*>(a.xh + frow + c) = make_float2(h0, h1);
            *reinterpret_cast<float2*>(a.xl + frow + c) = make_float2(f0 - h0, f1 - h1);
        } else *reinterpret_cast<double2*>(row + c) = v;
    };
    const int o_vor = 0, o_div = a.vorgp ? a.kf_uv : 0;
    const int o_u = o_div + (a.divgp ? a.kf_uv : 0), o_v = o_u + a.kf_uv;
    const int o_sc = o_v + a.kf_uv, o_nsd = o_sc + a.kf_sc;
    for (int j = threadIdx.x; j < a.kf_uv + a.kf_sc; j += blockDim.x) {
        if (j < a.kf_uv) {
            const EctSpecField fv = a.vor[j], fd = a.div[j];
            double2 z0 = ld_spec<FP32>(fv, idx, v0, m0), zm = ld_spec<FP32>(fv, idx - 2, vm, m0), zp = ld_spec<FP32>(fv, idx + 2, vp, m0);
            double2 d0 = ld_spec<FP32>(fd, idx, v0, m0), dm = ld_spec<FP32>(fd, idx - 2, vm, m0), dp = ld_spec<FP32>(fd, idx + 2, vp, m0);
            if (a.adj) {
                z0.x *= g0; z0.y *= g0; d0.x *= g0; d0.y *= g0;
                zm.x *= gm; zm.y *= gm; dm.x *= gm; dm.y *= gm;
                zp.x *= gp; zp.y *= gp; dp.x *= gp; dp.y *= gp;
            }
            // vdtuv_mod.F90:121-139
            double2 u, v;
            u.x = -zl * d0.y + c1 * zm.x - c2 * zp.x;
            u.y = zl * d0.x + c1 * zm.y - c2 * zp.y;
            v.x = -zl * z0.y - c1 * dm.x + c2 * dp.x;
            v.y = zl * z0.x - c1 * dm.y + c2 * dp.y;
            if (m0) { u.y = 0.0; v.y = 0.0; }
            if (a.vorgp) put(2 * (o_vor + j), z0);
            if (a.divgp) put(2 * (o_div + j), d0);
            put(2 * (o_u + j), u);
            put(2 * (o_v + j), v);
        } else {
            const int s = j - a.kf_uv;
            const EctSpecField f = a.sc[s];
            const double2 f0 = ld_spec<FP32>(f, idx, v0, m0);
            put(2 * (o_sc + s), f0);
            if (a.scders) {
                const double2 fm = ld_spec<FP32>(f, idx - 2, vm, m0), fp = ld_spec<FP32>(f, idx + 2, vp, m0);
                put(2 * (o_nsd + s), make_double2(-e1 * fm.x + e2 * fp.x, -e1 * fm.y + e2 * fp.y));
            }
        }
    }
    }
}

void ect_launch_ltinv_prologue(EctHandle* h, const EctFieldCfg& f, const void* d_vor, const void* d_div,
                               const void* d_sc) {
    EctDevice* d = h->d;
    ProArgs a;
    a.legm = d->legm; a.nasm0 = d->nasm0;
    a.vor = (const EctSpecField*)d_vor; a.div = (const EctSpecField*)d_div; a.sc = (const EctSpecField*)d_sc;
    a.x = d->xwork; a.cp = f.cp; a.nsmax = h->hp.nsmax;
    a.kf_uv = f.kf_uv; a.kf_sc = f.kf_sc; a.scders = f.scders; a.vorgp = f.vorgp; a.divgp = f.divgp;
    a.adj = f.adj;
    a.xh = a.xl = nullptr;
    if (h->hp.nump == 0) return;
    if (f.fp32 && ect_tc_enabled(h) && ect_tc_operands(h, f.cp, &a.xh, &a.xl) != ECT_SUCCESS) { d->launch_error = 1; return; }
    dim3 grid((h->hp.nsmax + 2 + PRO_ROWS - 1) / PRO_ROWS, h->hp.nump);
    int items = f.kf_uv + f.kf_sc;
    int threads = items >= 192 ? 256 : (items >= 96 ? 128 : 64);
    a.fp32 = f.fp32;
    if (f.fp32) k_ltinv_prologue<true><<<grid, threads, 0, d->stream>>>(a);
    else k_ltinv_prologue<false><<<grid, threads, 0, d->stream>>>(a);
    d->launches++;
}

// ------------------------------------------------------------------------------------------
// direct epilogue: POA[row = (m, n)][c] -> user spectral arrays
// ------------------------------------------------------------------------------------------
struct EpiArgs {
    const EctLegM* legm; const int* nasm0;
    EctSpecField* vor; EctSpecField* div; EctSpecField* sc;    // bases are written through
    const double* x; int cp; int nsmax;
    int kf_uv, kf_sc;
    int adj;      // INV_TRANSAD: VDTUV^T = diag(-RLAPIN(n)) . UVTVD, and the results are added to the caller's arrays
};

template <bool FP32>
__device__ __forceinline__ void st_spec(const EctSpecField f, int idx, double re, double im, bool accumulate = false) {
    if (FP32) {
        float* b = reinterpret_cast<float*>(const_cast<double*>(f.base));
        if (accumulate) { re += (double)b[(long long)idx * f.stride]; im += (double)b[(long long)(idx + 1) * f.stride]; }
        b[(long long)idx * f.stride] = (float)re;
        b[(long long)(idx + 1) * f.stride] = (float)im;
    } else {
        double* b = const_cast<double*>(f.base);
        if (accumulate) { re += b[(long long)idx * f.stride]; im += b[(long long)(idx + 1) * f.stride]; }
        b[(long long)idx * f.stride] = re;
        b[(long long)(idx + 1) * f.stride] = im;
    }
}

template <bool FP32>
__global__ void k_ltdir_epilogue(EpiArgs a) {
    const EctLegM lm = a.legm[blockIdx.y];
    const int m = lm.m, T = a.nsmax;
    for (int r = blockIdx.x * PRO_ROWS; r < min((int)(blockIdx.x + 1) * PRO_ROWS, T - m + 1); ++r) {
    const int n = m + r;
    const double c1 = (double)n * d_eps(m, n + 1);          // ZN(JN)*PEPSNM(JN+1)
    const double c2 = (double)(n + 1) * d_eps(m, n);        // ZN(JN+1)*PEPSNM(JN)
    const double zkm = (double)m;
    const int idx = a.nasm0[blockIdx.y] + 2 * r;
    const bool m0 = (m == 0);
    const double* row = a.x + (lm.xrow0 + r) * (long long)a.cp;
    const double* rowp = row + a.cp;                   // n + 1 (always exists: rows go to T+1)
    const double* rowm = row - a.cp;                   // n - 1 (valid if r > 0)
    const int o_u = 0, o_v = a.kf_uv, o_sc = 2 * a.kf_uv;
    for (int j = threadIdx.x; j < a.kf_uv + a.kf_sc; j += blockDim.x) {
        if (j < a.kf_uv) {
            const double2 u0 = *reinterpret_cast<const double2*>(row + 2 * (o_u + j));
            const double2 v0 = *reinterpret_cast<const double2*>(row + 2 * (o_v + j));
            const double2 up = *reinterpret_cast<const double2*>(rowp + 2 * (o_u + j));
            const double2 vp = *reinterpret_cast<const double2*>(rowp + 2 * (o_v + j));
            double2 um = make_double2(0.0, 0.0), vm = um;
            if (r > 0) {
                um = *reinterpret_cast<const double2*>(rowm + 2 * (o_u + j));
                vm = *reinterpret_cast<const double2*>(rowm + 2 * (o_v + j));
            }
            // uvtvd_mod.F90:104-139
            double2 vor, div;
            vor.x = -zkm * v0.y - c1 * up.x + c2 * um.x;
            vor.y = zkm * v0.x - c1 * up.y + c2 * um.y;
            div.x = -zkm * u0.y + c1 * vp.x - c2 * vm.x;
            div.y = zkm * u0.x + c1 * vp.y - c2 * vm.y;
            if (m0) { vor.y = 0.0; div.y = 0.0; if (n == 0) { vor.x = 0.0; div.x = 0.0; } }
            if (a.adj) { const double sl = -d_lap(n); vor.x *= sl; vor.y *= sl; div.x *= sl; div.y *= sl; }
            st_spec<FP32>(a.vor[j], idx, vor.x, vor.y, a.adj != 0);
            st_spec<FP32>(a.div[j], idx, div.x, div.y, a.adj != 0);
        } else {
            const int s = j - a.kf_uv;
            double2 f0 = *reinterpret_cast<const double2*>(row + 2 * (o_sc + s));
            if (m0) f0.y = 0.0;
            st_spec<FP32>(a.sc[s], idx, f0.x, f0.y, a.adj != 0);
        }
    }
    }
}

void ect_launch_ltdir_epilogue(EctHandle* h, const EctFieldCfg& f, void* d_vor, void* d_div, void* d_sc) {
    EctDevice* d = h->d;
    EpiArgs a;
    a.legm = d->legm; a.nasm0 = d->nasm0;
    a.vor = (EctSpecField*)d_vor; a.div = (EctSpecField*)d_div; a.sc = (EctSpecField*)d_sc;
    a.x = d->xwork; a.cp = f.cp; a.nsmax = h->hp.nsmax;
    a.kf_uv = f.kf_uv; a.kf_sc = f.kf_sc; a.adj = f.adj;
    if (h->hp.nump == 0) return;
    dim3 grid((h->hp.nsmax + 1 + PRO_ROWS - 1) / PRO_ROWS, h->hp.nump);
    int items = f.kf_uv + f.kf_sc;
    int threads = items >= 192 ? 256 : (items >= 96 ? 128 : 64);
    if (f.fp32) k_ltdir_epilogue<true><<<grid, threads, 0, d->stream>>>(a);
    else k_ltdir_epilogue<false><<<grid, threads, 0, d->stream>>>(a);
    d->launches++;
}

// ------------------------------------------------------------------------------------------
// VORDIV_TO_UV (vd2uv_mod.F90:86-112): spectral (vor, div) -> spectral (U, V) cos(theta) / a for n <= T, arrays
// (nfld, nspec2).  legm == nullptr: one task holding every m, m-major (NASM0(m) = m (2T + 3 - m)).
// ------------------------------------------------------------------------------------------
template <bool FP32>
__global__ void k_vd2uv(const void* vor, const void* div, void* u, void* v, int nfld, int T, const EctLegM* legm,
                        const int* nasm0) {
    const int m = legm ? legm[blockIdx.y].m : (int)blockIdx.y;
    const int base = legm ? nasm0[blockIdx.y] : m * (2 * T + 3 - m);
    const bool m0 = (m == 0);
    const double ra_r = 1.0 / ECT_RA;
    for (int r = blockIdx.x * PRO_ROWS; r < min((int)(blockIdx.x + 1) * PRO_ROWS, T - m + 1); ++r) {
        const int n = m + r, idx = base + 2 * r;
        const double zl = (double)m * d_lap(n);
        const double c1 = ((double)(n - 1) * d_eps(m, n)) * d_lap(n - 1);
        const double c2 = ((double)(n + 2) * d_eps(m, n + 1)) * d_lap(n + 1);
        const bool vm = (n - 1 >= m), vp = (n + 1 <= T);
        for (int j = threadIdx.x; j < nfld; j += blockDim.x) {
            const EctSpecField fv{reinterpret_cast<const double*>(FP32 ? (const void*)(reinterpret_cast<const float*>(vor) + j) : (const void*)(reinterpret_cast<const double*>(vor) + j)), nfld};
            const EctSpecField fd{reinterpret_cast<const double*>(FP32 ? (const void*)(reinterpret_cast<const float*>(div) + j) : (const void*)(reinterpret_cast<const double*>(div) + j)), nfld};
            const double2 z0 = ld_spec<FP32>(fv, idx, true, m0), zm = ld_spec<FP32>(fv, idx - 2, vm, m0), zp = ld_spec<FP32>(fv, idx + 2, vp, m0);
            const double2 d0 = ld_spec<FP32>(fd, idx, true, m0), dm = ld_spec<FP32>(fd, idx - 2, vm, m0), dp = ld_spec<FP32>(fd, idx + 2, vp, m0);
            // vdtuv_mod.F90:121-139
            double2 uu, vv;
            uu.x = -zl * d0.y + c1 * zm.x - c2 * zp.x;
            uu.y = zl * d0.x + c1 * zm.y - c2 * zp.y;
            vv.x = -zl * z0.y - c1 * dm.x + c2 * dp.x;
            vv.y = zl * z0.x - c1 * dm.y + c2 * dp.y;
            if (m0) { uu.y = 0.0; vv.y = 0.0; }
            const EctSpecField fu{reinterpret_cast<const double*>(FP32 ? (void*)(reinterpret_cast<float*>(u) + j) : (void*)(reinterpret_cast<double*>(u) + j)), nfld};
            const EctSpecField fw{reinterpret_cast<const double*>(FP32 ? (void*)(reinterpret_cast<float*>(v) + j) : (void*)(reinterpret_cast<double*>(v) + j)), nfld};
            st_spec<FP32>(fu, idx, uu.x * ra_r, uu.y * ra_r);
            st_spec<FP32>(fw, idx, vv.x * ra_r, vv.y * ra_r);
        }
    }
}

int ect_launch_vd2uv(const void* vor, const void* div, void* u, void* v, int nfld, int nsmax, int nump,
                     const EctLegM* legm, const int* nasm0, bool fp32, cudaStream_t st) {
    dim3 grid((nsmax + 1 + PRO_ROWS - 1) / PRO_ROWS, nump);
    const int threads = nfld >= 192 ? 256 : (nfld >= 96 ? 128 : 64);
    if (fp32) k_vd2uv<true><<<grid, threads, 0, st>>>(vor, div, u, v, nfld, nsmax, legm, nasm0);
    else k_vd2uv<false><<<grid, threads, 0, st>>>(vor, div, u, v, nfld, nsmax, legm, nasm0);
    ECT_CUDA(cudaGetLastError());
    return ECT_SUCCESS;
}

// ------------------------------------------------------------------------------------------
// DMMA contraction kernels.  CTA = 128 threads = 4 warps as 2 (M) x 2 (N); CTA tile 64 x 64 for BOTH
// parities, warp tile 32 x 32 -> 2 x 16 m8n8 accumulators per thread; two CTAs per SM so that one CTA's
// barrier / cp.async waits are covered by the other's DMMAs (the first version, one 256-thread CTA per
// SM with a 64 x 128 tile, kept the DMMA pipe 75 % busy: profiles/r01_ncu_full_summary.txt).
// 64-wide field tiles also remove the 6.5 -> 7 tile rounding at 412 fields (832 columns = 13 x 64).
// Shared-memory pitches are = 4 (mod 16) doubles so that the fragment loads
// (row = lane/4, k = lane%4) hit 16 distinct 8-byte banks per half warp.
// ------------------------------------------------------------------------------------------
#define LEG_BM 64
#define LEG_BN 64
#define LEG_THREADS 128
#define LEG_KC 8
#define LEG_STAGES 5
#define LEG_LD (64 + 4)          // 68: pitch of the 8 x 64 tiles
#define DIR_LDA (LEG_KC + 4)     // 12: pitch of the 64 x 8 polynomial tile of the direct kernel
#ifndef INV_KC
#define INV_KC 8
#endif
#ifndef INV_STAGES
#define INV_STAGES 5
#endif
#define INV_STAGE_DOUBLES (4 * INV_KC * LEG_LD)
#define DIR_STAGE_DOUBLES (2 * LEG_BM * DIR_LDA + 2 * LEG_KC * LEG_LD)

struct LegArgs {
    const EctLegM* legm;
    const int2* tiles;            // (ml, tile index along M)
    int nct;                      // field tiles
    const double* ptab;
    double* x;                    // inverse: input X ; direct: output POA
    double* fb;                   // Fourier buffer (Legendre side)
    const int* rec_n; const int* rec_s;
    const double* rw; const double* racthe;
    int cp;
    int c_uv_end;                 // direct: columns < c_uv_end are u,v (scaled by 1/(a cos))
    int dbg;                      // ECT_LEG_DBG timing experiments: 1 skip B preparation, 2 contiguous (wrong) A loads
    // inverse epilogue destinations: rank + record in that rank's Fourier-side buffer (self when one rank / NCCL mode)
    double* const* peer; const int* dst_rank_n; const int* dst_rank_s; const int* dst_rec_n; const int* dst_rec_s;
};

// Variant with the operand tiles staged by the TMA engine (cp.async.bulk + mbarrier), selected by ECT_LEINV_TMA=1.
// Measured on B200 at TCo1279 L137: 70.9 ms against 48.4 ms for the cp.async loader of k_leinv below -- a stage is
// 32 row copies of 512 bytes into the padded (conflict-free) pitch, and two CTAs per SM issue 64 such copies per
// ~1000 clocks, which the bulk-copy path does not sustain at this granularity.  Kept for comparison only.
__global__ void __launch_bounds__(LEG_THREADS, 2) k_leinv_tma(LegArgs a) {
    extern __shared__ __align__(16) double smem[];
    const int tile = blockIdx.x / a.nct, ct = blockIdx.x - tile * a.nct;
    const int2 td = a.tiles[tile];
    const EctLegM lm = a.legm[td.x];
    const int i0 = td.y * LEG_BM, c0 = ct * LEG_BN;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    const int nchunks = (lm.ils + INV_KC - 1) / INV_KC;
    const double* ps = a.ptab + lm.ps_off + i0;
    const double* pa = a.ptab + lm.pa_off + i0;
    const double* xb = a.x + lm.xrow0 * (long long)a.cp + c0;

    // Operand tiles arrive by bulk asynchronous copies (TMA engine): a stage is 4 tiles (P sym, P antisym, X sym,
    // X antisym) x INV_KC rows of 64 doubles = 32 row copies of 512 bytes, one per lane of warp 0, completing on
    // the stage's mbarrier.  Rows past the end of a parity are not copied: their polynomial row is zeroed instead.
    __shared__ __align__(8) unsigned long long s_full[INV_STAGES];
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < INV_STAGES; ++s) mbar_init(&s_full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const unsigned rowb = (unsigned)(min(LEG_BN, a.cp - c0) * (int)sizeof(double));      // X row bytes inside this field tile
    auto load_chunk = [&](int chunk, int buf) {         // called by warp 0 only
        double* As = smem + (size_t)buf * INV_STAGE_DOUBLES;
        const int k0 = chunk * INV_KC;
        const int nvs = min(INV_KC, lm.ils - k0), nva = max(0, min(INV_KC, lm.ila - k0));
        if (lane == 0) mbar_expect_tx(&s_full[buf], (unsigned)(nvs + nva) * (LEG_BM * (unsigned)sizeof(double) + rowb));
        __syncwarp();
        if (4 * INV_KC != 32) __trap();      // one (tile, row) per lane: this variant exists for INV_KC = 8 only
        const int tile4 = lane / INV_KC, row = lane % INV_KC;
        const int k = k0 + row;
        const bool anti = tile4 & 1, isx = tile4 >= 2;
        const bool valid = k < (anti ? lm.ila : lm.ils);
        double* dst = As + (size_t)tile4 * INV_KC * LEG_LD + row * LEG_LD;      // tile order in a stage: As, Aa, Bs, Ba
        if (valid) {
            const double* src = !isx ? (anti ? pa : ps) + (long long)k * lm.ldp
                                     : xb + (long long)(2 * k + (anti ? 1 : 0)) * a.cp;
            bulk_g2s(dst, src, isx ? rowb : (unsigned)(LEG_BM * sizeof(double)), &s_full[buf]);
        } else {       // (last chunk only) zero row: the stale contents of a never-filled stage could be NaN patterns
#pragma unroll 8
            for (int c = 0; c < LEG_BM; c += 2) *reinterpret_cast<double2*>(dst + c) = make_double2(0.0, 0.0);
        }
    };

    double acs[4][4][2], aca[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acs[i][j][0] = acs[i][j][1] = 0.0; aca[i][j][0] = aca[i][j][1] = 0.0; }

    if (warp == 0) {
#pragma unroll
        for (int s = 0; s < INV_STAGES - 1; ++s)
            if (s < nchunks) load_chunk(s, s);
    }
    for (int ch = 0; ch < nchunks; ++ch) {
        mbar_wait(&s_full[ch % INV_STAGES], (unsigned)((ch / INV_STAGES) & 1));
        __syncthreads();           // everybody is done with chunk ch - 1: its stage can be refilled
        if (warp == 0) {
            const int nx = ch + INV_STAGES - 1;
            if (nx < nchunks) {
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                load_chunk(nx, nx % INV_STAGES);
            }
        }
        const double* As = smem + (size_t)(ch % INV_STAGES) * INV_STAGE_DOUBLES;
        const double* Aa = As + INV_KC * LEG_LD;
        const double* Bs = Aa + INV_KC * LEG_LD;
        const double* Ba = Bs + INV_KC * LEG_LD;
#pragma unroll
        for (int kk = 0; kk < INV_KC; kk += 4) {
            double fs[4], fa[4], bs[4], ba[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                fs[i] = As[(kk + t) * LEG_LD + wm * 32 + i * 8 + g];
                fa[i] = Aa[(kk + t) * LEG_LD + wm * 32 + i * 8 + g];
                bs[i] = Bs[(kk + t) * LEG_LD + wn * 32 + i * 8 + g];
                ba[i] = Ba[(kk + t) * LEG_LD + wn * 32 + i * 8 + g];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dmma884(acs[i][j][0], acs[i][j][1], fs[i], bs[j]);
                    dmma884(aca[i][j][0], aca[i][j][1], fa[i], ba[j]);
                }
        }
    }
    // epilogue: north = S + A, south = S - A  (asre1b_mod.F90:99-100)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int li = i0 + wm * 32 + i * 8 + g;
        if (li >= lm.ndglu) continue;
        // TRMTOL fused: the record goes straight into the buffer of the rank that owns the latitude
        double* pn = a.peer[a.dst_rank_n[lm.rec0 + li]] + (long long)a.dst_rec_n[lm.rec0 + li] * a.cp;
        double* psth = a.peer[a.dst_rank_s[lm.rec0 + li]] + (long long)a.dst_rec_s[lm.rec0 + li] * a.cp;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + wn * 32 + j * 8 + 2 * t;
            if (c >= a.cp) continue;
            *reinterpret_cast<double2*>(pn + c) =
                make_double2(acs[i][j][0] + aca[i][j][0], acs[i][j][1] + aca[i][j][1]);
            *reinterpret_cast<double2*>(psth + c) =
                make_double2(acs[i][j][0] - aca[i][j][0], acs[i][j][1] - aca[i][j][1]);
        }
    }
}

// Inverse contraction (default): 16-byte cp.async pieces issued by all threads into a 5-stage ring.
__global__ void __launch_bounds__(LEG_THREADS, 2) k_leinv(LegArgs a) {
    extern __shared__ __align__(16) double smem[];
    const int tile = blockIdx.x / a.nct, ct = blockIdx.x - tile * a.nct;
    const int2 td = a.tiles[tile];
    const EctLegM lm = a.legm[td.x];
    const int i0 = td.y * LEG_BM, c0 = ct * LEG_BN;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    const int nchunks = (lm.ils + INV_KC - 1) / INV_KC;
    const double* ps = a.ptab + lm.ps_off + i0;
    const double* pa = a.ptab + lm.pa_off + i0;
    const double* xb = a.x + lm.xrow0 * (long long)a.cp + c0;
    const int ni = min(4, (lm.ndglu - (i0 + wm * 32) + 7) >> 3);

    auto load_chunk = [&](int chunk, int buf) {
        double* As = smem + (size_t)buf * INV_STAGE_DOUBLES;
        double* Aa = As + INV_KC * LEG_LD;
        double* Bs = Aa + INV_KC * LEG_LD;
        double* Ba = Bs + INV_KC * LEG_LD;
        const int k0 = chunk * INV_KC;
#pragma unroll
        for (int e = 0; e < INV_KC / 4; ++e) {   // INV_KC rows x 64 columns = INV_KC * 32 16-byte pieces per tile
            const int idx = tid + e * LEG_THREADS;
            const int row = idx >> 5, c2 = (idx & 31) * 2;
            const int k = k0 + row;
            const bool vs = k < lm.ils, va = k < lm.ila;
            cp_async16(As + row * LEG_LD + c2, ps + (long long)(vs ? k : 0) * lm.ldp + c2, vs);
            cp_async16(Aa + row * LEG_LD + c2, pa + (long long)(va ? k : 0) * lm.ldp + c2, va);
            const bool cv = (c0 + c2) < a.cp;
            cp_async16(Bs + row * LEG_LD + c2, xb + (long long)(vs ? 2 * k : 0) * a.cp + (cv ? c2 : 0), vs && cv);
            cp_async16(Ba + row * LEG_LD + c2, xb + (long long)(va ? 2 * k + 1 : 0) * a.cp + (cv ? c2 : 0), va && cv);
        }
    };

    double acs[4][4][2], aca[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acs[i][j][0] = acs[i][j][1] = 0.0; aca[i][j][0] = aca[i][j][1] = 0.0; }

#pragma unroll
    for (int s = 0; s < INV_STAGES - 1; ++s) {
        if (s < nchunks) load_chunk(s, s);
        cp_async_commit();
    }
    for (int ch = 0; ch < nchunks; ++ch) {
        cp_async_wait<INV_STAGES - 2>();
        __syncthreads();
        {
            const int nx = ch + INV_STAGES - 1;
            if (nx < nchunks) load_chunk(nx, nx % INV_STAGES);
            cp_async_commit();
        }
        const double* As = smem + (size_t)(ch % INV_STAGES) * INV_STAGE_DOUBLES;
        const double* Aa = As + INV_KC * LEG_LD;
        const double* Bs = Aa + INV_KC * LEG_LD;
        const double* Ba = Bs + INV_KC * LEG_LD;
#pragma unroll
        for (int kk = 0; kk < INV_KC; kk += 4) {
            double fs[4], fa[4], bs[4], ba[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                fs[i] = As[(kk + t) * LEG_LD + wm * 32 + i * 8 + g];
                fa[i] = Aa[(kk + t) * LEG_LD + wm * 32 + i * 8 + g];
                bs[i] = Bs[(kk + t) * LEG_LD + wn * 32 + i * 8 + g];
                ba[i] = Ba[(kk + t) * LEG_LD + wn * 32 + i * 8 + g];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (i >= ni) continue;          // 8-latitude blocks past NDGLU(m) (warp-uniform)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dmma884(acs[i][j][0], acs[i][j][1], fs[i], bs[j]);
                    dmma884(aca[i][j][0], aca[i][j][1], fa[i], ba[j]);
                }
            }
        }
    }
    cp_async_wait<0>();
    // epilogue: north = S + A, south = S - A  (asre1b_mod.F90:99-100)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int li = i0 + wm * 32 + i * 8 + g;
        if (li >= lm.ndglu) continue;
        // TRMTOL fused: the record goes straight into the buffer of the rank that owns the latitude
        double* pn = a.peer[a.dst_rank_n[lm.rec0 + li]] + (long long)a.dst_rec_n[lm.rec0 + li] * a.cp;
        double* psth = a.peer[a.dst_rank_s[lm.rec0 + li]] + (long long)a.dst_rec_s[lm.rec0 + li] * a.cp;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + wn * 32 + j * 8 + 2 * t;
            if (c >= a.cp) continue;
            *reinterpret_cast<double2*>(pn + c) =
                make_double2(acs[i][j][0] + aca[i][j][0], acs[i][j][1] + aca[i][j][1]);
            *reinterpret_cast<double2*>(psth + c) =
                make_double2(acs[i][j][0] - aca[i][j][0], acs[i][j][1] - aca[i][j][1]);
        }
    }
}

// Direct: Psi[k][c] = sum_lat P[k][lat] * (w (N +- S))[lat][c].  North / south records arrive already scaled by
// the Gaussian weight (and 1/(a cos theta) for u, v) from the Fourier stage's store; they are staged with
// cp.async and only the N +- S combination (prfi2b_mod.F90:91-92) is formed when the B fragments are read.
__global__ void __launch_bounds__(LEG_THREADS, 2) k_ledir(LegArgs a) {
    extern __shared__ __align__(16) double smem[];
    const int tile = blockIdx.x / a.nct, ct = blockIdx.x - tile * a.nct;
    const int2 td = a.tiles[tile];
    const EctLegM lm = a.legm[td.x];
    const int kr0 = td.y * LEG_BM, c0 = ct * LEG_BN;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    const int nchunks = (lm.ndglu + LEG_KC - 1) / LEG_KC;
    const double* ps = a.ptab + lm.ps_off;
    const double* pa = a.ptab + lm.pa_off;
    const int ni = min(4, (lm.ils - (kr0 + wm * 32) + 7) >> 3);      // 8-row blocks of this warp that hold an n (ila <= ils)

    // record numbers of the north / south rows a thread stages: fetched one chunk ahead of the cp.async that
    // uses them, so that the dependent index load is never waited for in front of the tensor work
    int rn[2], rs[2];
    auto load_idx = [&](int chunk) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int li = chunk * LEG_KC + ((tid + e * LEG_THREADS) >> 5);
            const bool v = li < lm.ndglu;
            rn[e] = v ? a.rec_n[lm.rec0 + li] : -1;
            rs[e] = v ? a.rec_s[lm.rec0 + li] : -1;
        }
    };
    auto load_chunk = [&](int chunk, int buf) {
        double* As = smem + (size_t)buf * DIR_STAGE_DOUBLES;
        double* Aa = As + LEG_BM * DIR_LDA;
        double* Bn = Aa + LEG_BM * DIR_LDA;
        double* Bs = Bn + LEG_KC * LEG_LD;
        const int l0 = chunk * LEG_KC;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int idx = tid + e * LEG_THREADS;
            {   // polynomial tile P[k][lat chunk]: 64 rows x 8 latitudes = 256 pieces per parity
                const int row = idx >> 2, c2 = (idx & 3) * 2;
                const int k = kr0 + row;
                const bool vs = k < lm.ils, va = k < lm.ila;
                if (a.dbg & 2) {
                    cp_async16(As + row * DIR_LDA + c2, ps + (long long)(chunk & 63) * lm.ldp + idx * 2, true);
                    cp_async16(Aa + row * DIR_LDA + c2, pa + (long long)(chunk & 63) * lm.ldp + idx * 2, true);
                } else {
                cp_async16(As + row * DIR_LDA + c2, ps + (long long)(vs ? k : 0) * lm.ldp + l0 + c2, vs);
                cp_async16(Aa + row * DIR_LDA + c2, pa + (long long)(va ? k : 0) * lm.ldp + l0 + c2, va);
                }
            }
            {   // north / south records: 8 latitudes x 64 columns = 256 pieces per hemisphere
                const int lr = idx >> 5, c2 = (idx & 31) * 2;
                const bool v = rn[e] >= 0 && (c0 + c2) < a.cp;
                const long long on = v ? (long long)rn[e] * a.cp + c0 + c2 : 0;
                const long long os = v ? (long long)rs[e] * a.cp + c0 + c2 : 0;
                cp_async16(Bn + lr * LEG_LD + c2, a.fb + on, v);
                cp_async16(Bs + lr * LEG_LD + c2, a.fb + os, v);
            }
        }
    };

    double acs[4][4][2], aca[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acs[i][j][0] = acs[i][j][1] = 0.0; aca[i][j][0] = aca[i][j][1] = 0.0; }

#pragma unroll
    for (int s = 0; s < LEG_STAGES - 1; ++s) {
        if (s < nchunks) { load_idx(s); load_chunk(s, s); }
        cp_async_commit();
    }
    load_idx(LEG_STAGES - 1);
    for (int ch = 0; ch < nchunks; ++ch) {
        cp_async_wait<LEG_STAGES - 2>();
        __syncthreads();
        {
            const int nx = ch + LEG_STAGES - 1;
            if (nx < nchunks) load_chunk(nx, nx % LEG_STAGES);
            cp_async_commit();
            load_idx(nx + 1);
        }
        const double* As = smem + (size_t)(ch % LEG_STAGES) * DIR_STAGE_DOUBLES;
        const double* Aa = As + LEG_BM * DIR_LDA;
        const double* Bn = Aa + LEG_BM * DIR_LDA;
        const double* Bs = Bn + LEG_KC * LEG_LD;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int kk = 4 * h;
            double fs[4], fa[4], bs[4], ba[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                fs[i] = As[(wm * 32 + i * 8 + g) * DIR_LDA + kk + t];
                fa[i] = Aa[(wm * 32 + i * 8 + g) * DIR_LDA + kk + t];
                const double vn = Bn[(kk + t) * LEG_LD + wn * 32 + i * 8 + g];
                const double vs = Bs[(kk + t) * LEG_LD + wn * 32 + i * 8 + g];
                bs[i] = vn + vs;
                ba[i] = vn - vs;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (i >= ni) continue;          // 8-row blocks past the last n of this m: nothing to accumulate (warp-uniform)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dmma884(acs[i][j][0], acs[i][j][1], fs[i], bs[j]);
                    dmma884(aca[i][j][0], aca[i][j][1], fa[i], ba[j]);
                }
            }
        }
    }
    cp_async_wait<0>();
    // epilogue: symmetric part -> rows n - m even, antisymmetric -> odd  (ledir_mod.F90:174-179, :248-253)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = kr0 + wm * 32 + i * 8 + g;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + wn * 32 + j * 8 + 2 * t;
            if (c >= a.cp) continue;
            if (k < lm.ils)
                *reinterpret_cast<double2*>(a.x + (lm.xrow0 + 2 * k) * (long long)a.cp + c) =
                    make_double2(acs[i][j][0], acs[i][j][1]);
            if (k < lm.ila)
                *reinterpret_cast<double2*>(a.x + (lm.xrow0 + 2 * k + 1) * (long long)a.cp + c) =
                    make_double2(aca[i][j][0], aca[i][j][1]);
        }
    }
}

// cudaFuncSetAttribute applies to the current device only: the opt-in is taken once per device (a process may hold
// handles on several GPUs)
static bool g_leg_attr_set[64] = {};
static void leg_set_attrs() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 63;
    if (g_leg_attr_set[dev] && dev != 63) return;
    cudaFuncSetAttribute(k_leinv, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         INV_STAGES * INV_STAGE_DOUBLES * (int)sizeof(double));
    cudaFuncSetAttribute(k_leinv_tma, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(INV_STAGES * INV_STAGE_DOUBLES * sizeof(double)));
    cudaFuncSetAttribute(k_ledir, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         LEG_STAGES * DIR_STAGE_DOUBLES * (int)sizeof(double));
    g_leg_attr_set[dev] = true;
}

void ect_launch_leinv(EctHandle* h, const EctFieldCfg& f) {
    EctDevice* d = h->d;
    if (d->n_inv_tiles == 0) return;
    leg_set_attrs();
    LegArgs a;
    a.legm = d->legm; a.tiles = d->inv_tiles; a.nct = (f.cp + LEG_BN - 1) / LEG_BN;
    a.ptab = d->ptab; a.x = d->xwork; a.fb = d->fbuf_leg;
    a.rec_n = d->leg_rec_n; a.rec_s = d->leg_rec_s; a.rw = d->rw; a.racthe = d->racthe;
    a.cp = f.cp; a.c_uv_end = 0; a.dbg = 0;
    a.peer = d->peer_fft; a.dst_rank_n = d->leg_dst_rank_n; a.dst_rank_s = d->leg_dst_rank_s;
    a.dst_rec_n = d->leg_dst_rec_n; a.dst_rec_s = d->leg_dst_rec_s;
    const size_t smem = INV_STAGES * INV_STAGE_DOUBLES * sizeof(double);
    if (f.fp32 && ect_tc_enabled(h)) {
        // sp handle: m = 0 in double precision on the DMMA kernel (as the reference keeps it: leinv_mod.F90:264-288),
        // every other wavenumber as 3xTF32 on tcgen05
        if (d->n_inv_tiles_m0 > 0) {
            k_leinv<<<(unsigned)((long long)d->n_inv_tiles_m0 * a.nct), LEG_THREADS, smem, d->stream>>>(a);
            d->launches++;
        }
        if (ect_tc_launch_leinv(h, f) != ECT_SUCCESS) d->launch_error = 1;
        return;
    }
    static const char* tma = getenv("ECT_LEINV_TMA");      // bulk-copy loader: measured slower (70.9 vs 48.4 ms at TCo1279), opt-in
    if (tma && atoi(tma)) k_leinv_tma<<<(unsigned)((long long)d->n_inv_tiles * a.nct), LEG_THREADS, smem, d->stream>>>(a);
    else k_leinv<<<(unsigned)((long long)d->n_inv_tiles * a.nct), LEG_THREADS, smem, d->stream>>>(a);
    d->launches++;
}

void ect_launch_ledir(EctHandle* h, const EctFieldCfg& f) {
    EctDevice* d = h->d;
    if (d->n_dir_tiles == 0) return;
    leg_set_attrs();
    LegArgs a;
    a.legm = d->legm; a.tiles = d->dir_tiles; a.nct = (f.cp + LEG_BN - 1) / LEG_BN;
    a.ptab = d->ptab; a.x = d->xwork; a.fb = d->fbuf_leg;
    a.rec_n = d->leg_rec_n; a.rec_s = d->leg_rec_s; a.rw = d->rw; a.racthe = d->racthe;
    a.cp = f.cp; a.c_uv_end = 4 * f.kf_uv;
    static const char* dbg = getenv("ECT_LEG_DBG");
    a.dbg = dbg ? atoi(dbg) : 0;
    a.peer = nullptr; a.dst_rank_n = a.dst_rank_s = a.dst_rec_n = a.dst_rec_s = nullptr;
    const size_t smem = LEG_STAGES * DIR_STAGE_DOUBLES * sizeof(double);
    if (f.fp32 && ect_tc_enabled(h)) {         // sp handle: m = 0 in double (ledir_mod.F90:133-171), the rest on tcgen05
        if (d->n_dir_tiles_m0 > 0) {
            k_ledir<<<(unsigned)((long long)d->n_dir_tiles_m0 * a.nct), LEG_THREADS, smem, d->stream>>>(a);
            d->launches++;
        }
        if (ect_tc_launch_ledir(h, f) != ECT_SUCCESS) d->launch_error = 1;
        return;
    }
    k_ledir<<<(unsigned)((long long)d->n_dir_tiles * a.nct), LEG_THREADS, smem, d->stream>>>(a);
    d->launches++;
}

// ------------------------------------------------------------------------------------------
// device setup of the Legendre stage
// ------------------------------------------------------------------------------------------
int ect_legendre_setup(EctHandle* h) {
    EctHostPlan& P = h->hp;
    EctDevice* d = h->d;
    const int T = P.nsmax;
    d->h_legm.resize(P.nump);
    i64 off = 0, xrow = 0;
    std::vector<EctSupolfM> cms(P.nump);
    for (int ml = 0; ml < P.nump; ++ml) {
        EctLegM& lm = d->h_legm[ml];
        lm.m = P.myms[ml];
        lm.ndglu = P.ndglu[lm.m];
        lm.ldp = (lm.ndglu + ECT_LAT_PAD - 1) / ECT_LAT_PAD * ECT_LAT_PAD;
        lm.ila = (T - lm.m + 2) / 2;
        lm.ils = (T - lm.m + 3) / 2;
        lm.ps_off = off; off += (i64)lm.ils * lm.ldp;
        lm.pa_off = off; off += (i64)lm.ila * lm.ldp;
        lm.xrow0 = xrow; xrow += T - lm.m + 2;
        lm.rec0 = P.mrow0[ml];
        lm.isl = P.ndgnh - lm.ndglu;
        lm.pad = 0;
        ect_supolf_consts(lm.m, cms[ml]);
    }
    d->ptab_elems = off;
    d->xrows = xrow;
    if (P.nump == 0) return ECT_SUCCESS;
    ECT_CUDA(cudaMalloc(&d->ptab, std::max<i64>(off, 1) * sizeof(double)));
    ECT_CUDA(cudaMemsetAsync(d->ptab, 0, std::max<i64>(off, 1) * sizeof(double), d->stream));
    ECT_CUDA(cudaMalloc(&d->legm, P.nump * sizeof(EctLegM)));
    ECT_CUDA(cudaMemcpyAsync(d->legm, d->h_legm.data(), P.nump * sizeof(EctLegM), cudaMemcpyHostToDevice, d->stream));
    EctSupolfM* d_cm = nullptr;
    ECT_CUDA(cudaMalloc(&d_cm, P.nump * sizeof(EctSupolfM)));
    ECT_CUDA(cudaMemcpyAsync(d_cm, cms.data(), P.nump * sizeof(EctSupolfM), cudaMemcpyHostToDevice, d->stream));
    double* d_rmu = nullptr;
    ECT_CUDA(cudaMalloc(&d_rmu, P.ndgl * sizeof(double)));
    ECT_CUDA(cudaMemcpyAsync(d_rmu, P.rmu.data(), P.ndgl * sizeof(double), cudaMemcpyHostToDevice, d->stream));
    int maxdglu = 0;
    for (auto& lm : d->h_legm) maxdglu = std::max(maxdglu, lm.ndglu);
    if (maxdglu > 0 && !h->defer_table) {
        dim3 grid((maxdglu + 127) / 128, P.nump, 2);
        k_supolf_table<<<grid, 128, 0, d->stream>>>(d->legm, d_cm, d_rmu, d->ptab, T);
    }
    ECT_CUDA(cudaGetLastError());
    ECT_CUDA(cudaStreamSynchronize(d->stream));
    cudaFree(d_cm);
    cudaFree(d_rmu);
    // record tables
    const size_t nrow = (size_t)P.mrow0[P.nump];
    ECT_CUDA(cudaMalloc(&d->leg_rec_n, std::max<size_t>(nrow, 1) * sizeof(int)));
    ECT_CUDA(cudaMalloc(&d->leg_rec_s, std::max<size_t>(nrow, 1) * sizeof(int)));
    ECT_CUDA(cudaMemcpy(d->leg_rec_n, P.leg_rec_n.data(), nrow * sizeof(int), cudaMemcpyHostToDevice));
    ECT_CUDA(cudaMemcpy(d->leg_rec_s, P.leg_rec_s.data(), nrow * sizeof(int), cudaMemcpyHostToDevice));
    std::vector<int> nasm0(P.nump);
    for (int ml = 0; ml < P.nump; ++ml) nasm0[ml] = P.nasm0[P.myms[ml]];
    ECT_CUDA(cudaMalloc(&d->nasm0, P.nump * sizeof(int)));
    ECT_CUDA(cudaMemcpy(d->nasm0, nasm0.data(), P.nump * sizeof(int), cudaMemcpyHostToDevice));
    // tile schedules, heaviest (smallest m) first
    std::vector<int> order(P.nump);
    for (int i = 0; i < P.nump; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return P.myms[x] < P.myms[y]; });
    std::vector<int2> inv, dir;
    for (int ml : order) {
        const EctLegM& lm = d->h_legm[ml];
        if (lm.ndglu == 0) continue;
        for (int it = 0; it < (lm.ndglu + LEG_BM - 1) / LEG_BM; ++it) inv.push_back(make_int2(ml, it));
        for (int kt = 0; kt < (lm.ils + LEG_BM - 1) / LEG_BM; ++kt) dir.push_back(make_int2(ml, kt));
    }
    d->n_inv_tiles = (int)inv.size();
    d->n_dir_tiles = (int)dir.size();
    d->n_inv_tiles_m0 = d->n_dir_tiles_m0 = 0;
    for (const int2& t2 : inv) if (d->h_legm[t2.x].m == 0) d->n_inv_tiles_m0++;
    for (const int2& t2 : dir) if (d->h_legm[t2.x].m == 0) d->n_dir_tiles_m0++;
    ECT_CUDA(cudaMalloc(&d->inv_tiles, std::max<size_t>(inv.size(), 1) * sizeof(int2)));
    ECT_CUDA(cudaMalloc(&d->dir_tiles, std::max<size_t>(dir.size(), 1) * sizeof(int2)));
    ECT_CUDA(cudaMemcpy(d->inv_tiles, inv.data(), inv.size() * sizeof(int2), cudaMemcpyHostToDevice));
    ECT_CUDA(cudaMemcpy(d->dir_tiles, dir.data(), dir.size() * sizeof(int2), cudaMemcpyHostToDevice));
    return ECT_SUCCESS;
}

// debug / test access to the table: copies P[k][i] of local m index ml, parity par to host
int ect_legendre_get_table(EctHandle* h, int ml, int par, double* out, long long cap) {
    EctDevice* d = h->d;
    if (ml < 0 || ml >= (int)d->h_legm.size()) return ECT_ERR_BADARG;
    const EctLegM& lm = d->h_legm[ml];
    const int k = par ? lm.ila : lm.ils;
    if ((long long)k * lm.ndglu > cap) return ECT_ERR_BADARG;
    if (k == 0 || lm.ndglu == 0) return ECT_SUCCESS;
    ECT_CUDA(cudaMemcpy2D(out, lm.ndglu * sizeof(double), d->ptab + (par ? lm.pa_off : lm.ps_off),
                          lm.ldp * sizeof(double), lm.ndglu * sizeof(double), k, cudaMemcpyDeviceToHost));
    return ECT_SUCCESS;
}

int ect_legendre_set_table(EctHandle* h, int ml, int par, const double* in) {
    EctDevice* d = h->d;
    if (ml < 0 || ml >= (int)d->h_legm.size()) return ECT_ERR_BADARG;
    const EctLegM& lm = d->h_legm[ml];
    const int k = par ? lm.ila : lm.ils;
    if (k == 0 || lm.ndglu == 0) return ECT_SUCCESS;
    ECT_CUDA(cudaMemcpy2D(d->ptab + (par ? lm.pa_off : lm.ps_off), lm.ldp * sizeof(double), in,
                          lm.ndglu * sizeof(double), lm.ndglu * sizeof(double), k, cudaMemcpyHostToDevice));
    ect_tc_invalidate(h);          // the hi / lo float tables of the tensor-core path are rebuilt at the next sp transform
    return ECT_SUCCESS;
}
