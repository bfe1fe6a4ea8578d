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

// ------------------------------------------------------------------------------------------
// small PTX helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& d0, double& d1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(s), "l"(gmem), "r"(sz));
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
    int kf_uv, kf_sc, scders, vorgp, divgp;
};

__device__ __forceinline__ double d_eps(int m, int n) {     // pre_suleg_mod.F90:46-65
    const double dn = (double)n, dm = (double)m;
    return sqrt((dn * dn - dm * dm) / (4.0 * dn * dn - 1.0));
}
__device__ __forceinline__ double d_lap(int n) {            // RLAPIN
    return n >= 1 ? -(ECT_RA * ECT_RA / ((double)n * (double)(n + 1))) : 0.0;
}
__device__ __forceinline__ double2 ld_spec(const EctSpecField f, int idx, bool valid, bool m0) {
    if (!valid) return make_double2(0.0, 0.0);
    const double re = f.base[(long long)idx * f.stride];
    const double im = m0 ? 0.0 : f.base[(long long)(idx + 1) * f.stride];
    return make_double2(re, im);
}

__global__ void k_ltinv_prologue(ProArgs a) {
    const EctLegM lm = a.legm[blockIdx.y];
    const int m = lm.m, T = a.nsmax;
    const int r = blockIdx.x;
    if (r > T - m + 1) return;
    const int n = m + r;
    __shared__ double cst[5];
    if (threadIdx.x == 0) {
        cst[0] = (double)m * d_lap(n);                                    // zkm * lapin(n)
        cst[1] = ((double)(n - 1) * d_eps(m, n)) * d_lap(n - 1);           // c1
        cst[2] = ((double)(n + 2) * d_eps(m, n + 1)) * d_lap(n + 1);       // c2
        cst[3] = (double)(n - 1) * d_eps(m, n);                            // spnsde
        cst[4] = (double)(n + 2) * d_eps(m, n + 1);
    }
    __syncthreads();
    const double zl = cst[0], c1 = cst[1], c2 = cst[2], e1 = cst[3], e2 = cst[4];
    const int base = a.nasm0[blockIdx.y];
    const int idx = base + 2 * r;
    const bool m0 = (m == 0);
    const bool v0 = (n <= T), vm = (n - 1 >= m), vp = (n + 1 <= T);
    double* row = a.x + (lm.xrow0 + r) * (long long)a.cp;
    const int o_vor = 0, o_div = a.vorgp ? a.kf_uv : 0;
    const int o_u = o_div + (a.divgp ? a.kf_uv : 0), o_v = o_u + a.kf_uv;
    const int o_sc = o_v + a.kf_uv, o_nsd = o_sc + a.kf_sc;
    for (int j = threadIdx.x; j < a.kf_uv + a.kf_sc; j += blockDim.x) {
        if (j < a.kf_uv) {
            const EctSpecField fv = a.vor[j], fd = a.div[j];
            const double2 z0 = ld_spec(fv, idx, v0, m0), zm = ld_spec(fv, idx - 2, vm, m0), zp = ld_spec(fv, idx + 2, vp, m0);
            const double2 d0 = ld_spec(fd, idx, v0, m0), dm = ld_spec(fd, idx - 2, vm, m0), dp = ld_spec(fd, idx + 2, vp, m0);
            // vdtuv_mod.F90:121-139
            double2 u, v;
            u.x = -zl * d0.y + c1 * zm.x - c2 * zp.x;
            u.y = zl * d0.x + c1 * zm.y - c2 * zp.y;
            v.x = -zl * z0.y - c1 * dm.x + c2 * dp.x;
            v.y = zl * z0.x - c1 * dm.y + c2 * dp.y;
            if (m0) { u.y = 0.0; v.y = 0.0; }
            if (a.vorgp) *reinterpret_cast<double2*>(row + 2 * (o_vor + j)) = z0;
            if (a.divgp) *reinterpret_cast<double2*>(row + 2 * (o_div + j)) = d0;
            *reinterpret_cast<double2*>(row + 2 * (o_u + j)) = u;
            *reinterpret_cast<double2*>(row + 2 * (o_v + j)) = v;
        } else {
            const int s = j - a.kf_uv;
            const EctSpecField f = a.sc[s];
            const double2 f0 = ld_spec(f, idx, v0, m0);
            *reinterpret_cast<double2*>(row + 2 * (o_sc + s)) = f0;
            if (a.scders) {
                const double2 fm = ld_spec(f, idx - 2, vm, m0), fp = ld_spec(f, idx + 2, vp, m0);
                *reinterpret_cast<double2*>(row + 2 * (o_nsd + s)) =
                    make_double2(-e1 * fm.x + e2 * fp.x, -e1 * fm.y + e2 * fp.y);
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
    if (h->hp.nump == 0) return;
    dim3 grid(h->hp.nsmax + 2, h->hp.nump);
    int items = f.kf_uv + f.kf_sc;
    int threads = items >= 192 ? 256 : (items >= 96 ? 128 : 64);
    k_ltinv_prologue<<<grid, threads, 0, d->stream>>>(a);
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
};

__global__ void k_ltdir_epilogue(EpiArgs a) {
    const EctLegM lm = a.legm[blockIdx.y];
    const int m = lm.m, T = a.nsmax;
    const int r = blockIdx.x;
    if (r > T - m) return;
    const int n = m + r;
    __shared__ double cst[2];
    if (threadIdx.x == 0) {
        cst[0] = (double)n * d_eps(m, n + 1);          // ZN(JN)*PEPSNM(JN+1)
        cst[1] = (double)(n + 1) * d_eps(m, n);        // ZN(JN+1)*PEPSNM(JN)
    }
    __syncthreads();
    const double c1 = cst[0], c2 = cst[1], zkm = (double)m;
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
            const EctSpecField fv = a.vor[j], fd = a.div[j];
            double* bv = const_cast<double*>(fv.base);
            double* bd = const_cast<double*>(fd.base);
            bv[(long long)idx * fv.stride] = vor.x;
            bv[(long long)(idx + 1) * fv.stride] = vor.y;
            bd[(long long)idx * fd.stride] = div.x;
            bd[(long long)(idx + 1) * fd.stride] = div.y;
        } else {
            const int s = j - a.kf_uv;
            double2 f0 = *reinterpret_cast<const double2*>(row + 2 * (o_sc + s));
            if (m0) f0.y = 0.0;
            const EctSpecField f = a.sc[s];
            double* b = const_cast<double*>(f.base);
            b[(long long)idx * f.stride] = f0.x;
            b[(long long)(idx + 1) * f.stride] = f0.y;
        }
    }
}

void ect_launch_ltdir_epilogue(EctHandle* h, const EctFieldCfg& f, void* d_vor, void* d_div, void* d_sc) {
    EctDevice* d = h->d;
    EpiArgs a;
    a.legm = d->legm; a.nasm0 = d->nasm0;
    a.vor = (EctSpecField*)d_vor; a.div = (EctSpecField*)d_div; a.sc = (EctSpecField*)d_sc;
    a.x = d->xwork; a.cp = f.cp; a.nsmax = h->hp.nsmax;
    a.kf_uv = f.kf_uv; a.kf_sc = f.kf_sc;
    if (h->hp.nump == 0) return;
    dim3 grid(h->hp.nsmax + 1, h->hp.nump);
    int items = f.kf_uv + f.kf_sc;
    int threads = items >= 192 ? 256 : (items >= 96 ? 128 : 64);
    k_ltdir_epilogue<<<grid, threads, 0, d->stream>>>(a);
    d->launches++;
}

// ------------------------------------------------------------------------------------------
// DMMA contraction kernels.  CTA = 256 threads = 8 warps as 2 (M) x 4 (N); CTA tile 64 x 128,
// warp tile 32 x 32 for each parity -> 2 x 16 m8n8 accumulators per thread.
// Shared-memory pitches are = 4 (mod 16) doubles so that the fragment loads
// (row = lane/4, k = lane%4) hit 16 distinct 8-byte banks per half warp.
// ------------------------------------------------------------------------------------------
#define LEG_BM 64
#define LEG_BN 128
#define INV_KC 8
#define INV_STAGES 4
#define INV_LDA (LEG_BM + 4)     // 68
#define INV_LDB (LEG_BN + 4)     // 132
#define INV_STAGE_DOUBLES (2 * INV_KC * INV_LDA + 2 * INV_KC * INV_LDB)
#define DIR_KC 8
#define DIR_NB (DIR_KC / 4)     // register-staged double2 per thread and hemisphere
#define DIR_LDA (DIR_KC + 4)     // 20
#define DIR_LDB (LEG_BN + 4)
#define DIR_STAGE_DOUBLES (2 * LEG_BM * DIR_LDA + 2 * DIR_KC * DIR_LDB)

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
};

__global__ void __launch_bounds__(256, 1) k_leinv(LegArgs a) {
    extern __shared__ __align__(16) double smem[];
    const int tile = blockIdx.x / a.nct, ct = blockIdx.x - tile * a.nct;
    const int2 td = a.tiles[tile];
    const EctLegM lm = a.legm[td.x];
    const int i0 = td.y * LEG_BM, c0 = ct * LEG_BN;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    const int nchunks = (lm.ils + INV_KC - 1) / INV_KC;
    const double* ps = a.ptab + lm.ps_off + i0;
    const double* pa = a.ptab + lm.pa_off + i0;
    const double* xb = a.x + lm.xrow0 * (long long)a.cp + c0;

    auto load_chunk = [&](int chunk, int buf) {
        double* As = smem + (size_t)buf * INV_STAGE_DOUBLES;
        double* Aa = As + INV_KC * INV_LDA;
        double* Bs = Aa + INV_KC * INV_LDA;
        double* Ba = Bs + INV_KC * INV_LDB;
        const int k0 = chunk * INV_KC;
        {   // polynomial tiles: 8 rows x 64 latitudes per parity, one 16 B piece per thread
            const int row = tid >> 5, c2 = (tid & 31) * 2;
            const int k = k0 + row;
            const bool vs = k < lm.ils, va = k < lm.ila;
            cp_async16(As + row * INV_LDA + c2, ps + (long long)(vs ? k : 0) * lm.ldp + c2, vs);
            cp_async16(Aa + row * INV_LDA + c2, pa + (long long)(va ? k : 0) * lm.ldp + c2, va);
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {   // spectral tiles: 8 rows x 128 columns per parity
            const int idx = tid + e * 256;
            const int row = idx >> 6, c2 = (idx & 63) * 2;
            const int k = k0 + row;
            const bool cv = (c0 + c2) < a.cp;
            const bool vs = cv && k < lm.ils, va = cv && k < lm.ila;
            cp_async16(Bs + row * INV_LDB + c2, xb + (long long)(vs ? 2 * k : 0) * a.cp + (cv ? c2 : 0), vs);
            cp_async16(Ba + row * INV_LDB + c2, xb + (long long)(va ? 2 * k + 1 : 0) * a.cp + (cv ? c2 : 0), va);
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
        const double* Aa = As + INV_KC * INV_LDA;
        const double* Bs = Aa + INV_KC * INV_LDA;
        const double* Ba = Bs + INV_KC * INV_LDB;
#pragma unroll
        for (int kk = 0; kk < INV_KC; kk += 4) {
            double fs[4], fa[4], bs[4], ba[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                fs[i] = As[(kk + t) * INV_LDA + wm * 32 + i * 8 + g];
                fa[i] = Aa[(kk + t) * INV_LDA + wm * 32 + i * 8 + g];
                bs[i] = Bs[(kk + t) * INV_LDB + wn * 32 + i * 8 + g];
                ba[i] = Ba[(kk + t) * INV_LDB + wn * 32 + i * 8 + g];
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
    cp_async_wait<0>();
    // epilogue: north = S + A, south = S - A  (asre1b_mod.F90:99-100)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int li = i0 + wm * 32 + i * 8 + g;
        if (li >= lm.ndglu) continue;
        const long long rn = (long long)a.rec_n[lm.rec0 + li] * a.cp;
        const long long rs = (long long)a.rec_s[lm.rec0 + li] * a.cp;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + wn * 32 + j * 8 + 2 * t;
            if (c >= a.cp) continue;
            *reinterpret_cast<double2*>(a.fb + rn + c) =
                make_double2(acs[i][j][0] + aca[i][j][0], acs[i][j][1] + aca[i][j][1]);
            *reinterpret_cast<double2*>(a.fb + rs + c) =
                make_double2(acs[i][j][0] - aca[i][j][0], acs[i][j][1] - aca[i][j][1]);
        }
    }
}

__global__ void __launch_bounds__(256, 1) k_ledir(LegArgs a) {
    extern __shared__ __align__(16) double smem[];
    const int tile = blockIdx.x / a.nct, ct = blockIdx.x - tile * a.nct;
    const int2 td = a.tiles[tile];
    const EctLegM lm = a.legm[td.x];
    const int kr0 = td.y * LEG_BM, c0 = ct * LEG_BN;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    const int nchunks = (lm.ndglu + DIR_KC - 1) / DIR_KC;
    const double* ps = a.ptab + lm.ps_off;
    const double* pa = a.ptab + lm.pa_off;

    double2 rn_[DIR_NB], rs_[DIR_NB];     // register-staged north / south records of the next chunk

    auto load_a = [&](int chunk, int buf) {      // polynomial tile P[k][lat chunk]: 64 rows x DIR_KC latitudes per parity
        double* As = smem + (size_t)buf * DIR_STAGE_DOUBLES;
        double* Aa = As + LEG_BM * DIR_LDA;
        const int l0 = chunk * DIR_KC;
#pragma unroll
        for (int e = 0; e < DIR_KC / 8; ++e) {
            const int idx = tid + e * 256;
            const int row = idx / (DIR_KC / 2), c2 = (idx % (DIR_KC / 2)) * 2;
            const int k = kr0 + row;
            const bool vs = k < lm.ils, va = k < lm.ila;
            cp_async16(As + row * DIR_LDA + c2, ps + (long long)(vs ? k : 0) * lm.ldp + l0 + c2, vs);
            cp_async16(Aa + row * DIR_LDA + c2, pa + (long long)(va ? k : 0) * lm.ldp + l0 + c2, va);
        }
    };
    auto load_b_regs = [&](int chunk) {
        const int l0 = chunk * DIR_KC;
#pragma unroll
        for (int e = 0; e < DIR_NB; ++e) {
            const int idx = tid + e * 256;
            const int lr = idx >> 6, c2 = (idx & 63) * 2;
            const int li = l0 + lr;
            if (li < lm.ndglu && (c0 + c2) < a.cp) {
                const long long on = (long long)a.rec_n[lm.rec0 + li] * a.cp + c0 + c2;
                const long long os = (long long)a.rec_s[lm.rec0 + li] * a.cp + c0 + c2;
                rn_[e] = *reinterpret_cast<const double2*>(a.fb + on);
                rs_[e] = *reinterpret_cast<const double2*>(a.fb + os);
            } else {
                rn_[e] = make_double2(0.0, 0.0);
                rs_[e] = make_double2(0.0, 0.0);
            }
        }
    };
    auto store_b = [&](int chunk, int buf) {
        double* Bs = smem + (size_t)buf * DIR_STAGE_DOUBLES + 2 * LEG_BM * DIR_LDA;
        double* Ba = Bs + DIR_KC * DIR_LDB;
        const int l0 = chunk * DIR_KC;
#pragma unroll
        for (int e = 0; e < DIR_NB; ++e) {
            const int idx = tid + e * 256;
            const int lr = idx >> 6, c2 = (idx & 63) * 2;
            const int li = l0 + lr;
            double w = 0.0, ra = 1.0;
            if (li < lm.ndglu) { w = a.rw[lm.isl + li]; ra = a.racthe[lm.isl + li]; }
            double2 s = make_double2(rn_[e].x + rs_[e].x, rn_[e].y + rs_[e].y);    // prfi2b_mod.F90:91-92
            double2 d = make_double2(rn_[e].x - rs_[e].x, rn_[e].y - rs_[e].y);
            if (c0 + c2 < a.c_uv_end) { s.x *= ra; s.y *= ra; d.x *= ra; d.y *= ra; }   // ldfou2_mod.F90:90-96
            s.x *= w; s.y *= w; d.x *= w; d.y *= w;                                  // ledir_mod.F90:122
            *reinterpret_cast<double2*>(Bs + lr * DIR_LDB + c2) = s;
            *reinterpret_cast<double2*>(Ba + lr * DIR_LDB + c2) = d;
        }
    };

    double acs[4][4][2], aca[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acs[i][j][0] = acs[i][j][1] = 0.0; aca[i][j][0] = aca[i][j][1] = 0.0; }

    if (nchunks > 0) {
        load_a(0, 0);
        cp_async_commit();
        load_b_regs(0);
        store_b(0, 0);
        cp_async_wait<0>();
    }
    __syncthreads();
    for (int ch = 0; ch < nchunks; ++ch) {
        const int buf = ch & 1;
        const bool more = ch + 1 < nchunks;
        if (more) {
            load_a(ch + 1, buf ^ 1);
            cp_async_commit();
            load_b_regs(ch + 1);
        }
        const double* As = smem + (size_t)buf * DIR_STAGE_DOUBLES;
        const double* Aa = As + LEG_BM * DIR_LDA;
        const double* Bs = Aa + LEG_BM * DIR_LDA;
        const double* Ba = Bs + DIR_KC * DIR_LDB;
#pragma unroll
        for (int kk = 0; kk < DIR_KC; kk += 4) {
            double fs[4], fa[4], bs[4], ba[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                fs[i] = As[(wm * 32 + i * 8 + g) * DIR_LDA + kk + t];
                fa[i] = Aa[(wm * 32 + i * 8 + g) * DIR_LDA + kk + t];
                bs[i] = Bs[(kk + t) * DIR_LDB + wn * 32 + i * 8 + g];
                ba[i] = Ba[(kk + t) * DIR_LDB + wn * 32 + i * 8 + g];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dmma884(acs[i][j][0], acs[i][j][1], fs[i], bs[j]);
                    dmma884(aca[i][j][0], aca[i][j][1], fa[i], ba[j]);
                }
        }
        if (more) {
            store_b(ch + 1, buf ^ 1);
            cp_async_wait<0>();
        }
        __syncthreads();
    }
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

static bool g_leg_attr_set = false;
static void leg_set_attrs() {
    if (g_leg_attr_set) return;
    cudaFuncSetAttribute(k_leinv, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         INV_STAGES * INV_STAGE_DOUBLES * (int)sizeof(double));
    cudaFuncSetAttribute(k_ledir, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         2 * DIR_STAGE_DOUBLES * (int)sizeof(double));
    g_leg_attr_set = true;
}

void ect_launch_leinv(EctHandle* h, const EctFieldCfg& f) {
    EctDevice* d = h->d;
    if (d->n_inv_tiles == 0) return;
    leg_set_attrs();
    LegArgs a;
    a.legm = d->legm; a.tiles = d->inv_tiles; a.nct = (f.cp + LEG_BN - 1) / LEG_BN;
    a.ptab = d->ptab; a.x = d->xwork; a.fb = d->fbuf_leg;
    a.rec_n = d->leg_rec_n; a.rec_s = d->leg_rec_s; a.rw = d->rw; a.racthe = d->racthe;
    a.cp = f.cp; a.c_uv_end = 0;
    const size_t smem = INV_STAGES * INV_STAGE_DOUBLES * sizeof(double);
    k_leinv<<<(unsigned)((long long)d->n_inv_tiles * a.nct), 256, smem, d->stream>>>(a);
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
    const size_t smem = 2 * DIR_STAGE_DOUBLES * sizeof(double);
    k_ledir<<<(unsigned)((long long)d->n_dir_tiles * a.nct), 256, smem, d->stream>>>(a);
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
    if (maxdglu > 0) {
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
