// Fourier stage kernels for sm_100a: batched mixed-radix (+ chirp-z) real FFTs over the
// reduced grid's per-latitude lengths, two real fields per complex transform, fused with
//   inverse: FOURIER_IN + FSC + zero padding on load, TRLTOG's layout change on store
//            (reference cpu/internal/ftinv_ctl_mod.F90:171-192, fourier_in_mod.F90:58-77,
//             fsc_mod.F90:132-187, ftinv_mod.F90:65-84, trltog_mod.F90:579-731)
//   direct : TRGTOL's layout change on load, 1/N + FOURIER_OUT on store
//            (ftdir_ctl_mod.F90:160-192, ftdir_mod.F90:67-84, fourier_out_mod.F90:58-77)
// One CTA = one latitude x a chunk of field pairs; the work array (nlon or M double2) and the
// twiddle table of that length live in shared memory.
#include "ect_internal.h"
#include "fourier_phases.h"
#include <algorithm>

#define FT_PAIRS_PER_CTA 8

struct FtArgs {
    const EctLatPlan* latplans; const EctFftPlan* plans;
    const uint16_t* perm_pool; const double2* tw_pool; const double2* cz_pool; const double2* roots;
    const int* lat_plan; const i64* latrow0; const int* fft_rec;
    const int* lats;              // latitudes (local index) of this launch
    const int* gpoff; const int* nloen_loc; const double* racthe_loc;
    double* fb; int cp;
    int nfs; int npairs; int nchunks;    // pair chunks per latitude
    double* const* gp_base; const i64* gp_blk;   // per Fourier field
    const EctFsField* fsf;        // inverse only
    const int2* pairs;            // (field a, field b or -1): only fields of one group share a transform
    int nproma; int ngptot;
};

__device__ __forceinline__ i64 gp_index(int g, int nproma, i64 blkstride) {
    const int blk = g / nproma;
    return (i64)blk * blkstride + (g - blk * nproma);
}

template <bool INVERSE, int MAXR>
__global__ void k_fourier(FtArgs a) {
    extern __shared__ __align__(16) double2 sm[];
    __shared__ double2 s_roots[ECT_ROOTS_SIZE];
    const int item = blockIdx.x;
    const int l = a.lats[item / a.nchunks];
    const int chunk = item % a.nchunks;
    const EctLatPlan lp = a.latplans[a.lat_plan[l]];
    EctPairCtx c;
    c.nlon = lp.nlon; c.km = lp.km; c.racthe = a.racthe_loc[l];
    c.plan = a.plans[lp.plan];
    c.perm = a.perm_pool + c.plan.perm_off;
    c.bluestein = lp.bluestein; c.m = lp.m;
    c.chirp = lp.bluestein ? a.cz_pool + lp.chirp_off : nullptr;
    c.bhat = lp.bluestein ? a.cz_pool + (INVERSE ? lp.bhat_inv_off : lp.bhat_dir_off) : nullptr;
    c.rec = a.fft_rec + a.latrow0[l];
    c.cp = a.cp;
    const int len = c.plan.n;
    double2* data = sm;
    const double2* qt = a.tw_pool + c.plan.tw_off;
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int j = tid; j < ECT_ROOTS_SIZE; j += nthr) s_roots[j] = a.roots[j];
    c.qt = qt; c.roots = s_roots;
    const int g0 = a.gpoff[l];
    const bool oneblk = a.nproma >= a.ngptot;
    __syncthreads();
    const int p0 = chunk * FT_PAIRS_PER_CTA, p1 = min(p0 + FT_PAIRS_PER_CTA, a.npairs);
    for (int p = p0; p < p1; ++p) {
        const int2 pr = a.pairs[p];
        const int fa = pr.x, fb2 = pr.y;
        const bool hasb = fb2 >= 0;
        if (INVERSE) {
            EctFsField sfa = a.fsf[fa], sfb;
            if (hasb) sfb = a.fsf[fb2]; else { sfb.src_c = -1; sfb.pw = 0; sfb.deriv = 0; }
            ftinv_load(data, a.fb, c, sfa, sfb, tid, nthr);
        } else {
            const double* ba = a.gp_base[fa]; const i64 sa = a.gp_blk[fa];
            const double* bb = hasb ? a.gp_base[fb2] : nullptr; const i64 sb = hasb ? a.gp_blk[fb2] : 0;
            for (int j = tid; j < c.nlon; j += nthr) {
                const int g = g0 + j;
                const double va = ba[oneblk ? (i64)g : gp_index(g, a.nproma, sa)];
                const double vb = hasb ? bb[oneblk ? (i64)g : gp_index(g, a.nproma, sb)] : 0.0;
                ftdir_put(data, c, j, va, vb);
            }
            ftdir_zero_tail(data, c, tid, nthr);
        }
        __syncthreads();
        if (c.bluestein) {
            for (int s = c.plan.nst - 1; s >= 1; --s) {
                fft_stage<true, 7>(data, len, c.plan.radix[s], c.plan.sublen[s], c.plan.lshift[s], qt, s_roots, tid, nthr);
                __syncthreads();
            }
            blue_middle(data, len, c.plan.radix[0], c.bhat, tid, nthr);
            __syncthreads();
            for (int s = 1; s < c.plan.nst; ++s) {
                fft_stage<false, 7>(data, len, c.plan.radix[s], c.plan.sublen[s], c.plan.lshift[s], qt, s_roots, tid, nthr);
                __syncthreads();
            }
        } else {
            for (int s = 0; s < c.plan.nst; ++s) {
                fft_stage<false, MAXR>(data, len, c.plan.radix[s], c.plan.sublen[s], c.plan.lshift[s], qt, s_roots, tid, nthr);
                __syncthreads();
            }
        }
        if (INVERSE) {
            double* ba = a.gp_base[fa]; const i64 sa = a.gp_blk[fa];
            double* bb = hasb ? a.gp_base[fb2] : nullptr; const i64 sb = hasb ? a.gp_blk[fb2] : 0;
            for (int j = tid; j < c.nlon; j += nthr) {
                const double2 x = ftinv_out(data, c, j);
                const int g = g0 + j;
                ba[oneblk ? (i64)g : gp_index(g, a.nproma, sa)] = x.x;
                if (hasb) bb[oneblk ? (i64)g : gp_index(g, a.nproma, sb)] = x.y;
            }
        } else {
            ftdir_store(data, a.fb, c, 2 * fa, hasb ? 2 * fb2 : -1, tid, nthr);
        }
        __syncthreads();
    }
}

static void fill_args(EctHandle* h, const EctFieldCfg& f, FtArgs& a) {
    EctDevice* d = h->d;
    a.latplans = d->latplans; a.plans = d->plans;
    a.perm_pool = d->perm_pool; a.tw_pool = d->tw_pool; a.cz_pool = d->cz_pool; a.roots = d->roots;
    a.lat_plan = d->lat_plan; a.latrow0 = d->latrow0; a.fft_rec = d->fft_rec;
    a.gpoff = d->gpoff; a.nloen_loc = d->nloen; a.racthe_loc = d->racthe_loc;
    a.fb = d->fbuf_fft; a.cp = f.cp;
    a.nfs = f.nfs; a.npairs = f.npairs;
    a.nchunks = (a.npairs + FT_PAIRS_PER_CTA - 1) / FT_PAIRS_PER_CTA;
    a.ngptot = h->hp.ngptot;
}

template <bool INVERSE>
static void launch_fourier(EctHandle* h, FtArgs& a) {
    EctDevice* d = h->d;
    for (auto& b : d->buckets) {
        if (b.lats.empty()) continue;
        a.lats = b.d_lats;
        const unsigned grid = (unsigned)(b.lats.size() * (size_t)a.nchunks);
        if (b.maxr <= 7) k_fourier<INVERSE, 7><<<grid, b.threads, b.smem, d->stream>>>(a);
        else k_fourier<INVERSE, ECT_MAX_RADIX><<<grid, b.threads, b.smem, d->stream>>>(a);
        d->launches++;
    }
}

void ect_launch_ftinv(EctHandle* h, const EctFieldCfg& f, double* const* d_gp_base, const i64* d_gp_blkstride,
                      const void* d_fsfields, const void* d_pairs, int nproma) {
    if (h->hp.nlat == 0 || f.nfs == 0) return;
    FtArgs a;
    fill_args(h, f, a);
    a.gp_base = d_gp_base; a.gp_blk = d_gp_blkstride; a.fsf = (const EctFsField*)d_fsfields; a.pairs = (const int2*)d_pairs; a.nproma = nproma;
    launch_fourier<true>(h, a);
}

void ect_launch_ftdir(EctHandle* h, const EctFieldCfg& f, double* const* d_gp_base, const i64* d_gp_blkstride,
                      const void* d_pairs, int nproma) {
    if (h->hp.nlat == 0 || f.nfs == 0) return;
    FtArgs a;
    fill_args(h, f, a);
    a.gp_base = d_gp_base; a.gp_blk = d_gp_blkstride; a.fsf = nullptr; a.pairs = (const int2*)d_pairs; a.nproma = nproma;
    launch_fourier<false>(h, a);
}

template <typename T>
static int upload(T*& dptr, const std::vector<T>& v) {
    const size_t n = std::max<size_t>(v.size(), 1);
    ECT_CUDA(cudaMalloc(&dptr, n * sizeof(T)));
    if (!v.empty()) ECT_CUDA(cudaMemcpy(dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return ECT_SUCCESS;
}

int ect_fourier_setup(EctHandle* h) {
    EctHostPlan& P = h->hp;
    EctDevice* d = h->d;
    d->h_lat_plan.resize(P.nlat);
    for (int l = 0; l < P.nlat; ++l) {
        const int g = P.lat0 + l;
        const int id = d->fft.get_latplan(P.nloen[g], P.nmen[g]);
        if (id < 0 || d->fft.latplans[id].plan < 0) {
            ect_set_error("ect_setup: no FFT plan for nlon=%d", P.nloen[g]);
            return ECT_ERR_NOTIMPL;
        }
        d->h_lat_plan[l] = id;
    }
    int rc;
    if ((rc = upload(d->plans, d->fft.plans))) return rc;
    if ((rc = upload(d->latplans, d->fft.latplans))) return rc;
    if ((rc = upload(d->perm_pool, d->fft.perm_pool))) return rc;
    if ((rc = upload(d->tw_pool, d->fft.tw_pool))) return rc;
    if ((rc = upload(d->cz_pool, d->fft.cz_pool))) return rc;
    if (d->fft.roots.empty()) d->fft.roots.assign(ECT_ROOTS_SIZE, make_double2(0.0, 0.0));
    if ((rc = upload(d->roots, d->fft.roots))) return rc;
    if ((rc = upload(d->lat_plan, d->h_lat_plan))) return rc;
    if ((rc = upload(d->latrow0, P.latrow0))) return rc;
    if ((rc = upload(d->fft_rec, P.fft_rec))) return rc;
    std::vector<int> nl(P.nlat);
    std::vector<double> ra(P.nlat);
    for (int l = 0; l < P.nlat; ++l) { nl[l] = P.nloen[P.lat0 + l]; ra[l] = P.racthe[P.lat0 + l]; }
    if ((rc = upload(d->nloen, nl))) return rc;
    if ((rc = upload(d->racthe_loc, ra))) return rc;
    if ((rc = upload(d->gpoff, P.gpoff))) return rc;
    // shared-memory classes
    int maxsm = 0;
    cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, d->dev);
    const int limits[] = {12 * 1024, 24 * 1024, 48 * 1024, 72 * 1024, 108 * 1024, maxsm - 9 * 1024};
    const int threads[] = {64, 64, 128, 128, 128, 256};
    d->buckets.clear();
    for (int v = 0; v < 2; ++v)
        for (int i = 0; i < 6; ++i) {
            EctDevice::Bucket b;
            b.smem = limits[i];
            b.threads = threads[i];
            b.maxr = v ? ECT_MAX_RADIX : 7;
            d->buckets.push_back(b);
        }
    for (int l = 0; l < P.nlat; ++l) {
        const EctLatPlan& lp = d->fft.latplans[d->h_lat_plan[l]];
        const EctFftPlan& pl = d->fft.plans[lp.plan];
        const int need = lp.smem_bytes;
        int maxr = 2;
        for (int s = 0; s < pl.nst; ++s) maxr = std::max(maxr, pl.radix[s]);
        bool placed = false;
        for (auto& b : d->buckets)
            if (need <= b.smem && maxr <= b.maxr) { b.lats.push_back(l); placed = true; break; }
        if (!placed) {
            ect_set_error("ect_setup: latitude with nlon=%d needs %d bytes of shared memory (max %d)",
                          P.nloen[P.lat0 + l], need, maxsm);
            return ECT_ERR_NOTIMPL;
        }
    }
    for (auto& b : d->buckets) {
        // longest rows first; shrink the dynamic allocation to what the class needs
        std::sort(b.lats.begin(), b.lats.end(), [&](int x, int y) {
            return d->fft.latplans[d->h_lat_plan[x]].smem_bytes > d->fft.latplans[d->h_lat_plan[y]].smem_bytes;
        });
        if (b.lats.empty()) continue;
        b.smem = d->fft.latplans[d->h_lat_plan[b.lats[0]]].smem_bytes;
        if ((rc = upload(b.d_lats, b.lats))) return rc;
    }
    ECT_CUDA(cudaFuncSetAttribute(k_fourier<true, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm - 9 * 1024));
    ECT_CUDA(cudaFuncSetAttribute(k_fourier<false, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm - 9 * 1024));
    ECT_CUDA(cudaFuncSetAttribute(k_fourier<true, ECT_MAX_RADIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm - 9 * 1024));
    ECT_CUDA(cudaFuncSetAttribute(k_fourier<false, ECT_MAX_RADIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm - 9 * 1024));
    return ECT_SUCCESS;
}
