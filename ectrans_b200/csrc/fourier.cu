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
#include <cstdlib>

#define FT_PAIRS_PER_CTA 8

// phase timing probe (debug): cycles of block 0 at phase boundaries of its first pair
__device__ long long g_ft_probe[64];
#define FT_PROBE(i) do { if (blockIdx.x == 0 && tid == 0 && p == p0 + ((a.dbg >> 8) & 15)) g_ft_probe[i] = clock64(); } while (0)

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
    // direct store destinations: rank owning m + record in its Legendre-side buffer (TRLTOM fused)
    double* const* peer; const int* dst_rank; const int* dst_rec;
    const double* rw_loc;         // Gaussian weight per local latitude (direct: folded into the stored records)
    int n_uv_fields;              // direct: fields < n_uv_fields are u, v (also scaled by 1/(a cos theta), LDFOU2)
    int fp32;                     // grid-point arrays are float
    int dbg;                      // debug switches (ECT_FFT_DBG): 1 no output chirp, 2 no stores, 4 no middle kernel spectrum
};

__device__ __forceinline__ i64 gp_index(int g, int nproma, i64 blkstride) {
    const int blk = g / nproma;
    return (i64)blk * blkstride + (g - blk * nproma);
}

__device__ __forceinline__ void ft_cp_async16(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(sa), "l"(gmem));
}
__device__ __forceinline__ void ft_cp_async8(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" :: "r"(sa), "l"(gmem));
}
__device__ __forceinline__ void ft_cp_async4(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" :: "r"(sa), "l"(gmem));
}
__device__ __forceinline__ void ft_cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void ft_cp_wait_all() { asm volatile("cp.async.wait_all;\n" ::); }

// Shared memory of one CTA:
//   data [ECT_PADDED_LEN(len)] double2   work array of the pair in flight
//   stage                                raw inputs of the NEXT pair, filled by cp.async while this pair is transformed
//                                        inverse: (km+1) x {field a, field b} double2 ; direct: 2 x nlon doubles
//   t1, t2                               two-level twiddle table
//   roots                                odd-radix root tables
template <bool INVERSE, int MAXR, int TB, bool FP32>
__global__ void __launch_bounds__(TB) k_fourier(FtArgs a) {
    extern __shared__ __align__(16) double2 sm[];
    __shared__ EctFftPlan s_plan;        // stage list indexed at run time: keep it out of local memory
    constexpr int NROOTS = MAXR <= 7 ? ECT_ROOTS_OFF(8) : ECT_ROOTS_SIZE;
    const int item = blockIdx.x;
    const int l = a.lats[item / a.nchunks];
    const int chunk = item % a.nchunks;
    const EctLatPlan lp = a.latplans[a.lat_plan[l]];
    EctPairCtx c;
    c.nlon = lp.nlon; c.km = lp.km; c.racthe = a.racthe_loc[l];
    if (threadIdx.x == 0) s_plan = a.plans[lp.plan];
    __syncthreads();
    const int plan_n = s_plan.n, plan_nst = s_plan.nst;
    c.perm = a.perm_pool + s_plan.perm_off;
    c.bluestein = lp.bluestein; c.m = lp.m;
    c.chirp = lp.bluestein ? a.cz_pool + lp.chirp_off : nullptr;
    c.bhat = lp.bluestein ? a.cz_pool + (INVERSE ? lp.bhat_inv_off : lp.bhat_dir_off) : nullptr;
    c.rec = a.fft_rec + a.latrow0[l];
    c.cp = a.cp;
    const int len = plan_n;
    const int N = c.nlon, km = c.km;
    double2* data = sm;
    double2* stage = data + ECT_PADDED_LEN(len);
    // inverse chirp-z rows keep the chirp c[0 .. N/2] in shared memory behind the staged records (both fit in
    // the space the direct transform needs for its two staged rows whenever 2 (km+1) + N/2 + 1 <= N)
    const bool chirp_sm = INVERSE && lp.bluestein;
    const int nstage = INVERSE ? 2 * (km + 1) + (chirp_sm ? N / 2 + 1 : 0) : N;         // double2 elements
    double2* s_chirp = stage + 2 * (km + 1);
    double2* t1 = stage + nstage;
    double2* t2 = t1 + ECT_TW1_LEN(len);
    double2* s_roots = t2 + ECT_TW2_LEN;
    int* s_rec = reinterpret_cast<int*>(s_roots + NROOTS);   // per m: inverse local record, direct (dest rank << 24 | dest record)
    const int tid = threadIdx.x, nthr = blockDim.x;
    tw_build(t1, t2, a.tw_pool + s_plan.tw_off, len, tid, nthr);
    for (int j = tid; j < NROOTS; j += nthr) s_roots[j] = a.roots[j];
    if (chirp_sm) for (int j = tid; j <= N / 2; j += nthr) s_chirp[j] = c.chirp[j];
    for (int k = tid; k <= km; k += nthr)
        s_rec[k] = INVERSE ? c.rec[k] : ((a.dst_rank[a.latrow0[l] + k] << 24) | a.dst_rec[a.latrow0[l] + k]);
    const double2* chirp_tab = chirp_sm ? s_chirp : c.chirp;
    const EctTw qt{t1, t2};
    c.qt = qt; c.roots = s_roots;
    const int g0 = a.gpoff[l];
    const bool oneblk = a.nproma >= a.ngptot;
    const int p0 = chunk * FT_PAIRS_PER_CTA, p1 = min(p0 + FT_PAIRS_PER_CTA, a.npairs);
    constexpr int NB = 4;      // global loads issued per thread before the first use (latency batching)
    const double s1 = c.racthe, s2 = c.racthe * c.racthe;

    auto prefetch = [&](int p) {       // raw inputs of pair p -> stage (asynchronous)
        const int2 pr = a.pairs[p];
        const int fa = pr.x, fb2 = pr.y;
        if (INVERSE) {
            const int ca = a.fsf[fa].src_c;
            const int cb = fb2 >= 0 ? a.fsf[fb2].src_c : -1;
            for (int k0 = tid; k0 <= km; k0 += NB * nthr) {
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const int k = k0 + i * nthr;
                    if (k > km) continue;
                    const double* src = a.fb + (long long)s_rec[k] * c.cp;
                    ft_cp_async16(stage + 2 * k, src + ca);
                    if (cb >= 0) ft_cp_async16(stage + 2 * k + 1, src + cb);
                }
            }
        } else {
            const double* ba = a.gp_base[fa]; const i64 sa = a.gp_blk[fa];
            const double* bb = fb2 >= 0 ? a.gp_base[fb2] : nullptr; const i64 sb = fb2 >= 0 ? a.gp_blk[fb2] : 0;
            if (FP32) {
                float* st = reinterpret_cast<float*>(stage);
                const float* fa_ = reinterpret_cast<const float*>(ba); const float* fb_ = reinterpret_cast<const float*>(bb);
                for (int j = tid; j < N; j += nthr) {
                    const int g = g0 + j;
                    ft_cp_async4(st + j, fa_ + (oneblk ? (i64)g : gp_index(g, a.nproma, sa)));
                    if (bb) ft_cp_async4(st + N + j, fb_ + (oneblk ? (i64)g : gp_index(g, a.nproma, sb)));
                }
            } else {
                double* st = reinterpret_cast<double*>(stage);
                for (int j = tid; j < N; j += nthr) {
                    const int g = g0 + j;
                    ft_cp_async8(st + j, ba + (oneblk ? (i64)g : gp_index(g, a.nproma, sa)));
                    if (bb) ft_cp_async8(st + N + j, bb + (oneblk ? (i64)g : gp_index(g, a.nproma, sb)));
                }
            }
        }
        ft_cp_commit();
    };

    __syncthreads();                    // s_rec visible
    if (p0 < p1) prefetch(p0);
    for (int p = p0; p < p1; ++p) {
        const int2 pr = a.pairs[p];
        const int fa = pr.x, fb2 = pr.y;
        const bool hasb = fb2 >= 0;
        FT_PROBE(0);
        ft_cp_wait_all();
        __syncthreads();                 // staged inputs (and, first time, the twiddle tables) are visible
        if (INVERSE) {
            // FOURIER_IN + FSC + zero padding (same arithmetic as fourier_phases.h ftinv_load)
            EctFsField sfa = a.fsf[fa], sfb;
            if (hasb) sfb = a.fsf[fb2]; else { sfb.src_c = -1; sfb.pw = 0; sfb.deriv = 0; }
            if (!c.bluestein) { for (int k = km + 1 + tid; k < N - km; k += nthr) data[ECT_PAD((int)c.perm[k])] = make_double2(0.0, 0.0); }
            else { for (int u = 2 * km + 1 + tid; u < c.m; u += nthr) data[ECT_PAD(u)] = make_double2(0.0, 0.0); }
            const double sa_ = sfa.pw == 0 ? 1.0 : (sfa.pw == 1 ? s1 : s2);
            const double sb_ = sfb.pw == 0 ? 1.0 : (sfb.pw == 1 ? s1 : s2);
            for (int k0 = tid; k0 <= km; k0 += NB * nthr) {
                double2 ch[NB]; int pk[NB], pn[NB];
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const int k = k0 + i * nthr;
                    ch[i] = make_double2(1.0, 0.0); pk[i] = pn[i] = 0;
                    if (k <= km) {
                        if (c.bluestein) ch[i] = chirp_tab[k];
                        else { pk[i] = c.perm[k]; pn[i] = c.perm[k == 0 ? 0 : N - k]; }
                    }
                }
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const int k = k0 + i * nthr;
                    if (k > km) continue;
                    const double2 va = stage[2 * k];
                    const double2 vb = hasb ? stage[2 * k + 1] : make_double2(0.0, 0.0);
                    double2 fa_ = make_double2(va.x * sa_, k == 0 ? 0.0 : va.y * sa_);
                    double2 fb_ = make_double2(vb.x * sb_, k == 0 ? 0.0 : vb.y * sb_);
                    const double z = s1 * (double)k;
                    if (sfa.deriv) fa_ = make_double2(-fa_.y * z, fa_.x * z);
                    if (sfb.deriv) fb_ = make_double2(-fb_.y * z, fb_.x * z);
                    const double2 zp = make_double2(fa_.x - fb_.y, fa_.y + fb_.x);      // Z[k]
                    const double2 zm = make_double2(fa_.x + fb_.y, fb_.x - fa_.y);      // Z[-k]
                    if (!c.bluestein) {
                        data[ECT_PAD(pk[i])] = zp;
                        if (k > 0) data[ECT_PAD(pn[i])] = zm;
                    } else {
                        const double2 xp = c_mul(zp, ch[i]);
                        data[ECT_PAD(km + k)] = make_double2(xp.y, xp.x);
                        if (k > 0) { const double2 xm = c_mul(zm, ch[i]); data[ECT_PAD(km - k)] = make_double2(xm.y, xm.x); }
                    }
                }
            }
        } else {
            const double* st = reinterpret_cast<const double*>(stage);
            const float* stf = reinterpret_cast<const float*>(stage);
            // chirp factors come from global memory (no shared memory left next to two staged rows): the loads of
            // batch i+1 are in flight while batch i is scattered
            auto ld_batch = [&](int j0, double2* ch, int* pj) {
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const int j = j0 + i * nthr;
                    ch[i] = make_double2(1.0, 0.0); pj[i] = 0;
                    if (j < N) {
                        if (c.bluestein) ch[i] = c.chirp[j > N / 2 ? N - j : j];
                        else pj[i] = c.perm[j];
                    }
                }
            };
            auto put_batch = [&](int j0, const double2* ch, const int* pj) {
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const int j = j0 + i * nthr;
                    if (j >= N) continue;
                    const double va = FP32 ? (double)stf[j] : st[j];
                    const double vb = hasb ? (FP32 ? (double)stf[N + j] : st[N + j]) : 0.0;
                    if (!c.bluestein) data[ECT_PAD(pj[i])] = make_double2(vb, va);
                    else { const double2 t = c_mul(make_double2(vb, va), ch[i]); data[ECT_PAD(j)] = make_double2(t.y, t.x); }
                }
            };
            double2 chA[NB], chB[NB]; int pjA[NB], pjB[NB];
            ld_batch(tid, chA, pjA);
            for (int j0 = tid; j0 < N; j0 += 2 * NB * nthr) {
                ld_batch(j0 + NB * nthr, chB, pjB);
                put_batch(j0, chA, pjA);
                ld_batch(j0 + 2 * NB * nthr, chA, pjA);
                put_batch(j0 + NB * nthr, chB, pjB);
            }
            ftdir_zero_tail(data, c, tid, nthr);
        }
        __syncthreads();                 // work array complete, stage consumed
        if (p + 1 < p1) prefetch(p + 1);
        FT_PROBE(1);
        if (c.bluestein) {
            for (int s = plan_nst - 1; s >= 1; --s) {
                fft_stage<true, 7>(data, len, s_plan.radix[s], s_plan.sublen[s], s_plan.lshift[s], qt, s_roots, tid, nthr);
                __syncthreads();
                FT_PROBE(2 + (plan_nst - 1 - s));
            }
            blue_middle(data, len, s_plan.radix[0], (a.dbg & 4) ? a.cz_pool : c.bhat, tid, nthr);
            __syncthreads();
            FT_PROBE(10);
            for (int s = 1; s < plan_nst; ++s) {
                fft_stage<false, 7>(data, len, s_plan.radix[s], s_plan.sublen[s], s_plan.lshift[s], qt, s_roots, tid, nthr);
                __syncthreads();
                FT_PROBE(10 + s);
            }
        } else {
            for (int s = 0; s < plan_nst; ++s) {
                fft_stage<false, MAXR>(data, len, s_plan.radix[s], s_plan.sublen[s], s_plan.lshift[s], qt, s_roots, tid, nthr);
                __syncthreads();
            }
            FT_PROBE(19);
        }
        if (INVERSE) {
            double* ba = a.gp_base[fa]; const i64 sa = a.gp_blk[fa];
            double* bb = hasb ? a.gp_base[fb2] : nullptr; const i64 sb = hasb ? a.gp_blk[fb2] : 0;
            for (int j0 = tid; j0 < N; j0 += NB * nthr) {
                double2 x[NB], ch[NB];
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const int j = j0 + i * nthr;
                    if (j < N) {
                        x[i] = data[ECT_PAD(j)];
                        if (c.bluestein) ch[i] = chirp_tab[j > N / 2 ? N - j : j];
                    }
                }
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const int j = j0 + i * nthr;
                    if (j >= N) continue;
                    const double2 y = c.bluestein ? c_mul(ch[i], x[i]) : x[i];
                    if ((a.dbg & 2) && y.x != 12345.678) continue;
                    const int g = g0 + j;
                    if (FP32) {
                        reinterpret_cast<float*>(ba)[oneblk ? (i64)g : gp_index(g, a.nproma, sa)] = (float)y.x;
                        if (hasb) reinterpret_cast<float*>(bb)[oneblk ? (i64)g : gp_index(g, a.nproma, sb)] = (float)y.y;
                    } else {
                        ba[oneblk ? (i64)g : gp_index(g, a.nproma, sa)] = y.x;
                        if (hasb) bb[oneblk ? (i64)g : gp_index(g, a.nproma, sb)] = y.y;
                    }
                }
            }
        } else {
            // 1/N + FOURIER_OUT (same arithmetic as fourier_phases.h ftdir_store)
            // 1/N (tpm_fftw.F90:317-321), Gaussian weight (ledir_mod.F90:122) and, for u and v, 1/(a cos theta)
            // (ldfou2_mod.F90:90-96) in one factor, so that the Legendre loader only forms N +- S
            const double wl = a.rw_loc[l];
            const double sca = 0.5 / (double)N * wl * (fa < a.n_uv_fields ? s1 : 1.0);
            const double scb = 0.5 / (double)N * wl * (fb2 >= 0 && fb2 < a.n_uv_fields ? s1 : 1.0);
            const int ca = 2 * fa, cb = hasb ? 2 * fb2 : -1;
            for (int k0 = tid; k0 <= km; k0 += NB * nthr) {
                double2 zk[NB], zn[NB], ch[NB]; double* rb[NB];
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const int k = k0 + i * nthr;
                    if (k <= km) {
                        const int pk_ = s_rec[k];
                        rb[i] = a.peer[pk_ >> 24] + (long long)(pk_ & 0xffffff) * c.cp;
                        if (!c.bluestein) { zk[i] = data[ECT_PAD(k)]; zn[i] = data[ECT_PAD(k == 0 ? 0 : N - k)]; }
                        else { zk[i] = data[ECT_PAD(km + k)]; zn[i] = data[ECT_PAD(km - k)]; ch[i] = c.chirp[k]; }
                    }
                }
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const int k = k0 + i * nthr;
                    if (k > km) continue;
                    double2 a_ = zk[i], b_ = zn[i];
                    if (c.bluestein) { a_ = c_mul(ch[i], a_); b_ = c_mul(ch[i], b_); }
                    // stored values are swapped (sign - transform on the sign + core): Z = (y, x)
                    const double2 Zk = make_double2(a_.y, a_.x), Zn = make_double2(b_.y, b_.x);
                    *reinterpret_cast<double2*>(rb[i] + ca) = make_double2((Zk.x + Zn.x) * sca, (Zk.y - Zn.y) * sca);
                    if (cb >= 0) *reinterpret_cast<double2*>(rb[i] + cb) = make_double2((Zk.y + Zn.y) * scb, (Zn.x - Zk.x) * scb);
                }
            }
        }
        __syncthreads();
        FT_PROBE(20);
    }
}

static void fill_args(EctHandle* h, const EctFieldCfg& f, FtArgs& a) {
    EctDevice* d = h->d;
    a.latplans = d->latplans; a.plans = d->plans;
    a.perm_pool = d->perm_pool; a.tw_pool = d->tw_pool; a.cz_pool = d->cz_pool; a.roots = d->roots;
    a.lat_plan = d->lat_plan; a.latrow0 = d->latrow0; a.fft_rec = d->fft_rec;
    a.gpoff = d->gpoff; a.nloen_loc = d->nloen; a.racthe_loc = d->racthe_loc;
    a.fb = d->fbuf_fft; a.cp = f.cp;
    a.peer = d->peer_leg; a.dst_rank = d->fft_dst_rank; a.dst_rec = d->fft_dst_rec;
    a.nfs = f.nfs; a.npairs = f.npairs;
    a.nchunks = (a.npairs + FT_PAIRS_PER_CTA - 1) / FT_PAIRS_PER_CTA;
    a.ngptot = h->hp.ngptot;
    a.fp32 = f.fp32;
    a.rw_loc = d->rw_loc; a.n_uv_fields = 2 * f.kf_uv;
    static const char* dbg = getenv("ECT_FFT_DBG");
    a.dbg = dbg ? atoi(dbg) : 0;
}

template <bool INVERSE>
static void launch_fourier(EctHandle* h, FtArgs& a) {
    EctDevice* d = h->d;
    static const char* only = getenv("ECT_FFT_ONLY_BUCKET");     // debug: run a single shared-memory class
    static const char* serial = getenv("ECT_FFT_SERIAL");        // debug: one stream
    static const char* conc = getenv("ECT_FFT_CONCURRENT");   // side streams did not pay off in measurements: off by default
    const bool fork = conc && atoi(conc) && !(serial && atoi(serial)) && d->ev_fork != nullptr;
    if (fork) {
        cudaEventRecord(d->ev_fork, d->stream);
        for (int k = 0; k < EctDevice::kSide; ++k) cudaStreamWaitEvent(d->side[k], d->ev_fork, 0);
    }
    // largest classes first, round-robin over main + side streams
    std::vector<int> order;
    for (int i = (int)d->buckets.size() - 1; i >= 0; --i) order.push_back(i);
    static const char* oldorder = getenv("ECT_FFT_OLDORDER");
    if (oldorder && atoi(oldorder)) std::reverse(order.begin(), order.end());
    else std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return d->buckets[x].smem > d->buckets[y].smem; });
    int slot = 0;
    for (int bi : order) {
        auto& b = d->buckets[bi];
        if (b.lats.empty()) continue;
        if (only && atoi(only) != bi) continue;
        a.lats = b.d_lats;
        const unsigned grid = (unsigned)(b.lats.size() * (size_t)a.nchunks);
        cudaStream_t st = d->stream;
        if (fork) { const int k = slot % (EctDevice::kSide + 1); st = k == 0 ? d->stream : d->side[k - 1]; }
        ++slot;
        if (a.fp32) {
            if (b.maxr <= 7) k_fourier<INVERSE, 7, 256, true><<<grid, std::min(b.threads, 256), b.smem, st>>>(a);
            else k_fourier<INVERSE, ECT_MAX_RADIX, 256, true><<<grid, b.threads, b.smem, st>>>(a);
        } else if (b.maxr <= 7 && b.threads == 512) k_fourier<INVERSE, 7, 512, false><<<grid, b.threads, b.smem, st>>>(a);
        else if (b.maxr <= 7) k_fourier<INVERSE, 7, 256, false><<<grid, b.threads, b.smem, st>>>(a);
        else k_fourier<INVERSE, ECT_MAX_RADIX, 256, false><<<grid, b.threads, b.smem, st>>>(a);
        d->launches++;
    }
    if (fork)
        for (int k = 0; k < EctDevice::kSide; ++k) {
            cudaEventRecord(d->ev_join[k], d->side[k]);
            cudaStreamWaitEvent(d->stream, d->ev_join[k], 0);
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
    {
        std::vector<double> wl(P.nlat);
        for (int l = 0; l < P.nlat; ++l) wl[l] = P.rw[P.lat0 + l];
        if ((rc = upload(d->rw_loc, wl))) return rc;
    }
    if ((rc = upload(d->gpoff, P.gpoff))) return rc;
    // shared-memory classes
    int maxsm = 0;
    cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, d->dev);
    const int limits[] = {16 * 1024, 32 * 1024, 56 * 1024, 74 * 1024, 112 * 1024, maxsm - 1024};
    const int threads[] = {64, 64, 128, 128, 256, 256};
    d->buckets.clear();
    for (int v = 0; v < 2; ++v)
        for (int i = 0; i < 6; ++i) {
            EctDevice::Bucket b;
            b.smem = limits[i];
            b.threads = v ? std::min(threads[i], 256) : threads[i];
            static const char* t512 = getenv("ECT_FFT_T512");
            if (!v && i >= 4 && t512 && atoi(t512)) b.threads = 512;
            b.maxr = v ? ECT_MAX_RADIX : 7;
            d->buckets.push_back(b);
        }
    std::vector<int> need_of(P.nlat, 0);
    for (int l = 0; l < P.nlat; ++l) {
        const EctLatPlan& lp = d->fft.latplans[d->h_lat_plan[l]];
        const EctFftPlan& pl = d->fft.plans[lp.plan];
        const int len_ = pl.n;
        int need = (ECT_PADDED_LEN(len_) + std::max(2 * (lp.km + 1) + (lp.bluestein ? lp.nlon / 2 + 1 : 0), lp.nlon) + ECT_TW1_LEN(len_) + ECT_TW2_LEN) * (int)sizeof(double2);

        int maxr = 2;
        for (int s = 0; s < pl.nst; ++s) if (pl.radix[s] & 1) maxr = std::max(maxr, pl.radix[s]);   // 2,4,8,16 are in every variant
        need += (maxr <= 7 ? ECT_ROOTS_OFF(8) : ECT_ROOTS_SIZE) * (int)sizeof(double2);
        need += ((lp.km + 1) * (int)sizeof(int) + 15) / 16 * 16;
        need_of[l] = need;
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
        std::sort(b.lats.begin(), b.lats.end(), [&](int x, int y) { return need_of[x] > need_of[y]; });
        if (b.lats.empty()) continue;
        b.smem = need_of[b.lats[0]];
        if ((rc = upload(b.d_lats, b.lats))) return rc;
    }
    ECT_CUDA(cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming));
    for (int k = 0; k < EctDevice::kSide; ++k) {
        ECT_CUDA(cudaStreamCreateWithFlags(&d->side[k], cudaStreamNonBlocking));
        ECT_CUDA(cudaEventCreateWithFlags(&d->ev_join[k], cudaEventDisableTiming));
    }
#define FT_ATTR(...) ECT_CUDA(cudaFuncSetAttribute(k_fourier<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm - 1024))
    FT_ATTR(true, 7, 512, false); FT_ATTR(false, 7, 512, false);
    FT_ATTR(true, 7, 256, false); FT_ATTR(false, 7, 256, false);
    FT_ATTR(true, ECT_MAX_RADIX, 256, false); FT_ATTR(false, ECT_MAX_RADIX, 256, false);
    FT_ATTR(true, 7, 256, true); FT_ATTR(false, 7, 256, true);
    FT_ATTR(true, ECT_MAX_RADIX, 256, true); FT_ATTR(false, ECT_MAX_RADIX, 256, true);
#undef FT_ATTR
    return ECT_SUCCESS;
}

// debug: phase cycle stamps of block 0 (first pair) of the last Fourier launch
extern "C" int ect_debug_fft_probe(long long* out64) {
    ECT_CUDA(cudaDeviceSynchronize());
    ECT_CUDA(cudaMemcpyFromSymbol(out64, g_ft_probe, sizeof(long long) * 64));
    return ECT_SUCCESS;
}
