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
#include "fourier_cz.h"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#define FT_PAIRS_PER_CTA 8

// phase timing probe (debug): cycles of block 0 at phase boundaries of its first pair
__device__ long long g_ft_probe[64];
#define FT_PROBE(i) do { if (blockIdx.x == 0 && tid == 0 && p == p0 + ((a.dbg >> 8) & 15)) g_ft_probe[i] = clock64(); } while (0)

struct FtArgs {
    const EctLatPlan* latplans; const EctFftPlan* plans;
    const uint16_t* perm_pool; const double2* tw_pool; const void* cz_pool; const double2* roots;   // cz_pool: double2 (dp) or float2 (sp)
    const int* lat_plan; const i64* latrow0; const int* fft_rec;
    const int* lats;              // latitudes (local index) of this launch
    const int* lat_aff;           // per local latitude: bit 0: fft_rec[k] = fft_rec[0] + k, bit 1: one destination rank and dst_rec[k] = dst_rec[0] + k
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
    int fp32;                     // sp handle: grid-point arrays are float and the FFT arithmetic is float
    int adj;                      // adjoint call: direct pipeline without 1/N and Gaussian weight (INV_TRANSAD), inverse with them (DIR_TRANSAD)
    int nostage;                  // this launch's rows do not fit with a staging area: inputs are read straight from HBM
    int dbg;                      // debug switches (ECT_FFT_DBG): 1 no output chirp, 2 no stores, 4 no middle kernel spectrum, 8 direct: no chirp loads, 16 direct: no record stores (timing experiments, wrong results)
    // direct, peer mode: records of one CTA's field chunk are collected in a local slot ([m][FT_PUSH_W] double2) and leave
    // the GPU as runs of up to 256 bytes per (lat, m) record.  Slots: push_sps per SM, claimed through a bit mask per SM
    double2* push_scr; unsigned* push_mask; int push_sps; i64 push_slot;     // push_slot: double2 per slot
};
#define FT_PUSH_W (2 * FT_PAIRS_PER_CTA)

// Claim / release one of the push_sps scratch slots of the SM this CTA runs on (thread 0 of the CTA).  At most push_sps - 1
// CTAs of a launch are resident per SM (host: launch_fourier), so a free bit always exists; the loop only rides out the
// window between a neighbour's release and our read.
__device__ __forceinline__ int ft_slot_acquire(unsigned* masks, int sps) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    unsigned* m = masks + smid;
    const unsigned all = sps >= 32 ? 0xffffffffu : ((1u << sps) - 1u);
    for (;;) {
        const unsigned cur = *reinterpret_cast<volatile unsigned*>(m);
        const unsigned fr = ~cur & all;
        if (!fr) { __nanosleep(200); continue; }
        const int b = __ffs(fr) - 1;
        if (!(atomicOr(m, 1u << b) & (1u << b))) return (int)smid * 32 + b;
    }
}
// %smid is not guaranteed to be contiguous: the tables are sized by %nsmid, read once on the device
__global__ void k_ft_nsmid(unsigned* out) {
    unsigned n;
    asm volatile("mov.u32 %0, %%nsmid;" : "=r"(n));
    *out = n;
}
__device__ __forceinline__ void ft_slot_release(unsigned* masks, int id) {
    __threadfence();
    atomicAnd(masks + (id >> 5), ~(1u << (id & 31)));
}

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

// Byte offsets of the shared-memory areas of one CTA (host: class sizing, device: carve-up)
struct FtLayout { int stage, chirp, t1, t2, roots, rec, total; };
__host__ __device__ inline int ft_al16(int b) { return (b + 15) & ~15; }
__host__ __device__ inline FtLayout ft_layout(bool inverse, bool bluestein, int len, int N, int km, int nroots,
                                              int csize, int iosize, bool nostage) {
    FtLayout L;
    L.stage = ft_al16(ECT_PADDED_LEN(len) * csize);
    const int nst = nostage ? 0 : (inverse ? 2 * (km + 1) * (int)sizeof(double2) : 2 * N * iosize);
    L.chirp = L.stage + ft_al16(nst);
    L.t1 = L.chirp + ((inverse && bluestein) ? ft_al16((N / 2 + 1) * csize) : 0);
    L.t2 = L.t1 + ft_al16(ECT_TW1_LEN(len) * csize);
    L.roots = L.t2 + ECT_TW2_LEN * csize;
    L.rec = L.roots + ft_al16(nroots * csize);
    L.total = L.rec + ft_al16((km + 1) * (int)sizeof(int));
    return L;
}

// Shared memory of one CTA:
//   data [ECT_PADDED_LEN(len)] double2   work array of the pair in flight
//   stage                                raw inputs of the NEXT pair, filled by cp.async while this pair is transformed
//                                        inverse: (km+1) x {field a, field b} double2 ; direct: 2 x nlon doubles
//   t1, t2                               two-level twiddle table
//   roots                                odd-radix root tables
template <bool INVERSE, int MAXR, int TB, bool FP32>
__global__ void __launch_bounds__(TB) k_fourier(FtArgs a) {
    typedef typename std::conditional<FP32, float2, double2>::type C;     // arithmetic / work-array type
    typedef typename EctReal<C>::type R_;
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ EctFftPlan s_plan;        // stage list indexed at run time: keep it out of local memory
    __shared__ int s_slot;
    constexpr int NROOTS = MAXR <= 7 ? ECT_ROOTS_OFF(8) : ECT_ROOTS_SIZE;
    const int item = blockIdx.x;
    const int l = a.lats[item / a.nchunks];
    const int chunk = item % a.nchunks;
    const EctLatPlan lp = a.latplans[a.lat_plan[l]];
    const bool push = !INVERSE && a.push_scr != nullptr;
    if (threadIdx.x == 0) {
        s_plan = a.plans[lp.plan];
        if (push) s_slot = ft_slot_acquire(a.push_mask, a.push_sps);
    }
    __syncthreads();
    const int plan_n = s_plan.n, plan_nst = s_plan.nst;
    const uint16_t* perm = a.perm_pool + s_plan.perm_off;
    const bool blue = lp.bluestein != 0;
    const C* czp = reinterpret_cast<const C*>(a.cz_pool);
    const C* g_chirp = blue ? czp + lp.chirp_off : nullptr;
    const C* bhat = blue ? czp + (INVERSE ? lp.bhat_inv_off : lp.bhat_dir_off) : nullptr;
    const int* recs = a.fft_rec + a.latrow0[l];
    const int cp = a.cp;
    const int len = plan_n;
    const int N = lp.nlon, km = lp.km;
    const bool staged = !a.nostage;
    const bool oddrefl = (N & 1) != 0;     // odd row length (chirp-z always): c[N - j] = -c[j]
    const FtLayout lay = ft_layout(INVERSE, blue, len, N, km, NROOTS, (int)sizeof(C), FP32 ? 4 : 8, !staged);
    C* data = reinterpret_cast<C*>(smraw);
    double2* stage = reinterpret_cast<double2*>(smraw + lay.stage);
    // inverse chirp-z rows keep the chirp c[0 .. N/2] in shared memory
    const bool chirp_sm = INVERSE && blue;
    C* s_chirp = reinterpret_cast<C*>(smraw + lay.chirp);
    C* t1 = reinterpret_cast<C*>(smraw + lay.t1);
    C* t2 = reinterpret_cast<C*>(smraw + lay.t2);
    C* s_roots = reinterpret_cast<C*>(smraw + lay.roots);
    int* s_rec = reinterpret_cast<int*>(smraw + lay.rec);   // per m: inverse local record, direct (dest rank << 24 | dest record)
    const int tid = threadIdx.x, nthr = blockDim.x;
    tw_build(t1, t2, a.tw_pool + s_plan.tw_off, len, tid, nthr);
    for (int j = tid; j < NROOTS; j += nthr) s_roots[j] = c_cvt<C>(a.roots[j]);
    if (chirp_sm) for (int j = tid; j <= N / 2; j += nthr) s_chirp[j] = g_chirp[j];
    for (int k = tid; k <= km; k += nthr)
        s_rec[k] = INVERSE ? recs[k] : ((a.dst_rank[a.latrow0[l] + k] << 24) | a.dst_rec[a.latrow0[l] + k]);
    const C* chirp_tab = chirp_sm ? s_chirp : g_chirp;
    const EctTwT<C> qt{t1, t2};
    const int g0 = a.gpoff[l];
    const bool oneblk = a.nproma >= a.ngptot;
    const int p0 = chunk * FT_PAIRS_PER_CTA, p1 = min(p0 + FT_PAIRS_PER_CTA, a.npairs);
    constexpr int NB = 4;      // global loads issued per thread before the first use (latency batching)
    const double racthe = a.racthe_loc[l];
    const R_ s1 = (R_)racthe, s2 = (R_)(racthe * racthe);
    // push mode: slot of this CTA, first field of its chunk (fields of a chunk are consecutive, api.cu make_pairs)
    double2* scr = nullptr;
    int pf0 = 0;
    if (push) {
        scr = a.push_scr + (i64)((s_slot >> 5) * a.push_sps + (s_slot & 31)) * a.push_slot;
        pf0 = p0 < p1 ? a.pairs[p0].x : 0;
    }

    auto prefetch = [&](int p) {       // raw inputs of pair p -> stage (asynchronous)
        if (!staged) return;
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
                    const double* src = a.fb + (long long)s_rec[k] * cp;
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
            if (!blue) { for (int k = km + 1 + tid; k < N - km; k += nthr) data[ECT_PAD((int)perm[k])] = c_make<C>(0, 0); }
            else { for (int u = 2 * km + 1 + tid; u < lp.m; u += nthr) data[ECT_PAD(u)] = c_make<C>(0, 0); }
            // DIR_TRANSAD: (DIR_TRANS)^T M = diag(w / N) . INV_TRANS, the row factor rides on the load scaling
            const R_ rs_ = a.adj ? (R_)(a.rw_loc[l] / (double)N) : (R_)1;
            const R_ sa_ = (sfa.pw == 0 ? (R_)1 : (sfa.pw == 1 ? s1 : s2)) * rs_;
            const R_ sb_ = (sfb.pw == 0 ? (R_)1 : (sfb.pw == 1 ? s1 : s2)) * rs_;
            for (int k0 = tid; k0 <= km; k0 += NB * nthr) {
                C ch[NB]; int pk[NB], pn[NB]; double2 ra[NB], rb[NB];
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const int k = k0 + i * nthr;
                    ch[i] = c_make<C>(1, 0); pk[i] = pn[i] = 0;
                    ra[i] = rb[i] = make_double2(0.0, 0.0);
                    if (k <= km) {
                        if (blue) ch[i] = chirp_tab[k];
                        else { pk[i] = perm[k]; pn[i] = perm[k == 0 ? 0 : N - k]; }
                        if (staged) { ra[i] = stage[2 * k]; if (hasb) rb[i] = stage[2 * k + 1]; }
                        else {
                            const double* src = a.fb + (long long)s_rec[k] * cp;
                            ra[i] = *reinterpret_cast<const double2*>(src + sfa.src_c);
                            if (hasb) rb[i] = *reinterpret_cast<const double2*>(src + sfb.src_c);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const int k = k0 + i * nthr;
                    if (k > km) continue;
                    const C va = c_cvt<C>(ra[i]), vb = c_cvt<C>(rb[i]);
                    C fa_ = c_make<C>(va.x * sa_, k == 0 ? (R_)0 : va.y * sa_);
                    C fb_ = c_make<C>(vb.x * sb_, k == 0 ? (R_)0 : vb.y * sb_);
                    const R_ z = s1 * (R_)k;
                    if (sfa.deriv) fa_ = c_make<C>(-fa_.y * z, fa_.x * z);
                    if (sfb.deriv) fb_ = c_make<C>(-fb_.y * z, fb_.x * z);
                    const C zp = c_make<C>(fa_.x - fb_.y, fa_.y + fb_.x);      // Z[k]
                    const C zm = c_make<C>(fa_.x + fb_.y, fb_.x - fa_.y);      // Z[-k]
                    if (!blue) {
                        data[ECT_PAD(pk[i])] = zp;
                        if (k > 0) data[ECT_PAD(pn[i])] = zm;
                    } else {
                        const C xp = c_mul(zp, ch[i]);
                        data[ECT_PAD(km + k)] = c_make<C>(xp.y, xp.x);
                        if (k > 0) { const C xm = c_mul(zm, ch[i]); data[ECT_PAD(km - k)] = c_make<C>(xm.y, xm.x); }
                    }
                }
            }
        } else {
            const double* st = reinterpret_cast<const double*>(stage);
            const float* stf = reinterpret_cast<const float*>(stage);
            const double* ba = a.gp_base[fa]; const i64 sa = a.gp_blk[fa];
            const double* bb = hasb ? a.gp_base[fb2] : nullptr; const i64 sb = hasb ? a.gp_blk[fb2] : 0;
            // chirp factors come from global memory (no shared memory left next to two staged rows): the loads of
            // batch i+1 are in flight while batch i is scattered
            auto ld_batch = [&](int j0, C* ch, int* pj, R_* xa, R_* xb) {
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const int j = j0 + i * nthr;
                    ch[i] = c_make<C>(1, 0); pj[i] = 0; xa[i] = xb[i] = 0;
                    if (j < N) {
                        if (blue && !(a.dbg & 8)) { ch[i] = g_chirp[j > N / 2 ? N - j : j]; if (oddrefl && j > N / 2) ch[i] = c_make<C>(-ch[i].x, -ch[i].y); }
                        else if (!blue) pj[i] = perm[j];
                        if (!staged) {
                            const int g = g0 + j;
                            const i64 ia = oneblk ? (i64)g : gp_index(g, a.nproma, sa);
                            xa[i] = FP32 ? (R_)reinterpret_cast<const float*>(ba)[ia] : (R_)ba[ia];
                            if (hasb) {
                                const i64 ib = oneblk ? (i64)g : gp_index(g, a.nproma, sb);
                                xb[i] = FP32 ? (R_)reinterpret_cast<const float*>(bb)[ib] : (R_)bb[ib];
                            }
                        }
                    }
                }
            };
            auto put_batch = [&](int j0, const C* ch, const int* pj, const R_* xa, const R_* xb) {
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const int j = j0 + i * nthr;
                    if (j >= N) continue;
                    R_ va = xa[i], vb = xb[i];
                    if (staged) {
                        va = FP32 ? (R_)stf[j] : (R_)st[j];
                        vb = hasb ? (FP32 ? (R_)stf[N + j] : (R_)st[N + j]) : (R_)0;
                    }
                    if (!blue) data[ECT_PAD(pj[i])] = c_make<C>(vb, va);
                    else { const C t = c_mul(c_make<C>(vb, va), ch[i]); data[ECT_PAD(j)] = c_make<C>(t.y, t.x); }
                }
            };
            C chA[NB], chB[NB]; int pjA[NB], pjB[NB]; R_ xaA[NB], xbA[NB], xaB[NB], xbB[NB];
            ld_batch(tid, chA, pjA, xaA, xbA);
            for (int j0 = tid; j0 < N; j0 += 2 * NB * nthr) {
                ld_batch(j0 + NB * nthr, chB, pjB, xaB, xbB);
                put_batch(j0, chA, pjA, xaA, xbA);
                ld_batch(j0 + 2 * NB * nthr, chA, pjA, xaA, xbA);
                put_batch(j0 + NB * nthr, chB, pjB, xaB, xbB);
            }
            // zero tail of the chirp-z work array (fourier_phases.h ftdir_zero_tail)
            if (blue) for (int u = N + tid; u < lp.m; u += nthr) data[ECT_PAD(u)] = c_make<C>(0, 0);
        }
        __syncthreads();                 // work array complete, stage consumed
        if (p + 1 < p1) prefetch(p + 1);
        FT_PROBE(1);
        if (blue) {
            for (int s = plan_nst - 1; s >= 1; --s) {
                fft_stage<true, 7, false>(data, len, s_plan.radix[s], s_plan.sublen[s], s_plan.lshift[s], qt, (const C*)s_roots, tid, nthr);
                __syncthreads();
                FT_PROBE(2 + (plan_nst - 1 - s));
            }
            blue_middle(data, len, s_plan.radix[0], (a.dbg & 4) ? czp : bhat, tid, nthr);
            __syncthreads();
            FT_PROBE(10);
            for (int s = 1; s < plan_nst; ++s) {
                fft_stage<false, 7, false>(data, len, s_plan.radix[s], s_plan.sublen[s], s_plan.lshift[s], qt, (const C*)s_roots, tid, nthr);
                __syncthreads();
                FT_PROBE(10 + s);
            }
        } else {
            for (int s = 0; s < plan_nst; ++s) {
                fft_stage<false, MAXR, false>(data, len, s_plan.radix[s], s_plan.sublen[s], s_plan.lshift[s], qt, (const C*)s_roots, tid, nthr);
                __syncthreads();
            }
            FT_PROBE(19);
        }
        if (INVERSE) {
            double* ba = a.gp_base[fa]; const i64 sa = a.gp_blk[fa];
            double* bb = hasb ? a.gp_base[fb2] : nullptr; const i64 sb = hasb ? a.gp_blk[fb2] : 0;
            for (int j0 = tid; j0 < N; j0 += NB * nthr) {
                C x[NB], ch[NB];
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const int j = j0 + i * nthr;
                    if (j < N) {
                        x[i] = data[ECT_PAD(j)];
                        if (blue) { ch[i] = chirp_tab[j > N / 2 ? N - j : j]; if (oddrefl && j > N / 2) ch[i] = c_make<C>(-ch[i].x, -ch[i].y); }
                    }
                }
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const int j = j0 + i * nthr;
                    if (j >= N) continue;
                    const C y = blue ? c_mul(ch[i], x[i]) : x[i];
                    if ((a.dbg & 2) && y.x != (R_)12345.678) continue;
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
            // INV_TRANSAD: M^-1 (INV_TRANS)^T is the direct pipeline without 1/N and without the Gaussian weight
            const double wl = a.adj ? 1.0 : a.rw_loc[l] / (double)N;
            const R_ sca = (R_)(0.5 * wl * (fa < a.n_uv_fields ? racthe : 1.0));
            const R_ scb = (R_)(0.5 * wl * (fb2 >= 0 && fb2 < a.n_uv_fields ? racthe : 1.0));
            const int ca = 2 * fa, cb = hasb ? 2 * fb2 : -1;
            for (int k0 = tid; k0 <= km; k0 += NB * nthr) {
                C zk[NB], zn[NB], ch[NB]; double* rb[NB];
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const int k = k0 + i * nthr;
                    if (k <= km) {
                        const int pk_ = s_rec[k];
                        rb[i] = a.peer[pk_ >> 24] + (long long)(pk_ & 0xffffff) * cp;
                        if (!blue) { zk[i] = data[ECT_PAD(k)]; zn[i] = data[ECT_PAD(k == 0 ? 0 : N - k)]; }
                        else { zk[i] = data[ECT_PAD(km + k)]; zn[i] = data[ECT_PAD(km - k)]; ch[i] = (a.dbg & 8) ? c_make<C>(1, 0) : g_chirp[k]; }
                    }
                }
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const int k = k0 + i * nthr;
                    if (k > km) continue;
                    C a_ = zk[i], b_ = zn[i];
                    if (blue) { a_ = c_mul(ch[i], a_); b_ = c_mul(ch[i], b_); }
                    // stored values are swapped (sign - transform on the sign + core): Z = (y, x)
                    const C Zk = c_make<C>(a_.y, a_.x), Zn = c_make<C>(b_.y, b_.x);
                    const double2 ra_ = make_double2((double)((Zk.x + Zn.x) * sca), (double)((Zk.y - Zn.y) * sca));
                    const double2 rb_ = make_double2((double)((Zk.y + Zn.y) * scb), (double)((Zn.x - Zk.x) * scb));
                    if ((a.dbg & 16) && ra_.x != 12345.678) continue;      // timing experiment: no record stores
                    if (push) {
                        double2* q = scr + (i64)k * FT_PUSH_W + (fa - pf0);
                        q[0] = ra_;
                        if (cb >= 0) q[fb2 - fa] = rb_;
                    } else {
                        *reinterpret_cast<double2*>(rb[i] + ca) = ra_;
                        if (cb >= 0) *reinterpret_cast<double2*>(rb[i] + cb) = rb_;
                    }
                }
            }
        }
        __syncthreads();
        FT_PROBE(20);
    }
    if (push) {
        // the chunk's records, [m][field] in the slot (written by this CTA: L2 hits), go to their consumer as one run of
        // nfc * 16 bytes per (lat, m): 16 lanes per record
        if (p0 < p1) {
            const int2 pl = a.pairs[p1 - 1];
            const int nfc = (pl.y >= 0 ? pl.y : pl.x) - pf0 + 1;
            const int j = tid & (FT_PUSH_W - 1), kstep = nthr / FT_PUSH_W;
            constexpr int NP = 8;              // loads in flight per thread: the slot is read once, at L2 latency
            if (j < nfc)
                for (int k0 = tid / FT_PUSH_W; k0 <= km; k0 += NP * kstep) {
                    double2 v[NP];
#pragma unroll
                    for (int i = 0; i < NP; ++i) {
                        const int k = k0 + i * kstep;
                        if (k <= km) v[i] = __ldcg(scr + (i64)k * FT_PUSH_W + j);
                    }
#pragma unroll
                    for (int i = 0; i < NP; ++i) {
                        const int k = k0 + i * kstep;
                        if (k > km) continue;
                        const int pk_ = s_rec[k];
                        double* dst = a.peer[pk_ >> 24] + (long long)(pk_ & 0xffffff) * cp + 2 * (pf0 + j);
                        *reinterpret_cast<double2*>(dst) = v[i];
                    }
                }
        }
        __syncthreads();
        if (tid == 0) ft_slot_release(a.push_mask, s_slot);
    }
}


// ---------------------------------------------------------------------------------------
// Chirp-z rows, split over a CTA pair (fourier_cz.h).  Cluster of two CTAs = one latitude x a chunk of field pairs;
// CTA 0 convolves the even bins, CTA 1 the odd bins, each in a work array of H = M/2 elements; the halves are
// combined in the output phase through distributed shared memory.  No staging area and no chirp copy in shared
// memory: several CTAs per SM (three for the longest dp rows of TCo1279) hide the global-memory latency instead.
// ---------------------------------------------------------------------------------------
__host__ __device__ inline int cz_smem_bytes(int H, int N, int csize) {
    return (ECT_PADDED_LEN(H) + ECT_TW1_LEN(2 * H) + ECT_TW2_LEN + ECT_ROOTS_OFF(8) + ECT_TW1_LEN(2 * N) + ECT_TW2_LEN) * csize;
}
__device__ __forceinline__ void ft_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" :: "l"(p)); }
#define CZ_MINB(TB, FP32) ((TB) == 128 ? ((FP32) ? 4 : 3) : ((TB) == 192 || (TB) == 256 ? 2 : 1))

template <bool INVERSE, int TB, bool FP32>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TB, CZ_MINB(TB, FP32)) k_fourier_cz(FtArgs a) {
    typedef typename std::conditional<FP32, float2, double2>::type C;
    typedef typename EctReal<C>::type R_;
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ EctFftPlan s_plan;
    cg::cluster_group cluster = cg::this_cluster();
    const int half = (int)cluster.block_rank();
    const int item = blockIdx.x >> 1;
    const int l = a.lats[item / a.nchunks];
    const int chunk = item % a.nchunks;
    const EctLatPlan lp = a.latplans[a.lat_plan[l]];
    if (threadIdx.x == 0) s_plan = a.plans[lp.plan_h];
    __syncthreads();
    const int H = s_plan.n, plan_nst = s_plan.nst;
    const int N = lp.nlon, km = lp.km;
    const C* czp = reinterpret_cast<const C*>(a.cz_pool);
    const C* bhat = czp + (INVERSE ? lp.bhat_inv_eo[half] : lp.bhat_dir_eo[half]);
    C* data = reinterpret_cast<C*>(smraw);
    C* t1 = data + ECT_PADDED_LEN(H);
    C* t2 = t1 + ECT_TW1_LEN(2 * H);
    C* s_roots = t2 + ECT_TW2_LEN;
    C* t1c = s_roots + ECT_ROOTS_OFF(8);
    C* t2c = t1c + ECT_TW1_LEN(2 * N);
    const int tid = threadIdx.x, nthr = TB;
    // one two-level table of the M-th roots of unity serves the radix-2 step (w^u) and, with every second entry, the
    // length-H passes
    tw_build(t1, t2, a.tw_pool + a.plans[lp.plan].tw_off, 2 * H, tid, nthr);
    for (int j = tid; j < ECT_ROOTS_OFF(8); j += nthr) s_roots[j] = c_cvt<C>(a.roots[j]);
    // chirp c[j] = exp(i pi j^2 / N) as entry (j^2 mod 2N) of a two-level table of the 2N-th roots of unity
    for (int j = tid; j < ECT_TW1_LEN(2 * N) + ECT_TW2_LEN; j += nthr) t1c[j] = czp[lp.ctw_off + j];
    EctTwT<C> qth; qth.t1 = t1; qth.t2 = t2; qth.sh = 1;
    CzCtx<C> cx;
    cx.N = N; cx.km = km; cx.H = H; cx.half = half;
    cx.twm.t1 = t1; cx.twm.t2 = t2; cx.twm.sh = 0;
    cx.twc.t1 = t1c; cx.twc.t2 = t2c; cx.twc.sh = 0;
    cx.n2 = 2u * (unsigned)N; cx.magic = (unsigned)((0x100000000ull + cx.n2 - 1) / cx.n2);
    const C* other = cluster.map_shared_rank(data, half ^ 1);
    const int* recs = a.fft_rec + a.latrow0[l];
    // records of consecutive wavenumbers are consecutive (always on one rank): no index loads in front of the data loads
    const int aff = a.lat_aff[l];
    const int rec0 = recs[0];
    const int cp = a.cp;
    const int g0 = a.gpoff[l];
    const bool oneblk = a.nproma >= a.ngptot;
    const int p0 = chunk * FT_PAIRS_PER_CTA, p1 = min(p0 + FT_PAIRS_PER_CTA, a.npairs);
    const double racthe = a.racthe_loc[l];
    const bool warp_local = plan_nst >= 2 && s_plan.radix[0] == 16 && s_plan.radix[1] == 16 && (H & 511) == 0;
    __syncthreads();
    for (int p = p0; p < p1; ++p) {
        const int2 pr = a.pairs[p];
        const int fa = pr.x, fb2 = pr.y;
        const bool hasb = fb2 >= 0;
        if (INVERSE) {
            const EctFsField sfa = a.fsf[fa];
            EctFsField sfb; sfb.src_c = -1; sfb.pw = 0; sfb.deriv = 0;
            if (hasb) sfb = a.fsf[fb2];
            CzInvScale sc;
            sc.pwa = sfa.pw; sc.deriva = sfa.deriv; sc.pwb = sfb.pw; sc.derivb = sfb.deriv; sc.hasb = hasb;
            sc.s1 = racthe; sc.s2 = racthe * racthe;
            sc.rowscale = a.adj ? a.rw_loc[l] / (double)N : 1.0;       // DIR_TRANSAD: diag(w / N) . INV_TRANS
            const int ca = sfa.src_c, cb = sfb.src_c;
            cz_inv_load(data, cx, sc, [&](int k, double2& ra, double2& rb) {
                const double* src = a.fb + (long long)((aff & 1) ? rec0 + k : recs[k]) * cp;
                ra = *reinterpret_cast<const double2*>(src + ca);
                if (hasb) rb = *reinterpret_cast<const double2*>(src + cb);
            }, tid, nthr);
            if (p + 1 < p1) {        // the next pair's spectral values on their way into L2 while this pair is transformed
                const int cn = a.fsf[a.pairs[p + 1].x].src_c;
                for (int k = 2 * tid + half; k <= km; k += 2 * nthr)
                    ft_prefetch_l2(a.fb + (long long)((aff & 1) ? rec0 + k : recs[k]) * cp + cn);
            }
        } else {
            const double* ba = a.gp_base[fa]; const i64 sa = a.gp_blk[fa];
            const double* bb = hasb ? a.gp_base[fb2] : nullptr; const i64 sb = hasb ? a.gp_blk[fb2] : 0;
            cz_dir_load(data, cx, [&](int j, R_& va, R_& vb) {
                const int g = g0 + j;
                const i64 ia = oneblk ? (i64)g : gp_index(g, a.nproma, sa);
                va = FP32 ? (R_)reinterpret_cast<const float*>(ba)[ia] : (R_)ba[ia];
                if (hasb) {
                    const i64 ib = oneblk ? (i64)g : gp_index(g, a.nproma, sb);
                    vb = FP32 ? (R_)reinterpret_cast<const float*>(bb)[ib] : (R_)bb[ib];
                }
            }, tid, nthr);
            if (p + 1 < p1) {        // the next pair's rows on their way into L2 (one request per 128 bytes, shared by the two CTAs)
                const int2 pn = a.pairs[p + 1];
                const int es = FP32 ? 4 : 8, per = 128 / es;
                for (int t = 2 * tid + half; t * per < N; t += 2 * nthr) {
                    const int g = g0 + t * per;
                    ft_prefetch_l2(reinterpret_cast<const char*>(a.gp_base[pn.x]) + (oneblk ? (i64)g : gp_index(g, a.nproma, a.gp_blk[pn.x])) * es);
                    if (pn.y >= 0)
                        ft_prefetch_l2(reinterpret_cast<const char*>(a.gp_base[pn.y]) + (oneblk ? (i64)g : gp_index(g, a.nproma, a.gp_blk[pn.y])) * es);
                }
            }
        }
        __syncthreads();
        // Stages 0 and 1 of a 16 x 16 x ... plan work inside aligned blocks of 256 elements, and a warp's 32 consecutive
        // butterflies cover the same two blocks in either stage: DIF stage 1 -> fused middle -> DIT stage 1 need no
        // CTA-wide barrier, only the warp's own stores to be visible.  The warps of a CTA then drift apart and their
        // shared-memory and FP64 phases overlap instead of running in lock step.
        for (int s = plan_nst - 1; s >= 1; --s) {
            fft_stage<true, 7, true>(data, H, s_plan.radix[s], s_plan.sublen[s], s_plan.lshift[s], qth, (const C*)s_roots, tid, nthr);
            if (s == 1 && warp_local) __syncwarp(); else __syncthreads();
        }
        blue_middle_early(data, H, s_plan.radix[0], bhat, tid, nthr);
        if (warp_local) __syncwarp(); else __syncthreads();
        for (int s = 1; s < plan_nst; ++s) {
            fft_stage<false, 7, true>(data, H, s_plan.radix[s], s_plan.sublen[s], s_plan.lshift[s], qth, (const C*)s_roots, tid, nthr);
            __syncthreads();
        }
        cluster.sync();                   // both halves complete and visible to the partner
        if (INVERSE) {
            double* ba = a.gp_base[fa]; const i64 sa = a.gp_blk[fa];
            double* bb = hasb ? a.gp_base[fb2] : nullptr; const i64 sb = hasb ? a.gp_blk[fb2] : 0;
            cz_inv_out((const C*)data, other, cx, [&](int j, C y) {
                const int g = g0 + j;
                const i64 ia = oneblk ? (i64)g : gp_index(g, a.nproma, sa);
                if (FP32) reinterpret_cast<float*>(ba)[ia] = (float)y.x; else ba[ia] = (double)y.x;
                if (hasb) {
                    const i64 ib = oneblk ? (i64)g : gp_index(g, a.nproma, sb);
                    if (FP32) reinterpret_cast<float*>(bb)[ib] = (float)y.y; else bb[ib] = (double)y.y;
                }
            }, tid, nthr);
        } else {
            // 1/N (tpm_fftw.F90:317-321), Gaussian weight (ledir_mod.F90:122) and, for u and v, 1/(a cos theta)
            // (ldfou2_mod.F90:90-96) in one factor; INV_TRANSAD: neither 1/N nor the weight
            const double wl = a.adj ? 1.0 : a.rw_loc[l] / (double)N;
            const R_ sca = (R_)(0.5 * wl * (fa < a.n_uv_fields ? racthe : 1.0));
            const R_ scb = (R_)(0.5 * wl * (fb2 >= 0 && fb2 < a.n_uv_fields ? racthe : 1.0));
            const int ca = 2 * fa, cb = hasb ? 2 * fb2 : -1;
            const int* drank = a.dst_rank + a.latrow0[l];
            const int* drec = a.dst_rec + a.latrow0[l];
            double* const peer0 = a.peer[drank[0]];
            const int drec0 = drec[0];
            cz_dir_out((const C*)data, other, cx, [&](int k, C Zk, C Zn) {
                double* rb = (aff & 2) ? peer0 + (long long)(drec0 + k) * cp : a.peer[drank[k]] + (long long)drec[k] * cp;
                *reinterpret_cast<double2*>(rb + ca) = make_double2((double)((Zk.x + Zn.x) * sca), (double)((Zk.y - Zn.y) * sca));
                if (cb >= 0) *reinterpret_cast<double2*>(rb + cb) = make_double2((double)((Zk.y + Zn.y) * scb), (double)((Zn.x - Zk.x) * scb));
            }, tid, nthr);
        }
        cluster.sync();                   // the partner has read this CTA's half: the work array may be overwritten
    }
}


// ---------------------------------------------------------------------------------------
// Chirp-z rows, both halves of the split convolution in ONE CTA: two warp groups (E: even bins, O: odd bins), each with
// its own work array and its own named barrier, meeting only where the staged inputs are consumed and where the
// outputs are formed.  What this keeps from the undivided kernel: the next pair's inputs are staged by cp.async while the
// current pair is transformed (no exposed global-memory latency, the inputs are read once), one CTA per SM whose warps
// walk through the same code (instruction-cache working set = one pass).  What it takes from the pair kernel: two
// independent half-length transforms -- the groups drift apart, so one group's shared-memory phase overlaps the other's
// FP64 phase instead of all warps running each pass in lock step -- and two passes less per transform.
// ---------------------------------------------------------------------------------------
struct Cz2Layout { int o, t1, roots, tc, stage, rec, total; };
__host__ __device__ inline Cz2Layout cz2_layout(bool inverse, int H, int N, int km, int csize, int iosize, bool nostage) {
    Cz2Layout L;
    L.o = ft_al16(ECT_PADDED_LEN(H) * csize);
    L.t1 = 2 * L.o;
    L.roots = L.t1 + ft_al16((ECT_TW1_LEN(2 * H) + ECT_TW2_LEN) * csize);
    L.tc = L.roots + ft_al16(ECT_ROOTS_OFF(8) * csize);
    L.stage = L.tc + ft_al16((ECT_TW1_LEN(2 * N) + ECT_TW2_LEN) * csize);      // 16-byte aligned: cp.async destination
    const int nst = nostage ? 0 : (inverse ? 2 * (km + 1) * (int)sizeof(double2) : 2 * N * iosize);
    L.rec = L.stage + ft_al16(nst);
    L.total = L.rec + ft_al16((km + 1) * (int)sizeof(int));
    return L;
}
// named barriers 1 / 2 with constant ids (a register id makes ptxas reserve all 16 hardware barriers: one CTA per SM)
__device__ __forceinline__ void cz2_group_sync(int g, int n) {
    if (g == 0) asm volatile("bar.sync 1, %0;\n" :: "r"(n) : "memory");
    else asm volatile("bar.sync 2, %0;\n" :: "r"(n) : "memory");
}

template <bool INVERSE, int GS, bool FP32>
__global__ void __launch_bounds__(2 * GS, 384 / (2 * GS)) k_fourier_cz2(FtArgs a) {
    typedef typename std::conditional<FP32, float2, double2>::type C;
    typedef typename EctReal<C>::type R_;
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ EctFftPlan s_plan;
    const int item = blockIdx.x;
    const int l = a.lats[item / a.nchunks];
    const int chunk = item % a.nchunks;
    const EctLatPlan lp = a.latplans[a.lat_plan[l]];
    if (threadIdx.x == 0) s_plan = a.plans[lp.plan_h];
    __syncthreads();
    const int H = s_plan.n, plan_nst = s_plan.nst;
    const int N = lp.nlon, km = lp.km;
    const bool staged = !a.nostage;
    const Cz2Layout lay = cz2_layout(INVERSE, H, N, km, (int)sizeof(C), FP32 ? 4 : 8, !staged);
    const int tid = threadIdx.x, nthr = 2 * GS;
    const int g = tid / GS, gt = tid - g * GS;            // warp group (0: even bins, 1: odd bins) and thread in it
    const C* czp = reinterpret_cast<const C*>(a.cz_pool);
    const C* bhat = czp + (INVERSE ? lp.bhat_inv_eo[g] : lp.bhat_dir_eo[g]);
    C* mine = reinterpret_cast<C*>(smraw + (g ? lay.o : 0));
    const C* other = reinterpret_cast<const C*>(smraw + (g ? 0 : lay.o));
    C* t1 = reinterpret_cast<C*>(smraw + lay.t1);
    C* t2 = t1 + ECT_TW1_LEN(2 * H);
    C* s_roots = reinterpret_cast<C*>(smraw + lay.roots);
    C* t1c = reinterpret_cast<C*>(smraw + lay.tc);
    C* t2c = t1c + ECT_TW1_LEN(2 * N);
    double2* stage = reinterpret_cast<double2*>(smraw + lay.stage);
    int* s_rec = reinterpret_cast<int*>(smraw + lay.rec);
    tw_build(t1, t2, a.tw_pool + a.plans[lp.plan].tw_off, 2 * H, tid, nthr);
    for (int j = tid; j < ECT_ROOTS_OFF(8); j += nthr) s_roots[j] = c_cvt<C>(a.roots[j]);
    for (int j = tid; j < ECT_TW1_LEN(2 * N) + ECT_TW2_LEN; j += nthr) t1c[j] = czp[lp.ctw_off + j];
    const int* recs = a.fft_rec + a.latrow0[l];
    for (int k = tid; k <= km; k += nthr)
        s_rec[k] = INVERSE ? recs[k] : ((a.dst_rank[a.latrow0[l] + k] << 24) | a.dst_rec[a.latrow0[l] + k]);
    EctTwT<C> qth; qth.t1 = t1; qth.t2 = t2; qth.sh = 1;
    CzCtx<C> cx;
    cx.N = N; cx.km = km; cx.H = H; cx.half = g;
    cx.twm.t1 = t1; cx.twm.t2 = t2; cx.twm.sh = 0;
    cx.twc.t1 = t1c; cx.twc.t2 = t2c; cx.twc.sh = 0;
    cx.n2 = 2u * (unsigned)N; cx.magic = (unsigned)((0x100000000ull + cx.n2 - 1) / cx.n2);
    const int cp = a.cp;
    const int g0 = a.gpoff[l];
    const bool oneblk = a.nproma >= a.ngptot;
    const int p0 = chunk * FT_PAIRS_PER_CTA, p1 = min(p0 + FT_PAIRS_PER_CTA, a.npairs);
    const double racthe = a.racthe_loc[l];
    const bool warp_local = plan_nst >= 2 && s_plan.radix[0] == 16 && s_plan.radix[1] == 16 && (H & 511) == 0;

    auto prefetch = [&](int p) {       // raw inputs of pair p -> stage (asynchronous, all threads)
        if (!staged) return;
        const int2 pr = a.pairs[p];
        const int fa = pr.x, fb2 = pr.y;
        if (INVERSE) {
            const int ca = a.fsf[fa].src_c;
            const int cb = fb2 >= 0 ? a.fsf[fb2].src_c : -1;
            for (int k = tid; k <= km; k += nthr) {
                const double* src = a.fb + (long long)s_rec[k] * cp;
                ft_cp_async16(stage + 2 * k, src + ca);
                if (cb >= 0) ft_cp_async16(stage + 2 * k + 1, src + cb);
            }
        } else {
            const double* ba = a.gp_base[fa]; const i64 sa = a.gp_blk[fa];
            const double* bb = fb2 >= 0 ? a.gp_base[fb2] : nullptr; const i64 sb = fb2 >= 0 ? a.gp_blk[fb2] : 0;
            if (FP32) {
                float* st = reinterpret_cast<float*>(stage);
                const float* fa_ = reinterpret_cast<const float*>(ba); const float* fb_ = reinterpret_cast<const float*>(bb);
                for (int j = tid; j < N; j += nthr) {
                    const int gi = g0 + j;
                    ft_cp_async4(st + j, fa_ + (oneblk ? (i64)gi : gp_index(gi, a.nproma, sa)));
                    if (bb) ft_cp_async4(st + N + j, fb_ + (oneblk ? (i64)gi : gp_index(gi, a.nproma, sb)));
                }
            } else {
                double* st = reinterpret_cast<double*>(stage);
                for (int j = tid; j < N; j += nthr) {
                    const int gi = g0 + j;
                    ft_cp_async8(st + j, ba + (oneblk ? (i64)gi : gp_index(gi, a.nproma, sa)));
                    if (bb) ft_cp_async8(st + N + j, bb + (oneblk ? (i64)gi : gp_index(gi, a.nproma, sb)));
                }
            }
        }
        ft_cp_commit();
    };

    __syncthreads();                    // tables and s_rec visible
    if (p0 < p1) prefetch(p0);
    for (int p = p0; p < p1; ++p) {
        const int2 pr = a.pairs[p];
        const int fa = pr.x, fb2 = pr.y;
        const bool hasb = fb2 >= 0;
        ft_cp_wait_all();
        __syncthreads();                 // staged inputs visible; the previous pair's outputs have been read
        if (INVERSE) {
            const EctFsField sfa = a.fsf[fa];
            EctFsField sfb; sfb.src_c = -1; sfb.pw = 0; sfb.deriv = 0;
            if (hasb) sfb = a.fsf[fb2];
            CzInvScale sc;
            sc.pwa = sfa.pw; sc.deriva = sfa.deriv; sc.pwb = sfb.pw; sc.derivb = sfb.deriv; sc.hasb = hasb;
            sc.s1 = racthe; sc.s2 = racthe * racthe;
            sc.rowscale = a.adj ? a.rw_loc[l] / (double)N : 1.0;
            const int ca = sfa.src_c, cb = sfb.src_c;
            cz_inv_load(mine, cx, sc, [&](int k, double2& ra, double2& rb) {
                if (staged) { ra = stage[2 * k]; if (hasb) rb = stage[2 * k + 1]; }
                else {
                    const double* src = a.fb + (long long)s_rec[k] * cp;
                    ra = *reinterpret_cast<const double2*>(src + ca);
                    if (hasb) rb = *reinterpret_cast<const double2*>(src + cb);
                }
            }, gt, GS);
        } else {
            const double* ba = a.gp_base[fa]; const i64 sa = a.gp_blk[fa];
            const double* bb = hasb ? a.gp_base[fb2] : nullptr; const i64 sb = hasb ? a.gp_blk[fb2] : 0;
            const double* st = reinterpret_cast<const double*>(stage);
            const float* stf = reinterpret_cast<const float*>(stage);
            cz_dir_load(mine, cx, [&](int j, R_& va, R_& vb) {
                if (staged) {
                    va = FP32 ? (R_)stf[j] : (R_)st[j];
                    if (hasb) vb = FP32 ? (R_)stf[N + j] : (R_)st[N + j];
                } else {
                    const int gi = g0 + j;
                    const i64 ia = oneblk ? (i64)gi : gp_index(gi, a.nproma, sa);
                    va = FP32 ? (R_)reinterpret_cast<const float*>(ba)[ia] : (R_)ba[ia];
                    if (hasb) {
                        const i64 ib = oneblk ? (i64)gi : gp_index(gi, a.nproma, sb);
                        vb = FP32 ? (R_)reinterpret_cast<const float*>(bb)[ib] : (R_)bb[ib];
                    }
                }
            }, gt, GS);
        }
        __syncthreads();                 // both groups have consumed the stage
        if (p + 1 < p1) prefetch(p + 1);
        for (int s = plan_nst - 1; s >= 1; --s) {
            fft_stage<true, 7, true>(mine, H, s_plan.radix[s], s_plan.sublen[s], s_plan.lshift[s], qth, (const C*)s_roots, gt, GS);
            if (s == 1 && warp_local) __syncwarp(); else cz2_group_sync(g, GS);
        }
        blue_middle_early(mine, H, s_plan.radix[0], bhat, gt, GS);
        if (warp_local) __syncwarp(); else cz2_group_sync(g, GS);
        for (int s = 1; s < plan_nst; ++s) {
            fft_stage<false, 7, true>(mine, H, s_plan.radix[s], s_plan.sublen[s], s_plan.lshift[s], qth, (const C*)s_roots, gt, GS);
            if (s + 1 < plan_nst) cz2_group_sync(g, GS);
        }
        __syncthreads();                 // both halves complete
        if (INVERSE) {
            double* ba = a.gp_base[fa]; const i64 sa = a.gp_blk[fa];
            double* bb = hasb ? a.gp_base[fb2] : nullptr; const i64 sb = hasb ? a.gp_blk[fb2] : 0;
            cz_inv_out((const C*)mine, other, cx, [&](int j, C y) {
                const int gi = g0 + j;
                const i64 ia = oneblk ? (i64)gi : gp_index(gi, a.nproma, sa);
                if (FP32) reinterpret_cast<float*>(ba)[ia] = (float)y.x; else ba[ia] = (double)y.x;
                if (hasb) {
                    const i64 ib = oneblk ? (i64)gi : gp_index(gi, a.nproma, sb);
                    if (FP32) reinterpret_cast<float*>(bb)[ib] = (float)y.y; else bb[ib] = (double)y.y;
                }
            }, gt, GS);
        } else {
            const double wl = a.adj ? 1.0 : a.rw_loc[l] / (double)N;
            const R_ sca = (R_)(0.5 * wl * (fa < a.n_uv_fields ? racthe : 1.0));
            const R_ scb = (R_)(0.5 * wl * (fb2 >= 0 && fb2 < a.n_uv_fields ? racthe : 1.0));
            const int ca = 2 * fa, cb = hasb ? 2 * fb2 : -1;
            cz_dir_out((const C*)mine, other, cx, [&](int k, C Zk, C Zn) {
                const int pk_ = s_rec[k];
                double* rb = a.peer[pk_ >> 24] + (long long)(pk_ & 0xffffff) * cp;
                *reinterpret_cast<double2*>(rb + ca) = make_double2((double)((Zk.x + Zn.x) * sca), (double)((Zk.y - Zn.y) * sca));
                if (cb >= 0) *reinterpret_cast<double2*>(rb + cb) = make_double2((double)((Zk.y + Zn.y) * scb), (double)((Zn.x - Zk.x) * scb));
            }, gt, GS);
        }
    }
}

static void fill_args(EctHandle* h, const EctFieldCfg& f, FtArgs& a) {
    EctDevice* d = h->d;
    a.latplans = d->latplans; a.plans = d->plans;
    a.perm_pool = d->perm_pool; a.tw_pool = d->tw_pool; a.roots = d->roots;
    a.cz_pool = f.fp32 ? (const void*)d->cz_pool_f : (const void*)d->cz_pool;
    a.lat_plan = d->lat_plan; a.latrow0 = d->latrow0; a.fft_rec = d->fft_rec; a.lat_aff = d->lat_aff;
    a.gpoff = d->gpoff; a.nloen_loc = d->nloen; a.racthe_loc = d->racthe_loc;
    a.fb = d->fbuf_fft; a.cp = f.cp;
    a.peer = d->peer_leg; a.dst_rank = d->fft_dst_rank; a.dst_rec = d->fft_dst_rec;
    a.nfs = f.nfs; a.npairs = f.npairs;
    a.nchunks = (a.npairs + FT_PAIRS_PER_CTA - 1) / FT_PAIRS_PER_CTA;
    a.ngptot = h->hp.ngpband;      // the Fourier stage addresses the latitude band (== ngptot unless gp_eq)
    a.fp32 = f.fp32; a.adj = f.adj;
    a.rw_loc = d->rw_loc; a.n_uv_fields = 2 * f.kf_uv;
    static const char* dbg = getenv("ECT_FFT_DBG");
    a.dbg = dbg ? atoi(dbg) : 0;
    a.push_scr = nullptr; a.push_mask = nullptr; a.push_sps = 0; a.push_slot = 0;
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
        a.nostage = b.nostage;
        // buckets on concurrent streams would share the slots with different slot sizes: direct stores then
        const bool pushb = !INVERSE && !fork && b.push_sps > 0 && d->push_scr != nullptr;
        a.push_scr = pushb ? (double2*)d->push_scr : nullptr;
        a.push_mask = d->push_mask; a.push_sps = b.push_sps; a.push_slot = b.push_slot;
        const unsigned grid = (unsigned)(b.lats.size() * (size_t)a.nchunks);
        cudaStream_t st = d->stream;
        if (fork) { const int k = slot % (EctDevice::kSide + 1); st = k == 0 ? d->stream : d->side[k - 1]; }
        ++slot;
        const int thr = INVERSE ? b.threads_inv : b.threads;
        if (b.cz == 2) {                           // both halves in one CTA, two warp groups
#define CZ2_LAUNCH(GS_) do { if (a.fp32) k_fourier_cz2<INVERSE, GS_, true><<<grid, 2 * GS_, b.smem, st>>>(a); \
                             else k_fourier_cz2<INVERSE, GS_, false><<<grid, 2 * GS_, b.smem, st>>>(a); } while (0)
            if (b.threads == 384) CZ2_LAUNCH(192); else if (b.threads == 192) CZ2_LAUNCH(96); else CZ2_LAUNCH(64);
#undef CZ2_LAUNCH
            d->launches++;
            continue;
        }
        if (b.cz) {
            const unsigned g2 = 2 * grid;          // clusters of two CTAs (__cluster_dims__)
#define CZ_LAUNCH(TB_) do { if (a.fp32) k_fourier_cz<INVERSE, TB_, true><<<g2, TB_, b.smem, st>>>(a); \
                            else k_fourier_cz<INVERSE, TB_, false><<<g2, TB_, b.smem, st>>>(a); } while (0)
            if (b.threads == 128) CZ_LAUNCH(128); else if (b.threads == 192) CZ_LAUNCH(192); else if (b.threads == 256) CZ_LAUNCH(256); else CZ_LAUNCH(384);
#undef CZ_LAUNCH
            d->launches++;
            continue;
        }
        if (a.fp32) {
            if (b.maxr <= 7 && thr == 512) k_fourier<INVERSE, 7, 512, true><<<grid, thr, b.smem, st>>>(a);
            else if (b.maxr <= 7) k_fourier<INVERSE, 7, 256, true><<<grid, thr, b.smem, st>>>(a);
            else k_fourier<INVERSE, ECT_MAX_RADIX, 256, true><<<grid, thr, b.smem, st>>>(a);
        } else if (b.maxr <= 7 && thr == 384) k_fourier<INVERSE, 7, 384, false><<<grid, thr, b.smem, st>>>(a);
        else if (b.maxr <= 7 && thr == 512) k_fourier<INVERSE, 7, 512, false><<<grid, thr, b.smem, st>>>(a);
        else if (b.maxr <= 7) k_fourier<INVERSE, 7, 256, false><<<grid, thr, b.smem, st>>>(a);
        else k_fourier<INVERSE, ECT_MAX_RADIX, 256, false><<<grid, thr, b.smem, st>>>(a);
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


// Called once the transposition mode (peer memory or NCCL) is known: which latitudes have record tables that are
// affine in the wavenumber (FtArgs::lat_aff), so that the kernels need no index loads in front of their data accesses
int ect_fourier_set_affine(EctHandle* h) {
    EctHostPlan& P = h->hp;
    EctDevice* d = h->d;
    int rc;
    {
        const bool fused = d->p2p;
        const std::vector<int>& drk = P.fft_dst_rank; const std::vector<int>& drc = fused ? P.fft_dst_rec : P.fft_rec;
        std::vector<int> aff(P.nlat, 0);
        for (int l = 0; l < P.nlat; ++l) {
            const i64 r0 = P.latrow0[l], n = P.latrow0[l + 1] - r0;
            bool a0 = true, a1 = true;
            for (i64 k = 1; k < n; ++k) {
                if (P.fft_rec[r0 + k] != P.fft_rec[r0] + k) a0 = false;
                if (drc[r0 + k] != drc[r0] + k || (fused && drk[r0 + k] != drk[r0])) a1 = false;
            }
            aff[l] = (a0 ? 1 : 0) | (a1 ? 2 : 0);
        }
        if ((rc = upload(d->lat_aff, aff))) return rc;
    }
    // Peer mode: scratch slots of the direct stage's record push (k_fourier, FtArgs::push_scr).  ECT_FFT_PUSH=0 keeps
    // the direct 16-byte remote stores, 2 forces the slots whatever the rank count (tests), 1 turns them on for any peer-mode
    // run.  Default: three ranks or more -- NVLink takes about 8 G of the 16-byte packets per second from one GPU, which two
    // ranks stay under (3.2 GB per rank in 28 ms of FFT work: 110.5 ms per step against 112.0 with the push, whose extra pass
    // through L2 costs 2.5 ms there), four and eight do not (61.3 -> 57.2 ms, 33.1 -> 30.0 ms; profiles/r02_scaling.md)
    {
        const char* pe = getenv("ECT_FFT_PUSH");
        const int mode = pe ? atoi(pe) : -1;
        const bool on = mode == 2 || (d->p2p && P.nranks > 1 && (mode == 1 || (mode < 0 && P.nranks >= 3)));
        int smem_sm = 0, nsm = 0, thr_sm = 2048;
        cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, d->dev);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, d->dev);
        cudaDeviceGetAttribute(&thr_sm, cudaDevAttrMaxThreadsPerMultiProcessor, d->dev);
        if (on && d->push_nsm == 0) {
            unsigned* dn = nullptr; unsigned hn = 0;
            ECT_CUDA(cudaMalloc(&dn, sizeof(unsigned)));
            k_ft_nsmid<<<1, 1, 0, d->stream>>>(dn);
            ECT_CUDA(cudaMemcpyAsync(&hn, dn, sizeof(unsigned), cudaMemcpyDeviceToHost, d->stream));
            ECT_CUDA(cudaStreamSynchronize(d->stream));
            cudaFree(dn);
            d->push_nsm = std::max(nsm, (int)hn);
        }
        nsm = std::max(nsm, d->push_nsm);
        size_t need = 0;
        for (auto& b : d->buckets) {
            b.push_sps = 0; b.push_slot = 0;
            if (!on || b.cz || b.lats.empty()) continue;
            int kmmax = 0;
            for (int l : b.lats) kmmax = std::max(kmmax, d->fft.latplans[d->h_lat_plan[l]].km);
            // resident CTAs per SM are bounded by shared memory and threads; one spare slot
            const int res = std::max(1, std::min(smem_sm / (b.smem + 1024), thr_sm / std::max(b.threads, 32)));
            b.push_sps = std::min(32, res + 1);
            b.push_slot = (i64)(kmmax + 1) * FT_PUSH_W;
            need = std::max(need, (size_t)nsm * b.push_sps * (size_t)b.push_slot * sizeof(double2));
        }
        if (need > 0) {
            if (need > d->push_bytes) {
                if (d->push_scr) cudaFree(d->push_scr);
                d->push_scr = nullptr; d->push_bytes = 0;
                if (cudaMalloc(&d->push_scr, need) != cudaSuccess) {
                    // no room: the direct stage keeps storing straight into the consumer's buffer
                    cudaGetLastError();
                    for (auto& b : d->buckets) b.push_sps = 0;
                    need = 0;
                } else d->push_bytes = need;
            }
            if (need > 0 && !d->push_mask) {
                ECT_CUDA(cudaMalloc(&d->push_mask, (size_t)nsm * sizeof(unsigned)));
                ECT_CUDA(cudaMemset(d->push_mask, 0, (size_t)nsm * sizeof(unsigned)));
            }
        }
    }
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
    if (h->precision == ECT_PREC_SP) {
        std::vector<float2> czf(d->fft.cz_pool.size());
        for (size_t i = 0; i < czf.size(); ++i) czf[i] = make_float2((float)d->fft.cz_pool[i].x, (float)d->fft.cz_pool[i].y);
        if ((rc = upload(d->cz_pool_f, czf))) return rc;
    } else if ((rc = upload(d->cz_pool, d->fft.cz_pool))) return rc;
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
    const int limits[] = {16 * 1024, 32 * 1024, 56 * 1024, 74 * 1024, 112 * 1024, maxsm - 1024, maxsm - 1024};
    const int threads[] = {64, 64, 128, 128, 256, 256, 256};
    const bool sp = h->precision == ECT_PREC_SP;
    const int csize = sp ? (int)sizeof(float2) : (int)sizeof(double2), iosize = sp ? 4 : 8;
    d->buckets.clear();
    for (int v = 0; v < 2; ++v)
        for (int i = 0; i < 7; ++i) {
            EctDevice::Bucket b;
            b.smem = limits[i];
            b.threads = v ? std::min(threads[i], 256) : threads[i];
            static const char* t512 = getenv("ECT_FFT_T512");
            if (!v && i >= 4 && t512 && atoi(t512)) b.threads = 512;
            if (sp && !v && i >= 5) b.threads = 512;      // float work arrays: 128 registers/thread, 16 warps fit
            b.threads_inv = b.threads;
            // dp, big classes, radices <= 7 (the chirp-z rows): the inverse kernel fits 12 warps at 168 registers
            // (49.3 against 52.7 ms at TCo1279); the direct one does not gain (59.2 against 57.1) and keeps 8 warps
            static const char* t384 = getenv("ECT_FFT_T384");
            if (!sp && !v && i >= 5 && !(t384 && !atoi(t384))) b.threads_inv = 384;
            b.maxr = v ? ECT_MAX_RADIX : 7;
            b.nostage = i == 6;        // last class: rows too long for a staging area next to the work array
            d->buckets.push_back(b);
        }
    // chirp-z rows: split over CTA pairs (k_fourier_cz), one class per number of CTAs that fit an SM
    // ECT_FFT_CZ=1: every chirp-z row on the CTA-pair kernel; 2: on the two-warp-group kernel (k_fourier_cz2; measured
    // 112.8 ms against 108.8 ms, profiles/r02_fourier_cz.md); default: only the rows whose undivided work array does not fit one SM (dp rows
    // longer than ~5400 points, e.g. TCo2559 in double precision) -- on TCo1279 the pair kernel is no faster than the
    // undivided one (profiles/r02_fourier_cz.md: three CTAs per SM walk through 150 KB of unrolled code in different
    // phases and saturate the instruction cache, 84 % of the GPC instruction-fetch peak against 27 %)
    const char* czenv = getenv("ECT_FFT_CZ");          // read at every setup: tests switch it per handle
    const int cz_mode = czenv ? atoi(czenv) : -1;
    int smem_sm = 0;
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, d->dev);
    const int n_old = (int)d->buckets.size();
    for (int c = 1; c <= 6; ++c) {
        EctDevice::Bucket b;
        b.cz = 1; b.smem = 0; b.maxr = 7; b.nostage = 1;
        b.threads = c >= 3 ? 128 : (c == 2 ? 192 : 384);
        static const char* t256 = getenv("ECT_CZ_T256");       // experiment: 2 x 256 threads at 128 registers
        if (t256 && atoi(t256) && c >= 2) b.threads = 256;
        b.threads_inv = b.threads;
        d->buckets.push_back(b);
    }
    // cz_mode 2: chirp-z rows in k_fourier_cz2 (both halves in one CTA); classes by CTA size (small rows: several
    // small CTAs per SM) and staged / unstaged inputs
    const int n_cz1 = (int)d->buckets.size();
    for (int v = 0; v < 2; ++v)
        for (int thr : {128, 192, 384}) {
            EctDevice::Bucket b;
            b.cz = 2; b.smem = 0; b.maxr = 7; b.nostage = v; b.threads = thr; b.threads_inv = thr;
            d->buckets.push_back(b);
        }
    std::vector<int> need_of(P.nlat, 0);
    for (int l = 0; l < P.nlat; ++l) {
        const EctLatPlan& lp = d->fft.latplans[d->h_lat_plan[l]];
        bool placed = false;
        int need = 0, maxr = 2;
        if (lp.bluestein && cz_mode == 2) {
            const int H = lp.m / 2;
            const int ti = H >= 2048 ? 2 : (H >= 1024 ? 1 : 0);            // 384 / 192 / 128 threads
            for (int v = 0; v < 2 && !placed; ++v) {
                need = std::max(cz2_layout(true, H, lp.nlon, lp.km, csize, iosize, v != 0).total,
                                cz2_layout(false, H, lp.nlon, lp.km, csize, iosize, v != 0).total);
                if (need <= maxsm - 1024) { d->buckets[n_cz1 + 3 * v + ti].lats.push_back(l); placed = true; }
            }
        }
        if (!placed && !(lp.bluestein && cz_mode == 1)) {
            const EctFftPlan& pl = d->fft.plans[lp.plan];
            for (int s = 0; s < pl.nst; ++s) if (pl.radix[s] & 1) maxr = std::max(maxr, pl.radix[s]);   // 2,4,8,16 are in every variant
            const int nroots = maxr <= 7 ? ECT_ROOTS_OFF(8) : ECT_ROOTS_SIZE;
            for (auto& b : d->buckets) {
                if (b.cz) continue;
                need = std::max(ft_layout(true, lp.bluestein != 0, pl.n, lp.nlon, lp.km, nroots, csize, iosize, b.nostage).total,
                                ft_layout(false, lp.bluestein != 0, pl.n, lp.nlon, lp.km, nroots, csize, iosize, b.nostage).total);
                if (need <= b.smem && maxr <= b.maxr) { b.lats.push_back(l); placed = true; break; }
            }
        }
        if (!placed && lp.bluestein) {
            need = cz_smem_bytes(lp.m / 2, lp.nlon, csize);
            if (need <= maxsm - 1024) {
                const int c = std::max(1, std::min(6, smem_sm / (need + 1024 + 256)));
                d->buckets[n_old + c - 1].lats.push_back(l);
                placed = true;
            }
        }
        need_of[l] = need;
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
    FT_ATTR(true, 7, 384, false);
    FT_ATTR(true, 7, 256, false); FT_ATTR(false, 7, 256, false);
    FT_ATTR(true, ECT_MAX_RADIX, 256, false); FT_ATTR(false, ECT_MAX_RADIX, 256, false);
    FT_ATTR(true, 7, 256, true); FT_ATTR(false, 7, 256, true);
    FT_ATTR(true, 7, 512, true); FT_ATTR(false, 7, 512, true);
    FT_ATTR(true, ECT_MAX_RADIX, 256, true); FT_ATTR(false, ECT_MAX_RADIX, 256, true);
#undef FT_ATTR
#define CZ_ATTR(...) ECT_CUDA(cudaFuncSetAttribute(k_fourier_cz<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm - 1024))
    CZ_ATTR(true, 128, false); CZ_ATTR(false, 128, false); CZ_ATTR(true, 192, false); CZ_ATTR(false, 192, false);
    CZ_ATTR(true, 384, false); CZ_ATTR(false, 384, false);
    CZ_ATTR(true, 128, true); CZ_ATTR(false, 128, true); CZ_ATTR(true, 192, true); CZ_ATTR(false, 192, true);
    CZ_ATTR(true, 384, true); CZ_ATTR(false, 384, true);
    CZ_ATTR(true, 256, false); CZ_ATTR(false, 256, false); CZ_ATTR(true, 256, true); CZ_ATTR(false, 256, true);
#undef CZ_ATTR
#define CZ2_ATTR(...) ECT_CUDA(cudaFuncSetAttribute(k_fourier_cz2<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm - 1024))
    CZ2_ATTR(true, 192, false); CZ2_ATTR(false, 192, false); CZ2_ATTR(true, 96, false); CZ2_ATTR(false, 96, false);
    CZ2_ATTR(true, 64, false); CZ2_ATTR(false, 64, false);
    CZ2_ATTR(true, 192, true); CZ2_ATTR(false, 192, true); CZ2_ATTR(true, 96, true); CZ2_ATTR(false, 96, true);
    CZ2_ATTR(true, 64, true); CZ2_ATTR(false, 64, true);
#undef CZ2_ATTR
    return ECT_SUCCESS;
}

// debug: phase cycle stamps of block 0 (first pair) of the last Fourier launch
extern "C" int ect_debug_fft_probe(long long* out64) {
    ECT_CUDA(cudaDeviceSynchronize());
    ECT_CUDA(cudaMemcpyFromSymbol(out64, g_ft_probe, sizeof(long long) * 64));
    return ECT_SUCCESS;
}
