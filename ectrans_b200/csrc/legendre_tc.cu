// FP32-class Legendre contraction for sp handles on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// The reference's sp build runs LEINV / LEDIR as SGEMMs, on its GPU branch as 3xTF32 CUTLASS GEMMs
// (gpu/algor/hicblas_cutlass.cuda.h:41-72), with m = 0 kept in double precision
// (cpu/internal/ledir_mod.F90:133-171, gpu/internal/leinv_mod.F90:264-288).  Here:
//   * one kernel for both directions: D[M x N] = sum_k A[k][M] . B[k][N], both operands MN-major (the layouts the
//     tables P[k][lat] and the field operands [k][column] have in HBM), accumulators in TMEM, operands staged by TMA
//     tensor maps (cp.async.bulk.tensor, SWIZZLE_128B_ATOM_32B boxes of 32 floats x 16 k rows) into a 4-stage ring;
//   * 3xTF32: every operand is kept as hi = tf32(x) and lo = x - hi; D += hi.hi + hi.lo + lo.hi, three
//     tcgen05.mma.kind::tf32 per k step, fp32 accumulation (accuracy budget in DESIGN.md: 4-5e-7 relative L2);
//   * both parities of a tile in one CTA (256 of the 512 TMEM columns), so that the epilogue forms north = S + A /
//     south = S - A (ASRE1B) and writes the records of the consumer rank, exactly like k_leinv;
//   * the direct transform uses the same kernel on a transposed copy of the table (Pt[lat][k]) and on N +- S formed by
//     a small preparation kernel, so that no K-major descriptor is needed;
//   * m = 0 stays on the FP64 DMMA kernels (legendre.cu), as in the reference.
// Shared-memory / instruction descriptors are the ones the round-1 probe settled on B200 (tools/probes):
// layout type 1 (SWIZZLE_128B_BASE32B), LBO = stride of the 32-element MN chunks, SBO = 512 (4 k rows), a K = 8
// instruction consumes two groups.
#include "ect_internal.h"
#include <cuda.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>

#define TC_BM 128
#define TC_BN 128
#define TC_KC 16
#define TC_NST 4
#define TC_THREADS 128
#define TC_TILE_BYTES (TC_BM * TC_KC * 4)          // one operand tile: 128 (MN) x 16 (k) floats = 4 boxes of 2 KB
#define TC_STAGE_BYTES (4 * TC_TILE_BYTES)         // A hi, A lo, B hi, B lo
#define TC_SMEM_BYTES (TC_NST * TC_STAGE_BYTES + 1024)

struct EctTcM { long long arow0[2]; long long trow0[2]; };    // first row of (m, parity) in the inverse / transposed tables

struct alignas(64) TcMaps { CUtensorMap a[2]; CUtensorMap b[2][2]; };     // a[hi / lo], b[parity][hi / lo]

struct EctTcState {
    float *ainv[2] = {}, *adir[2] = {};         // [hi / lo]; inverse: rows (m, parity, k) x ldpu latitudes; direct: rows (m, parity, lat) x ldk
    long long ainv_rows = 0, adir_rows = 0;
    int ldpu = 0, ldk = 0;
    EctTcM* d_tcm = nullptr;
    float* b[4] = {};                           // operand arrays [rows][cp]: inverse X hi / lo; direct (N+S) hi / lo, (N-S) hi / lo
    long long b_rows = 0; int b_cp = 0;
    int2 *inv_tiles = nullptr, *dir_tiles = nullptr;
    int n_inv_tiles = 0, n_dir_tiles = 0;
    CUtensorMap map_ainv[2], map_adir[2];
    bool tables_ready = false;
    bool x_split_by_prologue = false;           // the prologue of the current call wrote b[0] / b[1] itself
    int enabled = -1;
};

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled g_encode = nullptr;

static int tc_make_map(CUtensorMap* map, const float* base, long long inner, long long rows, long long pitch_elems) {
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
            ect_set_error("tcgen05 path: cuTensorMapEncodeTiled not available");
            return ECT_ERR_CUDA;
        }
        g_encode = (PFN_tmapEncodeTiled)fn;
    }
    // The MN dimension is split into (32 floats, chunks of 32): one box {32, 16 k rows, 4 chunks} fetches a whole
    // 128 x 16 operand tile in the order the MN-major SWIZZLE_128B_BASE32B descriptor wants it ([chunk][k row][128 bytes]);
    // with a 2-D map the same tile takes four instructions, and the TMA issue rate of the single producer thread
    // (16 instructions per stage) dominated the first version (profiles/r02_sp_tc.md).
    const cuuint64_t dims[3] = {32, (cuuint64_t)std::max<long long>(rows, 1), (cuuint64_t)((inner + 31) / 32)};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch_elems * sizeof(float), 128};
    const cuuint32_t box[3] = {32, TC_KC, 4};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ect_set_error("tcgen05 path: cuTensorMapEncodeTiled failed (%d)", (int)r); return ECT_ERR_CUDA; }
    return ECT_SUCCESS;
}

// ------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned tc_s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ void tc_mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(tc_s32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(tc_s32(bar)), "r"(bytes) : "memory");
}
// Bounded: a transfer or MMA that never completes (a rejected descriptor) traps after ~2 s instead of hanging the GPU
__device__ __forceinline__ void tc_mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned addr = tc_s32(bar);
    const long long t0 = clock64();
    for (;;) {
        unsigned done;
        asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
// box {32 floats, 16 rows, 4 chunks} at (row y, first MN element x, a multiple of 32)
__device__ __forceinline__ void tc_tma_load(unsigned dst, const CUtensorMap* map, int x, int y, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
                 :: "r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(0), "r"(y), "r"(x >> 5), "r"(tc_s32(bar)) : "memory");
}
// shared-memory matrix descriptor, MN-major, SWIZZLE_128B_BASE32B (layout type 1)
__device__ __forceinline__ unsigned long long tc_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr >> 4) & 0x3fff);
    d |= (unsigned long long)((lbo >> 4) & 0x3fff) << 16;
    d |= (unsigned long long)((sbo >> 4) & 0x3fff) << 32;
    d |= 1ull << 46;
    d |= 1ull << 61;
    return d;
}
__device__ __forceinline__ void tc_mma(unsigned taddr, unsigned long long da, unsigned long long db, unsigned idesc, unsigned acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(taddr), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(tc_s32(bar)) : "memory");
}
__device__ __forceinline__ void tc_ld16(unsigned addr, float* v) {
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------------------------------
// operand preparation
// ------------------------------------------------------------------------------------------
// hi / lo float tables from the double table: inverse layout [(m, parity, k)][lat], direct layout [(m, parity, lat)][k]
__global__ void k_tc_tables(const EctLegM* __restrict__ legm, const EctTcM* __restrict__ tcm, const double* __restrict__ ptab,
                            float* ainv_hi, float* ainv_lo, float* adir_hi, float* adir_lo, int ldpu, int ldk) {
    const int ml = blockIdx.z >> 1, par = blockIdx.z & 1;
    const EctLegM lm = legm[ml];
    const EctTcM tm = tcm[ml];
    const int i = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    const int cnt = par ? lm.ila : lm.ils;
    if (k >= cnt || i >= lm.ndglu) return;
    const double p = ptab[(par ? lm.pa_off : lm.ps_off) + (long long)k * lm.ldp + i];
    const float hi = tc_tf32((float)p);
    const float lo = (float)(p - (double)hi);
    const long long ai = (tm.arow0[par] + k) * (long long)ldpu + i;
    ainv_hi[ai] = hi; ainv_lo[ai] = lo;
    const long long di = (tm.trow0[par] + i) * (long long)ldk + k;
    adir_hi[di] = hi; adir_lo[di] = lo;
}

// inverse: X[(m, n)][c] (double, from the prologue) -> hi / lo float rows in parity-split order (m, parity, k)
__global__ void k_tc_split_x(const EctLegM* __restrict__ legm, const double* __restrict__ x, int cp, int nsmax,
                             float* __restrict__ xh, float* __restrict__ xl) {
    const EctLegM lm = legm[blockIdx.y];
    if (lm.m == 0) return;                          // m = 0 runs on the FP64 kernels
    const int nrows = nsmax - lm.m + 2;
    for (int r = blockIdx.x * 4; r < min((int)(blockIdx.x + 1) * 4, nrows); ++r) {
        const int par = r & 1, k = r >> 1;
        const double* src = x + (lm.xrow0 + r) * (long long)cp;
        const long long dst = (lm.xrow0 + (par ? lm.ils : 0) + k) * (long long)cp;
        for (int c = threadIdx.x * 2; c < cp; c += blockDim.x * 2) {
            const double2 v = *reinterpret_cast<const double2*>(src + c);
            const float a = (float)v.x, b = (float)v.y;
            const float ah = tc_tf32(a), bh = tc_tf32(b);
            *reinterpret_cast<float2*>(xh + dst + c) = make_float2(ah, bh);
            *reinterpret_cast<float2*>(xl + dst + c) = make_float2(a - ah, b - bh);
        }
    }
}

// direct: PRFI2B (prfi2b_mod.F90:91-92): N + S and N - S of the (already weighted) records, as hi / lo float rows (m, lat)
__global__ void k_tc_prep_dir(const EctLegM* __restrict__ legm, const double* __restrict__ fb, const int* __restrict__ rec_n,
                              const int* __restrict__ rec_s, int cp, float* __restrict__ fsh, float* __restrict__ fsl,
                              float* __restrict__ fah, float* __restrict__ fal) {
    const EctLegM lm = legm[blockIdx.y];
    if (lm.m == 0) return;
    for (int i = blockIdx.x * 4; i < min((int)(blockIdx.x + 1) * 4, lm.ndglu); ++i) {
        const double* pn = fb + (long long)rec_n[lm.rec0 + i] * cp;
        const double* ps = fb + (long long)rec_s[lm.rec0 + i] * cp;
        const long long dst = (lm.rec0 + i) * (long long)cp;
        for (int c = threadIdx.x * 2; c < cp; c += blockDim.x * 2) {
            const double2 n = *reinterpret_cast<const double2*>(pn + c), s = *reinterpret_cast<const double2*>(ps + c);
            const float s0 = (float)(n.x + s.x), s1 = (float)(n.y + s.y), a0 = (float)(n.x - s.x), a1 = (float)(n.y - s.y);
            const float s0h = tc_tf32(s0), s1h = tc_tf32(s1), a0h = tc_tf32(a0), a1h = tc_tf32(a1);
            *reinterpret_cast<float2*>(fsh + dst + c) = make_float2(s0h, s1h);
            *reinterpret_cast<float2*>(fsl + dst + c) = make_float2(s0 - s0h, s1 - s1h);
            *reinterpret_cast<float2*>(fah + dst + c) = make_float2(a0h, a1h);
            *reinterpret_cast<float2*>(fal + dst + c) = make_float2(a0 - a0h, a1 - a1h);
        }
    }
}

// ------------------------------------------------------------------------------------------
// the contraction
// ------------------------------------------------------------------------------------------
struct TcArgs {
    const EctLegM* legm; const EctTcM* tcm;
    const int2* tiles; int nct;      // (local m, tile along M); column tiles
    int cp;
    double* x;                       // direct: output POA rows (m, n)
    // inverse epilogue (TRMTOL fused, as k_leinv)
    double* const* peer; const int* dst_rank_n; const int* dst_rank_s; const int* dst_rec_n; const int* dst_rec_s;
};

// Accuracy: the tensor core adds into its fp32 accumulator with truncation, a bias of about half an ulp per MMA
// instruction (measured: relative error 0.8e-6 at T47, 2.8e-6 at T399 with all three products in one accumulator).  So
//   * the dominant hi.hi products have an accumulator of their own (one instruction per K = 8 step); the hi.lo and lo.hi
//     corrections, 2^-11 smaller, go to a second one -- all 512 TMEM columns: {main, correction} x {symmetric,
//     antisymmetric} x 128; the epilogue adds them in fp32 with rounding;
//   * K is cut into segments of TC_KSEG; every segment ends with an epilogue, later segments add to what is stored.
#define TC_KSEG 1024
#define TC_EPI_LD 68          // pitch (floats) of the epilogue staging rows: 64 columns + 4, 16-byte stores of a quarter warp hit 8 distinct 16-byte bank groups

template <bool DIRECT>
__global__ void __launch_bounds__(TC_THREADS, 1) k_leg_tc(const __grid_constant__ TcMaps maps, TcArgs a) {
    extern __shared__ unsigned char tc_smem_raw[];
    __shared__ __align__(8) unsigned long long s_full[TC_NST], s_empty[TC_NST], s_acc;
    __shared__ unsigned s_tmem;
    const int tile = blockIdx.x / a.nct, ct = blockIdx.x - tile * a.nct;
    const int2 td = a.tiles[tile];
    const EctLegM lm = a.legm[td.x];
    const EctTcM tm = a.tcm[td.x];
    const int m0 = td.y * TC_BM, c0 = ct * TC_BN;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned sbase = (tc_s32(tc_smem_raw) + 1023u) & ~1023u;
    // K extent and chunk count per parity (0: n - m even / symmetric, 1: odd / antisymmetric)
    const int cnt[2] = {lm.ils, lm.ila};
    const int nk0 = DIRECT ? lm.ndglu : lm.ils, nk1 = DIRECT ? (lm.ila > 0 ? lm.ndglu : 0) : lm.ila;
    const int nch0 = (nk0 + TC_KC - 1) / TC_KC, nch1 = (nk1 + TC_KC - 1) / TC_KC;
    constexpr int SEGCH = TC_KSEG / TC_KC;
    const int nseg = (max(nch0, nch1) + SEGCH - 1) / SEGCH;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TC_NST; ++s) { tc_mbar_init(&s_full[s], 1); tc_mbar_init(&s_empty[s], 1); }
        tc_mbar_init(&s_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" :: "r"(tc_s32(&s_tmem)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const unsigned taddr = s_tmem;
    const int r = 32 * warp + lane;                       // accumulator row = TMEM lane
    // inverse: record addresses of this tile's rows (TRMTOL fused: the buffer of the rank that owns the latitude), looked
    // up once -- three dependent loads per row inside the store loop cost more than the contraction
    __shared__ double* s_pn[TC_BM];
    __shared__ double* s_ps[TC_BM];
    if (!DIRECT) {
        const int li = m0 + r;
        if (li < lm.ndglu) {
            s_pn[r] = a.peer[a.dst_rank_n[lm.rec0 + li]] + (long long)a.dst_rec_n[lm.rec0 + li] * a.cp;
            s_ps[r] = a.peer[a.dst_rank_s[lm.rec0 + li]] + (long long)a.dst_rec_s[lm.rec0 + li] * a.cp;
        }
    }
    const unsigned lane_addr = taddr + ((unsigned)(32 * warp) << 16);
    int it_base = 0;                                      // ring position carried across segments (same in both roles)

    for (int seg = 0; seg < nseg; ++seg) {
        const int lo0 = min(seg * SEGCH, nch0), hi0 = min((seg + 1) * SEGCH, nch0);
        const int lo1 = min(seg * SEGCH, nch1), hi1 = min((seg + 1) * SEGCH, nch1);
        const int n0 = hi0 - lo0, n1 = hi1 - lo1, nit = n0 + n1;
        if (warp == 0) {
            if (lane == 0) {
                // ---- TMA producer: one box (32 floats x 16 k rows x 4 chunks along MN) per operand tile, four per stage ----
                for (int i = 0; i < nit; ++i) {
                    const int it = it_base + i, st = it % TC_NST, par = i >= n0 ? 1 : 0, ch = par ? lo1 + i - n0 : lo0 + i;
                    tc_mbar_wait(&s_empty[st], (unsigned)(((it / TC_NST) & 1) ^ 1));
                    tc_mbar_expect_tx(&s_full[st], TC_STAGE_BYTES);
                    const int k0 = ch * TC_KC;
                    const int ya = (int)((DIRECT ? tm.trow0[par] : tm.arow0[par]) + k0);
                    const int yb = DIRECT ? (int)(lm.rec0 + k0) : (int)(lm.xrow0 + (par ? lm.ils : 0) + k0);
                    const unsigned dst = sbase + (unsigned)st * TC_STAGE_BYTES;
                    tc_tma_load(dst + 0 * TC_TILE_BYTES, &maps.a[0], m0, ya, &s_full[st]);
                    tc_tma_load(dst + 1 * TC_TILE_BYTES, &maps.a[1], m0, ya, &s_full[st]);
                    tc_tma_load(dst + 2 * TC_TILE_BYTES, &maps.b[par][0], c0, yb, &s_full[st]);
                    tc_tma_load(dst + 3 * TC_TILE_BYTES, &maps.b[par][1], c0, yb, &s_full[st]);
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            if (lane == 0) {
                // ---- MMA issuer: hi.hi into the main accumulator, hi.lo and lo.hi into the correction accumulator ----
                const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                       ((unsigned)(TC_BN >> 3) << 17) | ((unsigned)(TC_BM >> 4) << 24);
                for (int i = 0; i < nit; ++i) {
                    const int it = it_base + i, st = it % TC_NST, par = i >= n0 ? 1 : 0, first = par ? (i == n0) : (i == 0);
                    tc_mbar_wait(&s_full[st], (unsigned)((it / TC_NST) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;\n");
                    const unsigned base = sbase + (unsigned)st * TC_STAGE_BYTES;
                    const unsigned acc_main = taddr + (unsigned)(par * TC_BN), acc_corr = taddr + (unsigned)(2 * TC_BN + par * TC_BN);
#pragma unroll
                    for (int ks = 0; ks < TC_KC / 8; ++ks) {
                        const unsigned off = (unsigned)ks * 1024u;          // two groups of 4 k rows
                        const unsigned long long ah = tc_desc(base + 0 * TC_TILE_BYTES + off, TC_KC * 128, 512);
                        const unsigned long long al = tc_desc(base + 1 * TC_TILE_BYTES + off, TC_KC * 128, 512);
                        const unsigned long long bh = tc_desc(base + 2 * TC_TILE_BYTES + off, TC_KC * 128, 512);
                        const unsigned long long bl = tc_desc(base + 3 * TC_TILE_BYTES + off, TC_KC * 128, 512);
                        const unsigned acc = (!first || ks > 0) ? 1u : 0u;
                        tc_mma(acc_main, ah, bh, idesc, acc);
                        tc_mma(acc_corr, ah, bl, idesc, acc);
                        tc_mma(acc_corr, al, bh, idesc, 1u);
                    }
                    tc_commit(&s_empty[st]);         // the stage is free once these MMAs have read it
                }
                tc_commit(&s_acc);                   // the accumulators of this segment are complete
            }
            __syncwarp();
        }
        it_base += nit;
        // ---- epilogue: TMEM -> registers -> records (inverse: ASRE1B, asre1b_mod.F90:99-100) / POA rows (direct) ----
        tc_mbar_wait(&s_acc, (unsigned)(seg & 1));
        asm volatile("tcgen05.fence::after_thread_sync;\n");
        const bool add = seg > 0;
        // Accumulator rows live one per thread (TMEM lane = row); stored like that every lane would write its own
        // record (one 32-byte sector per lane and instruction).  The rows are staged through the (now idle) operand ring,
        // 64 columns at a time, and leave as whole 512-byte pieces of a record per warp instruction.
        float* stS = reinterpret_cast<float*>(tc_smem_raw + (sbase - tc_s32(tc_smem_raw)));
        float* stA = stS + TC_BM * TC_EPI_LD;
        for (int ch0 = 0; ch0 < TC_BN && c0 + ch0 < a.cp; ch0 += 64) {
#pragma unroll
            for (int c = 0; c < 64; c += 16) {
                float s[16], v[16], t[16];
                tc_ld16(lane_addr + (unsigned)(ch0 + c), s);
                tc_ld16(lane_addr + (unsigned)(2 * TC_BN + ch0 + c), t);
#pragma unroll
                for (int j = 0; j < 16; ++j) s[j] += t[j];
                if (n1 > 0) {
                    tc_ld16(lane_addr + (unsigned)(TC_BN + ch0 + c), v);
                    tc_ld16(lane_addr + (unsigned)(3 * TC_BN + ch0 + c), t);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += t[j];
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    *reinterpret_cast<float4*>(stS + r * TC_EPI_LD + c + j) = make_float4(s[j], s[j + 1], s[j + 2], s[j + 3]);
                    *reinterpret_cast<float4*>(stA + r * TC_EPI_LD + c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
            }
            __syncthreads();
            const int col = c0 + ch0 + 2 * lane;
            for (int row = warp; row < TC_BM; row += 4) {
                const float2 sv = *reinterpret_cast<const float2*>(stS + row * TC_EPI_LD + 2 * lane);
                const float2 av = *reinterpret_cast<const float2*>(stA + row * TC_EPI_LD + 2 * lane);
                if (col >= a.cp) continue;
                if (!DIRECT) {
                    // north = S + A, south = S - A (asre1b_mod.F90:99-100); TRMTOL fused: the record of the consumer rank
                    const int li = m0 + row;
                    if (li >= lm.ndglu) break;
                    double2* qn = reinterpret_cast<double2*>(s_pn[row] + col);
                    double2* qs = reinterpret_cast<double2*>(s_ps[row] + col);
                    double2 on = make_double2((double)(sv.x + av.x), (double)(sv.y + av.y));
                    double2 os = make_double2((double)(sv.x - av.x), (double)(sv.y - av.y));
                    if (add) { const double2 pn_ = *qn, ps_ = *qs; on.x += pn_.x; on.y += pn_.y; os.x += ps_.x; os.y += ps_.y; }
                    *qn = on; *qs = os;
                } else {
                    // symmetric part -> rows n - m even, antisymmetric -> odd (ledir_mod.F90:174-179, :248-253)
                    const int k = m0 + row;
                    if (k >= cnt[0]) break;
                    double2* q0 = reinterpret_cast<double2*>(a.x + (lm.xrow0 + 2 * k) * (long long)a.cp + col);
                    double2 o0 = make_double2((double)sv.x, (double)sv.y);
                    if (add) { const double2 p_ = *q0; o0.x += p_.x; o0.y += p_.y; }
                    *q0 = o0;
                    if (k < cnt[1]) {
                        double2* q1 = reinterpret_cast<double2*>(a.x + (lm.xrow0 + 2 * k + 1) * (long long)a.cp + col);
                        double2 o1 = make_double2((double)av.x, (double)av.y);
                        if (add) { const double2 p_ = *q1; o1.x += p_.x; o1.y += p_.y; }
                        *q1 = o1;
                    }
                }
            }
            __syncthreads();
        }
        // the staging writes (generic proxy) precede the next segment's TMA writes (async proxy) to the same bytes
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n");
        __syncthreads();                                  // every warp has read its accumulator rows: the next segment may overwrite them
        asm volatile("tcgen05.fence::after_thread_sync;\n");
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" :: "r"(taddr));
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static EctTcState* tc_state(EctDevice* d) {
    if (!d->tc) d->tc = new EctTcState();
    return d->tc;
}

// sp handles use the tensor-core path unless ECT_SP_TC=0 (then every m runs on the FP64 DMMA kernels, as in round 1)
static int tc_build_tables(EctHandle* h);
static long long tc_table_bytes(EctHandle* h);
bool ect_tc_enabled(EctHandle* h) {
    if (h->precision != ECT_PREC_SP || !h->d || h->hp.nump == 0) return false;
    EctTcState* t = tc_state(h->d);
    if (t->enabled < 0) {
        const char* e = getenv("ECT_SP_TC");
        t->enabled = (e && atoi(e) == 0) ? 0 : 1;
        if (t->enabled) {
            // the hi / lo float tables (both orientations) live next to the double table: without the memory for them
            // (TCo2559 on ONE GPU: 137 GB) the handle stays on the FP64 kernels
            size_t fr = 0, tot = 0;
            cudaMemGetInfo(&fr, &tot);
            if ((long long)fr < tc_table_bytes(h) + (8ll << 30)) t->enabled = 0;
        }
    }
    if (t->enabled == 1 && !t->tables_ready && tc_build_tables(h) != ECT_SUCCESS) { cudaGetLastError(); t->enabled = 0; }
    return t->enabled == 1;
}

void ect_tc_invalidate(EctHandle* h) { if (h->d && h->d->tc) h->d->tc->tables_ready = false; }

void ect_tc_free(EctDevice* d) {
    EctTcState* t = d->tc;
    if (!t) return;
    for (int i = 0; i < 2; ++i) { if (t->ainv[i]) cudaFree(t->ainv[i]); if (t->adir[i]) cudaFree(t->adir[i]); }
    for (int i = 0; i < 4; ++i) if (t->b[i]) cudaFree(t->b[i]);
    if (t->d_tcm) cudaFree(t->d_tcm);
    if (t->inv_tiles) cudaFree(t->inv_tiles);
    if (t->dir_tiles) cudaFree(t->dir_tiles);
    delete t;
    d->tc = nullptr;
}

static long long tc_table_bytes(EctHandle* h) {
    long long ar = 0, tr = 0; int maxdglu = 0, maxk = 0;
    auto pad = [](long long v) { return (v + TC_KC - 1) / TC_KC * TC_KC; };
    for (const EctLegM& lm : h->d->h_legm) {
        ar += pad(lm.ils) + pad(lm.ila); tr += 2 * pad(lm.ndglu);
        maxdglu = std::max(maxdglu, lm.ndglu); maxk = std::max(maxk, lm.ils);
    }
    const long long ldpu = (maxdglu + TC_BM - 1) / TC_BM * TC_BM, ldk = (maxk + TC_BM - 1) / TC_BM * TC_BM;
    return 2 * sizeof(float) * ((ar + TC_KC) * ldpu + (tr + TC_KC) * ldk);
}

static int tc_build_tables(EctHandle* h) {
    EctDevice* d = h->d;
    EctTcState* t = tc_state(d);
    const EctHostPlan& P = h->hp;
    const int nump = P.nump;
    if (!t->d_tcm) {
        std::vector<EctTcM> tcm(nump);
        long long ar = 0, tr = 0;
        int maxdglu = 0, maxk = 0;
        auto pad = [](long long v) { return (v + TC_KC - 1) / TC_KC * TC_KC; };
        for (int ml = 0; ml < nump; ++ml) {
            const EctLegM& lm = d->h_legm[ml];
            tcm[ml].arow0[0] = ar; ar += pad(lm.ils);
            tcm[ml].arow0[1] = ar; ar += pad(lm.ila);
            tcm[ml].trow0[0] = tr; tr += pad(lm.ndglu);
            tcm[ml].trow0[1] = tr; tr += pad(lm.ndglu);
            maxdglu = std::max(maxdglu, lm.ndglu); maxk = std::max(maxk, lm.ils);
        }
        t->ainv_rows = ar + TC_KC; t->adir_rows = tr + TC_KC;
        t->ldpu = (maxdglu + TC_BM - 1) / TC_BM * TC_BM;
        t->ldk = (maxk + TC_BM - 1) / TC_BM * TC_BM;
        ECT_CUDA(cudaMalloc(&t->d_tcm, nump * sizeof(EctTcM)));
        ECT_CUDA(cudaMemcpy(t->d_tcm, tcm.data(), nump * sizeof(EctTcM), cudaMemcpyHostToDevice));
        for (int i = 0; i < 2; ++i) {
            ECT_CUDA(cudaMalloc(&t->ainv[i], (size_t)t->ainv_rows * t->ldpu * sizeof(float)));
            ECT_CUDA(cudaMalloc(&t->adir[i], (size_t)t->adir_rows * t->ldk * sizeof(float)));
        }
        // tile lists (m > 0), heaviest first
        std::vector<int> order(nump);
        for (int i = 0; i < nump; ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](int x, int y) { return P.myms[x] < P.myms[y]; });
        std::vector<int2> inv, dir;
        for (int ml : order) {
            const EctLegM& lm = d->h_legm[ml];
            if (lm.m == 0 || lm.ndglu == 0) continue;
            for (int it = 0; it < (lm.ndglu + TC_BM - 1) / TC_BM; ++it) inv.push_back(make_int2(ml, it));
            for (int kt = 0; kt < (lm.ils + TC_BM - 1) / TC_BM; ++kt) dir.push_back(make_int2(ml, kt));
        }
        t->n_inv_tiles = (int)inv.size(); t->n_dir_tiles = (int)dir.size();
        ECT_CUDA(cudaMalloc(&t->inv_tiles, std::max<size_t>(inv.size(), 1) * sizeof(int2)));
        ECT_CUDA(cudaMalloc(&t->dir_tiles, std::max<size_t>(dir.size(), 1) * sizeof(int2)));
        ECT_CUDA(cudaMemcpy(t->inv_tiles, inv.data(), inv.size() * sizeof(int2), cudaMemcpyHostToDevice));
        ECT_CUDA(cudaMemcpy(t->dir_tiles, dir.data(), dir.size() * sizeof(int2), cudaMemcpyHostToDevice));
        int rc;
        for (int i = 0; i < 2; ++i) {
            if ((rc = tc_make_map(&t->map_ainv[i], t->ainv[i], t->ldpu, t->ainv_rows, t->ldpu))) return rc;
            if ((rc = tc_make_map(&t->map_adir[i], t->adir[i], t->ldk, t->adir_rows, t->ldk))) return rc;
        }
        ECT_CUDA(cudaFuncSetAttribute(k_leg_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        ECT_CUDA(cudaFuncSetAttribute(k_leg_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    }
    // zero padding rows / columns multiply whatever finite operand rows the boxes overlap
    for (int i = 0; i < 2; ++i) {
        ECT_CUDA(cudaMemsetAsync(t->ainv[i], 0, (size_t)t->ainv_rows * t->ldpu * sizeof(float), d->stream));
        ECT_CUDA(cudaMemsetAsync(t->adir[i], 0, (size_t)t->adir_rows * t->ldk * sizeof(float), d->stream));
    }
    int maxdglu = 0, maxk = 0;
    for (auto& lm : d->h_legm) { maxdglu = std::max(maxdglu, lm.ndglu); maxk = std::max(maxk, lm.ils); }
    if (maxdglu > 0) {
        dim3 grid((maxdglu + 127) / 128, maxk, 2 * nump);
        k_tc_tables<<<grid, 128, 0, d->stream>>>(d->legm, t->d_tcm, d->ptab, t->ainv[0], t->ainv[1], t->adir[0], t->adir[1], t->ldpu, t->ldk);
        ECT_CUDA(cudaGetLastError());
    }
    t->tables_ready = true;
    return ECT_SUCCESS;
}

static int tc_ensure_operands(EctHandle* h, int cp) {
    EctDevice* d = h->d;
    EctTcState* t = tc_state(d);
    const long long rows = std::max<long long>(d->xrows, (long long)h->hp.mrow0[h->hp.nump]) + 2 * TC_KC;
    if (cp <= t->b_cp && rows <= t->b_rows && t->b[0]) return ECT_SUCCESS;
    ECT_CUDA(cudaStreamSynchronize(d->stream));
    for (int i = 0; i < 4; ++i) {
        if (t->b[i]) { ECT_CUDA(cudaFree(t->b[i])); t->b[i] = nullptr; }
        ECT_CUDA(cudaMalloc(&t->b[i], (size_t)rows * cp * sizeof(float) + 256));
        ECT_CUDA(cudaMemsetAsync(t->b[i], 0, (size_t)rows * cp * sizeof(float), d->stream));      // never NaN patterns under zero table rows
    }
    t->b_rows = rows; t->b_cp = cp;
    return ECT_SUCCESS;
}

static void tc_fill_args(EctHandle* h, const EctFieldCfg& f, TcArgs& a) {
    EctDevice* d = h->d;
    EctTcState* t = d->tc;
    a.legm = d->legm; a.tcm = t->d_tcm; a.cp = f.cp; a.nct = (f.cp + TC_BN - 1) / TC_BN;
    a.x = d->xwork;
    a.peer = d->peer_fft; a.dst_rank_n = d->leg_dst_rank_n; a.dst_rank_s = d->leg_dst_rank_s;
    a.dst_rec_n = d->leg_dst_rec_n; a.dst_rec_s = d->leg_dst_rec_s;
}

// Operand rows of the inverse contraction (hi / lo float rows, parity-split order): written by k_ltinv_prologue itself
int ect_tc_operands(EctHandle* h, int cp, float** xh, float** xl) {
    int rc = tc_ensure_operands(h, cp);
    if (rc) return rc;
    EctTcState* t = h->d->tc;
    *xh = t->b[0]; *xl = t->b[1];
    t->x_split_by_prologue = true;
    return ECT_SUCCESS;
}

// LEINV for the wavenumbers m > 0 of an sp handle (m = 0: k_leinv on its own tiles, see ect_launch_leinv)
int ect_tc_launch_leinv(EctHandle* h, const EctFieldCfg& f) {
    EctDevice* d = h->d;
    EctTcState* t = tc_state(d);
    int rc;
    if (!t->tables_ready && (rc = tc_build_tables(h))) return rc;
    if ((rc = tc_ensure_operands(h, f.cp))) return rc;
    if (t->n_inv_tiles == 0) return ECT_SUCCESS;
    if (!t->x_split_by_prologue) {      // (callers that fill the double rows themselves)
        dim3 g((h->hp.nsmax + 2 + 3) / 4, h->hp.nump);
        k_tc_split_x<<<g, 128, 0, d->stream>>>(d->legm, d->xwork, f.cp, h->hp.nsmax, t->b[0], t->b[1]);
        d->launches++;
    }
    t->x_split_by_prologue = false;
    TcMaps maps;
    maps.a[0] = t->map_ainv[0]; maps.a[1] = t->map_ainv[1];
    for (int hl = 0; hl < 2; ++hl) {
        if ((rc = tc_make_map(&maps.b[0][hl], t->b[hl], f.cp, t->b_rows, f.cp))) return rc;
        maps.b[1][hl] = maps.b[0][hl];
    }
    TcArgs a;
    tc_fill_args(h, f, a);
    a.tiles = t->inv_tiles;
    k_leg_tc<false><<<(unsigned)((long long)t->n_inv_tiles * a.nct), TC_THREADS, TC_SMEM_BYTES, d->stream>>>(maps, a);
    d->launches += 1;
    ECT_CUDA(cudaGetLastError());
    return ECT_SUCCESS;
}

// LEDIR for m > 0 of an sp handle
int ect_tc_launch_ledir(EctHandle* h, const EctFieldCfg& f) {
    EctDevice* d = h->d;
    EctTcState* t = tc_state(d);
    int rc;
    if (!t->tables_ready && (rc = tc_build_tables(h))) return rc;
    if ((rc = tc_ensure_operands(h, f.cp))) return rc;
    if (t->n_dir_tiles == 0) return ECT_SUCCESS;
    int maxdglu = 0;
    for (auto& lm : d->h_legm) maxdglu = std::max(maxdglu, lm.ndglu);
    dim3 g((maxdglu + 3) / 4, h->hp.nump);
    k_tc_prep_dir<<<g, 128, 0, d->stream>>>(d->legm, d->fbuf_leg, d->leg_rec_n, d->leg_rec_s, f.cp, t->b[0], t->b[1], t->b[2], t->b[3]);
    TcMaps maps;
    maps.a[0] = t->map_adir[0]; maps.a[1] = t->map_adir[1];
    for (int par = 0; par < 2; ++par)
        for (int hl = 0; hl < 2; ++hl)
            if ((rc = tc_make_map(&maps.b[par][hl], t->b[2 * par + hl], f.cp, t->b_rows, f.cp))) return rc;
    TcArgs a;
    tc_fill_args(h, f, a);
    a.tiles = t->dir_tiles;
    k_leg_tc<true><<<(unsigned)((long long)t->n_dir_tiles * a.nct), TC_THREADS, TC_SMEM_BYTES, d->stream>>>(maps, a);
    d->launches += 2;
    ECT_CUDA(cudaGetLastError());
    return ECT_SUCCESS;
}
