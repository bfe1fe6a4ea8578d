// tcgen05 TF32 probe (preparation for the sp Legendre contraction; NOT part of the library).
// RESULT on B200 (round 1, last GPU seconds): MATCH (rel. error 6.5e-8) with layout type 1 = SWIZZLE_128B_BASE32B,
// swizzle_fill = 2 (atoms of 32 floats x 4 k rows, Swizzle<2,5,2>), LBO = stride of the 32-element MN chunks, SBO = stride of
// the 4-row K groups, two groups per K = 8 instruction; with layout type 2 (SWIZZLE_128B, the first guess described below)
// the MMA runs and writes zeros; K-major bits on MN-major descriptors fault the launch.
//
// One CTA computes D[128 x N] = sum_k A[k][m] * B[k][n] with both operands MN-major in shared memory -- the layouts the
// Legendre tables (P[k][lat], latitude contiguous) and the spectral / Fourier operands ([k][column]) have in HBM -- through
// tcgen05.mma.cta_group::1.kind::tf32 with the accumulator in TMEM, and writes D and a status word back.  The host side
// (tools/probes/tcgen05_probe.py) compares with a CPU product of the TF32-rounded inputs for a list of descriptor
// variants, so that the first GPU call of the next round settles the encodings instead of guessing them one by one.
//
// Encodings follow the vendored CUTLASS headers (cute/arch/mma_sm100_desc.hpp, cute/atom/mma_traits_sm100.hpp:167-260):
//   shared-memory descriptor: start >> 4 in bits [0,14), LBO >> 4 in [16,30), SBO >> 4 in [32,46), version 1 in [46,48),
//     layout type in [61,64) (2 = SWIZZLE_128B);  Major-MN, 128B swizzle, in 16-byte units:
//     ((8, n), (8, k)) : ((1, LBO), (8, SBO))  -- 32 floats of MN contiguous (128 B), the next 32 at LBO, the 8 k rows of a
//     group 128 B apart, the next group at SBO; Swizzle<3,4,3>: 16-byte chunk index ^= k row (address bits [4,7) ^= [7,10)).
//   instruction descriptor (32 bit): c_format F32 = 1 at [4,6), a/b_format TF32 = 2 at [7,10) / [10,13), a_major / b_major
//     (1 = MN) at 15 / 16, N >> 3 at [17,23), M >> 4 at [24,29).
// Every wait is bounded (the status word says which one gave up) so that a wrong guess cannot hang the GPU.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define PROBE_M 128
#define PROBE_MAXN 256
#define PROBE_MAXK 64

struct ProbeArgs {
    const float* a;      // [K][128]  MN-major A (m contiguous)
    const float* b;      // [K][N]    MN-major B (n contiguous)
    float* d;            // [128][N]
    int* status;         // 0 ok, 1 mbarrier wait gave up, 2 bad TMEM address
    int n, k;            // N multiple of 32 (<= 256), K multiple of 8 (<= 64)
    unsigned lbo_a, sbo_a, lbo_b, sbo_b;     // bytes; also the strides the fill uses
    unsigned layout_type;                    // 2 = SWIZZLE_128B
    unsigned a_major, b_major;               // 1 = MN
    int swizzle_fill;                        // 1: the fill applies Swizzle<3,4,3>
    int diag;                                // bit 0: operands are all ones (D = K whatever the layout); bit 1: the accumulator is
                                             // pre-set to 7.0 with tcgen05.st (left untouched = the MMA did not write)
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned long long make_desc(unsigned saddr, unsigned lbo, unsigned sbo, unsigned layout) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr >> 4) & 0x3fff);
    d |= (unsigned long long)((lbo >> 4) & 0x3fff) << 16;
    d |= (unsigned long long)((sbo >> 4) & 0x3fff) << 32;
    d |= 1ull << 46;                                  // descriptor version (Blackwell)
    d |= (unsigned long long)(layout & 7) << 61;
    return d;
}

// byte offset of element (mn, k) of an MN-major operand tile: chunks of 32 along MN at `lbo`, groups of 8 k rows at `sbo`
__device__ __forceinline__ unsigned tile_off(int mn, int k, unsigned lbo, unsigned sbo, int swz) {
    if (swz == 2) {      // SWIZZLE_128B_BASE32B (layout type 1): groups of 4 k rows, Swizzle<2,5,2>: 32-byte chunk ^= row
        const unsigned row = (unsigned)(k & 3);
        unsigned inrow = (unsigned)(mn & 31) * 4u;
        inrow ^= row << 5;
        return (unsigned)(mn >> 5) * lbo + (unsigned)(k >> 2) * sbo + row * 128u + inrow;
    }
    const unsigned row = (unsigned)(k & 7);
    unsigned inrow = (unsigned)(mn & 31) * 4u;        // byte inside the 128-byte row
    if (swz) inrow ^= row << 4;                       // Swizzle<3,4,3>: 16-byte chunk ^= row
    return (unsigned)(mn >> 5) * lbo + (unsigned)(k >> 3) * sbo + row * 128u + inrow;
}

extern "C" __global__ void __launch_bounds__(128, 1) k_tcgen05_tf32_probe(ProbeArgs p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // operand regions, 1024-byte aligned (swizzle atoms): A first, then B
    const unsigned a_bytes = (unsigned)(PROBE_M / 32) * max(p.lbo_a, 1024u) + (unsigned)(p.k / 8) * max(p.sbo_a, 1024u);
    unsigned char* sa = smem;
    unsigned char* sb = smem + ((a_bytes + 1023u) & ~1023u);
    for (int i = tid; i < p.k * PROBE_M; i += blockDim.x) {
        const int k = i / PROBE_M, m = i - k * PROBE_M;
        *reinterpret_cast<float*>(sa + tile_off(m, k, p.lbo_a, p.sbo_a, p.swizzle_fill)) = (p.diag & 1) ? 1.0f : p.a[i];
    }
    for (int i = tid; i < p.k * p.n; i += blockDim.x) {
        const int k = i / p.n, n = i - k * p.n;
        *reinterpret_cast<float*>(sb + tile_off(n, k, p.lbo_b, p.sbo_b, p.swizzle_fill)) = (p.diag & 1) ? 1.0f : p.b[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    // generic-proxy writes of the operands -> visible to the tensor core (async proxy)
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    if (tid == 0) tmem_base = 0xdeadbeefu;      // an allocation that does not write its address shows up in status[1]
    __syncthreads();
    if (warp == 0) {       // TMEM: N fp32 columns x 128 lanes, power of two >= 32
        unsigned cols = 32; while ((int)cols < p.n) cols <<= 1;
        if (cols == 32) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;\n" :: "r"(smem_u32(&tmem_base)));
        else if (cols == 64) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;\n" :: "r"(smem_u32(&tmem_base)));
        else if (cols == 128) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;\n" :: "r"(smem_u32(&tmem_base)));
        else asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" :: "r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const unsigned taddr = tmem_base;
    if (tid == 0) p.status[1] = (int)taddr;
    if (p.diag & 2) {      // sentinel in every accumulator element this warp can reach
        const unsigned sv = __float_as_uint(7.0f);
        for (int c0 = 0; c0 < p.n; c0 += 8) {
            const unsigned addr = taddr + ((unsigned)(warp * 32) << 16) + (unsigned)c0;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};\n" :: "r"(addr), "r"(sv) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        {   // read the sentinel straight back (status[2]: warp 0, status[3]: warp 1)
            unsigned r[8];
            const unsigned addr = taddr + ((unsigned)(warp * 32) << 16);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr));
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            if (lane == 0 && warp < 2) p.status[2 + warp] = (int)r[0];
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;\n");
    }
    if (tid == 0) {
        // instruction descriptor: F32 accumulate, TF32 x TF32, majors as given, N, M = 128
        const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((p.a_major & 1u) << 15) | ((p.b_major & 1u) << 16) |
                               ((unsigned)(p.n >> 3) << 17) | ((unsigned)(PROBE_M >> 4) << 24);
        for (int ks = 0; ks < p.k / 8; ++ks) {        // one MMA per group of 8 k rows
            const unsigned long long da = make_desc(smem_u32(sa) + (unsigned)ks * p.sbo_a * (p.swizzle_fill == 2 ? 2u : 1u), p.lbo_a, p.sbo_a, p.layout_type);
            const unsigned long long db = make_desc(smem_u32(sb) + (unsigned)ks * p.sbo_b * (p.swizzle_fill == 2 ? 2u : 1u), p.lbo_b, p.sbo_b, p.layout_type);
            const unsigned acc = ks > 0 ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                         :: "r"(taddr), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        // completion of all MMAs above -> one arrival on the mbarrier (implies tcgen05.fence::before_thread_sync)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(&bar)) : "memory");
    }
    // bounded wait for phase 0 (test_wait never blocks: 2^22 polls are a fraction of a second)
    int ok = 0;
    for (int spin = 0; spin < (1 << 22) && !ok; ++spin) {
        unsigned done;
        asm volatile("{\n\t.reg .pred q;\n\tmbarrier.test_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        ok = (int)done;
    }
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    if (!ok) { if (tid == 0) *p.status = 1; }
    else {
        // accumulator row m = 32 * warp + lane lives in TMEM lane m; a warp may only read its own 32 lanes
        for (int c0 = 0; c0 < p.n; c0 += 8) {
            unsigned v[8];
            const unsigned addr = taddr + ((unsigned)(warp * 32) << 16) + (unsigned)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr));
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            for (int j = 0; j < 8; ++j) p.d[(size_t)(warp * 32 + lane) * p.n + c0 + j] = __uint_as_float(v[j]);
        }
        if (tid == 0) *p.status = 0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (warp == 0) {
        unsigned cols = 32; while ((int)cols < p.n) cols <<= 1;
        if (cols == 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;\n" :: "r"(taddr));
        else if (cols == 64) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;\n" :: "r"(taddr));
        else if (cols == 128) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;\n" :: "r"(taddr));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" :: "r"(taddr));
    }
}

// C entry for ctypes: runs one variant; all pointers are device pointers
extern "C" int tcgen05_probe_run(const float* a, const float* b, float* d, int* status, int n, int k, unsigned lbo_a, unsigned sbo_a,
                                 unsigned lbo_b, unsigned sbo_b, unsigned layout_type, unsigned a_major, unsigned b_major, int swizzle_fill, int diag) {
    if (n % 32 || n < 32 || n > PROBE_MAXN || k % 8 || k < 8 || k > PROBE_MAXK) return -4;
    ProbeArgs p{a, b, d, status, n, k, lbo_a, sbo_a, lbo_b, sbo_b, layout_type, a_major, b_major, swizzle_fill, diag};
    const unsigned a_bytes = (PROBE_M / 32) * (lbo_a > 1024u ? lbo_a : 1024u) + (k / 8) * (sbo_a > 1024u ? sbo_a : 1024u);
    const unsigned b_bytes = (n / 32) * (lbo_b > 1024u ? lbo_b : 1024u) + (k / 8) * (sbo_b > 1024u ? sbo_b : 1024u);
    const size_t smem = ((a_bytes + 1023u) & ~1023u) + b_bytes + 1024;
    if (smem > 200 * 1024) return -4;
    if (cudaFuncSetAttribute(k_tcgen05_tf32_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -6;
    k_tcgen05_tf32_probe<<<1, 128, smem>>>(p);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -6;
}
