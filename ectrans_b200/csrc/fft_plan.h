// Host-side FFT plan construction for the Fourier stage: factorisation, digit-reversal
// permutation, quarter-wave twiddle tables and Bluestein (chirp-z) tables for lengths
// with a prime factor > ECT_MAX_RADIX.  Replaces the reference's FFTW plan cache
// (cpu/internal/tpm_fftw.F90:84-213) and cuFFT plan cache (gpu/algor/hicfft.cuda.cu:136-163).
#pragma once
#include <vector>
#include <map>
#include "fft_core.h"

struct EctLatPlan {        // one per distinct (nlon, nmen)
    int nlon, km;
    int plan;              // index into plans[]: the length-nlon plan (direct) or the length-M plan (Bluestein)
    int bluestein;         // 0 / 1
    int m;                 // Bluestein convolution length (0 if direct)
    int chirp_off;         // double2 pool: c[j] = exp(+i pi j^2 / nlon), j = 0 .. nlon/2
    int ctw_off;           // double2 pool: two-level table of exp(2 pi i t / (2 nlon)): [ECT_TW1_LEN(2 nlon)] then [128]
    int bhat_inv_off;      // double2 pool: permuted spectrum of the inverse-direction kernel (includes 1/M)
    int bhat_dir_off;      // same for the direct direction
    int smem_bytes;        // dynamic shared memory for one field pair + twiddle table
    // chirp-z split over a CTA pair: the length-M convolution as two independent length-H (H = M/2) ones, over the
    // even bins (e[u] = x[u] + x[u+H]) and the odd bins (o[u] = (x[u] - x[u+H]) exp(-2 pi i u / M)),
    // y[j] = E[j] + w^j O[j], y[j+H] = E[j] - w^j O[j], w = exp(2 pi i / M)
    int plan_h;            // index into plans[]: the length-H plan (composite radices allowed)
    int bhat_inv_eo[2];    // double2 pool: kernel spectrum of the even / odd bins, permuted for the length-H middle step
    int bhat_dir_eo[2];
};

struct EctFftTables {
    std::vector<EctFftPlan> plans;
    std::vector<uint16_t> perm_pool;
    std::vector<double2> tw_pool;
    std::vector<double2> cz_pool;       // chirps and Bluestein kernels
    std::vector<double2> roots;         // [ECT_ROOTS_SIZE]: odd radix R at ECT_ROOTS_OFF(R)
    std::vector<EctLatPlan> latplans;
    std::map<int, int> plan_of_len;
    std::map<std::pair<int, int>, int> latplan_of;
    int get_plan(int n, bool pow2_inner, bool half = false);   // smooth even n; pow2_inner: chirp-z stage order; half: length-H plan of the split chirp-z
    int get_latplan(int nlon, int km);
};

bool ect_fft_factorize(int n, std::vector<int>& radices, bool pow2_inner);   // false if a prime factor > ECT_MAX_RADIX
bool ect_fft_factorize_half(int n, std::vector<int>& radices);               // chirp-z half plans: 16s innermost, one composite radix (6 / 10 / 12 / 14) outermost
int ect_fft_smooth_size(int need);                           // smallest r * 2^k >= need, r in {1,3,5,7}
// reference host FFT (uses the same core single-threaded); sign +, unnormalised, natural order in/out
void ect_fft_host(const EctFftTables& T, int plan, std::vector<double2>& data);
