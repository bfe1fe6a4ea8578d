// Host-side FFT plan construction (see fft_plan.h).
#include "fft_plan.h"
#include <cmath>
#include <cstdio>
#include <algorithm>
#include <cstdlib>

int g_ect_force_bluestein = 0;   // test knob: route every length through the chirp-z path
static const int kPrimes[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31};

bool ect_fft_factorize(int n, std::vector<int>& radices, bool pow2_inner) {
    radices.clear();
    if (n < 2 || n % 2 != 0) return false;
    int rem = n, twos = 0;
    while (rem % 2 == 0) { rem /= 2; ++twos; }
    std::vector<int> odd;
    for (int p : kPrimes) {
        if (p == 2) continue;
        while (rem % p == 0) { odd.push_back(p); rem /= p; }
    }
    if (rem != 1) return false;
    std::sort(odd.begin(), odd.end(), [](int a, int b) { return a > b; });
    std::vector<int> p2;                      // power-of-two part as 16s (or 8s) plus one smaller radix
    static const char* r8 = getenv("ECT_FFT_R8");
    const int lg = (pow2_inner && r8 && atoi(r8)) ? 3 : 4;
    for (int i = 0; i < twos / lg; ++i) p2.push_back(1 << lg);
    if (twos % lg) p2.push_back(1 << (twos % lg));
    if (pow2_inner) {
        // chirp-z lengths r * 2^k: 16s innermost (fused middle step), small power of two next, odd r outermost
        radices = p2;
        radices.insert(radices.end(), odd.begin(), odd.end());
    } else {
        // direct lengths: odd radices innermost (descending; odd strides spread over the banks), powers of two outside
        radices = odd;
        std::reverse(p2.begin(), p2.end());
        radices.insert(radices.end(), p2.begin(), p2.end());
    }
    return (int)radices.size() <= ECT_MAX_STAGES;
}

// Half plans of the split chirp-z (H = r 2^k, r in {1, 3, 5, 7}): 16s innermost (the fused middle step wants a power
// of two), then what is left of the power of two merged with r into ONE register-resident composite radix where it
// fits (6, 10, 12, 14), so that H = 2560, 3072, 3584, 4096 are all three-stage plans 16 x 16 x {10, 12, 14, 16}.
bool ect_fft_factorize_half(int n, std::vector<int>& radices) {
    radices.clear();
    if (n < 2 || n % 2 != 0) return false;
    int rem = n, twos = 0;
    while (rem % 2 == 0) { rem /= 2; ++twos; }
    const int r = rem;
    if (r != 1 && r != 3 && r != 5 && r != 7) return ect_fft_factorize(n, radices, true);
    const int a = twos / 4, j = twos % 4;
    if (a == 0) return ect_fft_factorize(n, radices, true);
    for (int i = 0; i < a; ++i) radices.push_back(16);
    if (r == 1) { if (j) radices.push_back(1 << j); }
    else if (j == 0) radices.push_back(r);
    else if ((r << j) <= 14) radices.push_back(r << j);
    else if (r == 3) { radices.push_back(2); radices.push_back(12); }            // 24 = 2 x 12
    else { radices.push_back(1 << (j - 1)); radices.push_back(2 * r); }          // 20, 40 = {2, 4} x 10; 28, 56 = {2, 4} x 14
    return (int)radices.size() <= ECT_MAX_STAGES;
}

int ect_fft_smooth_size(int need) {          // smallest r * 2^k >= need, r in {1, 3, 5, 7}, k >= 4
    int best = 0;
    for (int r : {1, 3, 5, 7}) {
        int m = r * 16;
        while (m < need) m *= 2;
        if (best == 0 || m < best) best = m;
    }
    return best;
}

static double2 expi2pi(long long num, long long den) {   // exp(2 pi i num/den), exact argument reduction
    num %= den;
    if (num < 0) num += den;
    // reduce to first octant for accuracy
    long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)num / (long double)den;
    return make_double2((double)cosl(a), (double)sinl(a));
}

int EctFftTables::get_plan(int n, bool pow2_inner, bool half) {
    const int key = half ? n + (1 << 24) : (pow2_inner ? -n : n);
    auto it = plan_of_len.find(key);
    if (it != plan_of_len.end()) return it->second;
    std::vector<int> rad;
    if (!(half ? ect_fft_factorize_half(n, rad) : ect_fft_factorize(n, rad, pow2_inner))) return -1;
    if (roots.empty()) {
        roots.assign(ECT_ROOTS_SIZE, make_double2(0.0, 0.0));
        for (int p : kPrimes) {
            if (p == 2) continue;
            for (int j = 0; j < p; ++j) roots[ECT_ROOTS_OFF(p) + j] = expi2pi(j, p);
        }
    }
    EctFftPlan p{};
    p.n = n;
    p.nst = (int)rad.size();
    int L = 1;
    for (int s = 0; s < p.nst; ++s) {
        p.radix[s] = rad[s];
        p.sublen[s] = L;
        p.lshift[s] = -1;
        if ((L & (L - 1)) == 0) { int sh = 0; while ((1 << sh) < L) ++sh; p.lshift[s] = sh; }
        L *= rad[s];
    }
    p.perm_off = (int)perm_pool.size();
    for (int i = 0; i < n; ++i) {
        int rem = i, pos = 0;
        for (int s = p.nst - 1; s >= 0; --s) {
            int q = rem % p.radix[s];
            rem /= p.radix[s];
            pos += q * p.sublen[s];
        }
        perm_pool.push_back((uint16_t)pos);
    }
    p.tw_off = (int)tw_pool.size();
    p.quarter = (n % 4 == 0);
    p.tw_len = p.quarter ? n / 4 + 1 : n / 2;
    for (int j = 0; j < p.tw_len; ++j) tw_pool.push_back(expi2pi(j, n));
    plans.push_back(p);
    int id = (int)plans.size() - 1;
    plan_of_len[key] = id;
    return id;
}

void ect_fft_host(const EctFftTables& T, int plan, std::vector<double2>& data) {
    const EctFftPlan& p = T.plans[plan];
    std::vector<double2> tmp(ECT_PADDED_LEN(p.n));
    for (int i = 0; i < p.n; ++i) tmp[ECT_PAD((int)T.perm_pool[p.perm_off + i])] = data[i];
    std::vector<double2> t1(ECT_TW1_LEN(p.n)), t2(ECT_TW2_LEN);
    tw_build(t1.data(), t2.data(), T.tw_pool.data() + p.tw_off, p.n, 0, 1);
    EctTw tw{t1.data(), t2.data()};
    for (int s = 0; s < p.nst; ++s)
        fft_stage<false>(tmp.data(), p.n, p.radix[s], p.sublen[s], p.lshift[s], tw, T.roots.data(), 0, 1);
    for (int i = 0; i < p.n; ++i) data[i] = tmp[ECT_PAD(i)];
}

int EctFftTables::get_latplan(int nlon, int km) {
    auto key = std::make_pair(nlon, km);
    auto it = latplan_of.find(key);
    if (it != latplan_of.end()) return it->second;
    EctLatPlan lp{};
    lp.nlon = nlon;
    lp.km = km;
    int direct = g_ect_force_bluestein ? -1 : get_plan(nlon, false);
    if (direct >= 0) {
        lp.plan = direct;
        lp.bluestein = 0;
        lp.m = 0;
        lp.chirp_off = lp.bhat_inv_off = lp.bhat_dir_off = lp.ctw_off = -1;
        lp.plan_h = -1; lp.bhat_inv_eo[0] = lp.bhat_inv_eo[1] = lp.bhat_dir_eo[0] = lp.bhat_dir_eo[1] = -1;
        lp.smem_bytes = ECT_PADDED_LEN(nlon) * (int)sizeof(double2);
    } else {
        const int N = nlon;
        const int ni_inv = 2 * km + 1, no_inv = N;       // inverse: inputs n in [-km, km], outputs k in [0, N)
        const int M = ect_fft_smooth_size(ni_inv + no_inv - 1);
        lp.bluestein = 1;
        lp.m = M;
        lp.plan = get_plan(M, true);
        lp.plan_h = get_plan(M / 2, true, true);      // before any reference into plans[] is taken
        lp.smem_bytes = ECT_PADDED_LEN(M) * (int)sizeof(double2);
        // chirp c[j] = exp(+i pi j^2 / N) = exp(2 pi i (j^2 mod 2N) / (2N)), j = 0 .. N/2
        lp.chirp_off = (int)cz_pool.size();
        for (int j = 0; j <= N / 2; ++j) {
            long long jj = ((long long)j * j) % (2LL * N);
            cz_pool.push_back(expi2pi(jj, 2LL * N));
        }
        lp.ctw_off = (int)cz_pool.size();
        for (int a = 0; a < ECT_TW1_LEN(2 * N); ++a) cz_pool.push_back(expi2pi((128LL * a) % (2LL * N), 2LL * N));
        for (int b = 0; b < ECT_TW2_LEN; ++b) cz_pool.push_back(expi2pi(b % (2LL * N), 2LL * N));
        // c[d] for any integer d from the stored c[0 .. N/2]: c is even, c[j + N] = (-1)^N c[j] and hence
        // c[N - j] = (-1)^N c[j] -- for odd row lengths (classic reduced grids) the reflections flip the sign
        auto chirp = [&](long long d) -> double2 {
            long long j = std::llabs(d) % (2LL * N);
            double sg = 1.0;
            if (j >= N) { j -= N; if (N & 1) sg = -sg; }
            if (j > N / 2) { j = N - j; if (N & 1) sg = -sg; }
            const double2 c = cz_pool[lp.chirp_off + (int)j];
            return make_double2(sg * c.x, sg * c.y);
        };
        const EctFftPlan& pm = plans[lp.plan];
        for (int dir = 0; dir < 2; ++dir) {
            // kernel h[e mod M] = conj(c[e + o0 - i0]),  e in [-(ni-1), no-1]
            const int i0 = dir == 0 ? -km : 0;
            const int ni = dir == 0 ? 2 * km + 1 : N;
            const int o0 = dir == 0 ? 0 : -km;
            const int no = dir == 0 ? N : 2 * km + 1;
            std::vector<double2> h(M, make_double2(0.0, 0.0));
            for (int e = -(ni - 1); e <= no - 1; ++e) {
                double2 c = chirp((long long)e + o0 - i0);
                h[((e % M) + M) % M] = make_double2(c.x, -c.y);
            }
            // forward (sign -) FFT via swap trick on the sign-+ core
            for (auto& v : h) std::swap(v.x, v.y);
            ect_fft_host(*this, lp.plan, h);
            for (auto& v : h) std::swap(v.x, v.y);
            int off = (int)cz_pool.size();
            cz_pool.resize(off + M);
            const double inv = 1.0 / (double)M;
            const int r0 = pm.radix[0], nb0 = M / r0;       // middle step reads bhat[q * nb0 + b] for position b * r0 + q
            for (int n = 0; n < M; ++n) {
                const int pos = perm_pool[pm.perm_off + n];
                cz_pool[off + (pos % r0) * nb0 + pos / r0] = make_double2(h[n].x * inv, h[n].y * inv);
            }
            if (dir == 0) lp.bhat_inv_off = off; else lp.bhat_dir_off = off;
            // split form: bins 2 r (even) and 2 r + 1 (odd) of the same spectrum, each permuted for the length-H plan
            const EctFftPlan& ph = plans[lp.plan_h];
            const int H = M / 2, rh = ph.radix[0], nbh = H / rh;
            for (int par = 0; par < 2; ++par) {
                const int offh = (int)cz_pool.size();
                cz_pool.resize(offh + H);
                for (int rr = 0; rr < H; ++rr) {
                    const int pos = perm_pool[ph.perm_off + rr];
                    cz_pool[offh + (pos % rh) * nbh + pos / rh] = make_double2(h[2 * rr + par].x * inv, h[2 * rr + par].y * inv);
                }
                if (dir == 0) lp.bhat_inv_eo[par] = offh; else lp.bhat_dir_eo[par] = offh;
            }
        }
    }
    latplans.push_back(lp);
    int id = (int)latplans.size() - 1;
    latplan_of[key] = id;
    return id;
}
