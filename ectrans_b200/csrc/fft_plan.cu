// Host-side FFT plan construction (see fft_plan.h).
#include "fft_plan.h"
#include <cmath>
#include <cstdio>
#include <algorithm>

int g_ect_force_bluestein = 0;   // test knob: route every length through the chirp-z path
static const int kPrimes[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31};

bool ect_fft_factorize(int n, std::vector<int>& radices) {
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
    // innermost first: odd radices (descending), then a single 2 if needed, then 4s outermost
    std::sort(odd.begin(), odd.end(), [](int a, int b) { return a > b; });
    radices = odd;
    if (twos % 2) radices.push_back(2);
    for (int i = 0; i < twos / 2; ++i) radices.push_back(4);
    return (int)radices.size() <= ECT_MAX_STAGES;
}

int ect_fft_smooth_size(int need) {
    for (int m = (need + 3) / 4 * 4;; m += 4) {
        int r = m;
        for (int p : {2, 3, 5, 7}) while (r % p == 0) r /= p;
        if (r == 1) return m;
    }
}

static double2 expi2pi(long long num, long long den) {   // exp(2 pi i num/den), exact argument reduction
    num %= den;
    if (num < 0) num += den;
    // reduce to first octant for accuracy
    long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)num / (long double)den;
    return make_double2((double)cosl(a), (double)sinl(a));
}

int EctFftTables::get_plan(int n) {
    auto it = plan_of_len.find(n);
    if (it != plan_of_len.end()) return it->second;
    std::vector<int> rad;
    if (!ect_fft_factorize(n, rad)) return -1;
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
    plan_of_len[n] = id;
    return id;
}

void ect_fft_host(const EctFftTables& T, int plan, std::vector<double2>& data) {
    const EctFftPlan& p = T.plans[plan];
    std::vector<double2> tmp(p.n);
    for (int i = 0; i < p.n; ++i) tmp[T.perm_pool[p.perm_off + i]] = data[i];
    for (int s = 0; s < p.nst; ++s)
        fft_stage<false>(tmp.data(), p.n, p.radix[s], p.sublen[s], T.tw_pool.data() + p.tw_off,
                         T.roots.data(), 0, 1);
    data = tmp;
}

int EctFftTables::get_latplan(int nlon, int km) {
    auto key = std::make_pair(nlon, km);
    auto it = latplan_of.find(key);
    if (it != latplan_of.end()) return it->second;
    EctLatPlan lp{};
    lp.nlon = nlon;
    lp.km = km;
    int direct = g_ect_force_bluestein ? -1 : get_plan(nlon);
    if (direct >= 0) {
        lp.plan = direct;
        lp.bluestein = 0;
        lp.m = 0;
        lp.chirp_off = lp.bhat_inv_off = lp.bhat_dir_off = -1;
        lp.smem_bytes = (nlon + plans[direct].tw_len) * (int)sizeof(double2);
    } else {
        const int N = nlon;
        const int ni_inv = 2 * km + 1, no_inv = N;       // inverse: inputs n in [-km, km], outputs k in [0, N)
        const int M = ect_fft_smooth_size(ni_inv + no_inv - 1);
        lp.bluestein = 1;
        lp.m = M;
        lp.plan = get_plan(M);
        lp.smem_bytes = (M + plans[lp.plan].tw_len) * (int)sizeof(double2);
        // chirp c[j] = exp(+i pi j^2 / N) = exp(2 pi i (j^2 mod 2N) / (2N)), j = 0 .. N/2
        lp.chirp_off = (int)cz_pool.size();
        for (int j = 0; j <= N / 2; ++j) {
            long long jj = ((long long)j * j) % (2LL * N);
            cz_pool.push_back(expi2pi(jj, 2LL * N));
        }
        auto chirp = [&](long long d) -> double2 {   // c[d] for any integer d (even N: period N, even symmetry)
            long long j = std::llabs(d) % N;
            if (j > N / 2) j = N - j;
            return cz_pool[lp.chirp_off + (int)j];
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
            for (int n = 0; n < M; ++n) {
                int pos = perm_pool[pm.perm_off + n];
                cz_pool[off + pos] = make_double2(h[n].x * inv, h[n].y * inv);
            }
            if (dir == 0) lp.bhat_inv_off = off; else lp.bhat_dir_off = off;
        }
    }
    latplans.push_back(lp);
    int id = (int)latplans.size() - 1;
    latplan_of[key] = id;
    return id;
}
