// Test-only CPU harness: runs the __host__ __device__ Fourier-stage phases of
// ectrans_b200/csrc single-threaded so that index logic is checked without a GPU.
// Built by tests/conftest.py into tests/hostemu/_build/; never loaded by the product.
#include "../../ectrans_b200/csrc/fourier_phases.h"
#include "../../ectrans_b200/csrc/fourier_cz.h"
#include <vector>
#include <cstring>
extern int g_ect_force_bluestein;

static EctFftPlan g_plan;
static void run_stages(double2* data, const EctPairCtx& c, bool dif, int nthr) {
    const EctFftPlan& p = g_plan;
    if (dif) {       // chirp-z: DIF stages, fused middle, DIT stages
        for (int s = p.nst - 1; s >= 1; --s)
            for (int t = 0; t < nthr; ++t) fft_stage<true>(data, p.n, p.radix[s], p.sublen[s], p.lshift[s], c.qt, c.roots, t, nthr);
        for (int t = 0; t < nthr; ++t) blue_middle(data, p.n, p.radix[0], c.bhat, t, nthr);
        for (int s = 1; s < p.nst; ++s)
            for (int t = 0; t < nthr; ++t) fft_stage<false>(data, p.n, p.radix[s], p.sublen[s], p.lshift[s], c.qt, c.roots, t, nthr);
    } else {
        for (int s = 0; s < p.nst; ++s)
            for (int t = 0; t < nthr; ++t) fft_stage<false>(data, p.n, p.radix[s], p.sublen[s], p.lshift[s], c.qt, c.roots, t, nthr);
    }
}

static std::vector<double2> g_t1, g_t2;
static void make_ctx(EctFftTables& T, EctPairCtx& c, int nlon, int km, int dir, std::vector<int>& rec) {
    int id = T.get_latplan(nlon, km);
    const EctLatPlan& lp = T.latplans[id];
    c.nlon = nlon; c.km = km; c.racthe = 1.0;
    g_plan = T.plans[lp.plan];
    c.perm = T.perm_pool.data() + g_plan.perm_off;
    g_t1.assign(ECT_TW1_LEN(g_plan.n), make_double2(0, 0)); g_t2.assign(ECT_TW2_LEN, make_double2(0, 0));
    tw_build(g_t1.data(), g_t2.data(), T.tw_pool.data() + g_plan.tw_off, g_plan.n, 0, 1);
    c.qt = EctTw{g_t1.data(), g_t2.data()};
    c.roots = T.roots.data();
    c.bluestein = lp.bluestein; c.m = lp.m;
    c.chirp = lp.bluestein ? T.cz_pool.data() + lp.chirp_off : nullptr;
    c.bhat = lp.bluestein ? T.cz_pool.data() + (dir == 0 ? lp.bhat_inv_off : lp.bhat_dir_off) : nullptr;
    rec.resize(km + 1);
    for (int k = 0; k <= km; ++k) rec[k] = k;
    c.rec = rec.data(); c.cp = 4;
}

extern "C" {
int emu_smooth(int n) { std::vector<int> r; return ect_fft_factorize(n, r, false) ? 1 : 0; }

// complex sign-+ FFT of smooth length n
int emu_fft(int n, const double* in, double* out) {
    EctFftTables T;
    int id = T.get_plan(n, false);
    if (id < 0) return -1;
    std::vector<double2> d(n);
    for (int i = 0; i < n; ++i) d[i] = make_double2(in[2 * i], in[2 * i + 1]);
    ect_fft_host(T, id, d);
    for (int i = 0; i < n; ++i) { out[2 * i] = d[i].x; out[2 * i + 1] = d[i].y; }
    return 0;
}

// float instantiation of the same core (sp handles): complex sign-+ FFT of smooth length n, and with
// chirpz = 1 the DIF / fused middle / DIT chain on length n against an all-ones kernel spectrum
// (= forward FFT on swapped data then inverse, i.e. n * swap(x) reproduced), exercising blue_middle<float2>
int emu_fft_float(int n, const double* in, double* out, int chirpz) {
    EctFftTables T;
    int id = T.get_plan(n, chirpz != 0);
    if (id < 0) return -1;
    const EctFftPlan& p = T.plans[id];
    std::vector<float2> t1(ECT_TW1_LEN(n)), t2(ECT_TW2_LEN), roots(T.roots.size()), data(ECT_PADDED_LEN(n));
    tw_build(t1.data(), t2.data(), T.tw_pool.data() + p.tw_off, n, 0, 1);
    for (size_t i = 0; i < roots.size(); ++i) roots[i] = c_cvt<float2>(T.roots[i]);
    const EctTwT<float2> tw{t1.data(), t2.data()};
    const uint16_t* perm = T.perm_pool.data() + p.perm_off;
    const int nthr = 3;
    if (!chirpz) {
        for (int i = 0; i < n; ++i) data[ECT_PAD((int)perm[i])] = make_float2((float)in[2 * i], (float)in[2 * i + 1]);
        for (int s = 0; s < p.nst; ++s)
            for (int t = 0; t < nthr; ++t) fft_stage<false>(data.data(), n, p.radix[s], p.sublen[s], p.lshift[s], tw, (const float2*)roots.data(), t, nthr);
    } else {
        if (p.radix[0] & 1) return -2;
        std::vector<float2> ones(n, make_float2(1.f, 0.f));
        for (int i = 0; i < n; ++i) data[ECT_PAD(i)] = make_float2((float)in[2 * i], (float)in[2 * i + 1]);
        for (int s = p.nst - 1; s >= 1; --s)
            for (int t = 0; t < nthr; ++t) fft_stage<true>(data.data(), n, p.radix[s], p.sublen[s], p.lshift[s], tw, (const float2*)roots.data(), t, nthr);
        for (int t = 0; t < nthr; ++t) blue_middle(data.data(), n, p.radix[0], (const float2*)ones.data(), t, nthr);
        for (int s = 1; s < p.nst; ++s)
            for (int t = 0; t < nthr; ++t) fft_stage<false>(data.data(), n, p.radix[s], p.sublen[s], p.lshift[s], tw, (const float2*)roots.data(), t, nthr);
    }
    for (int i = 0; i < n; ++i) { out[2 * i] = data[ECT_PAD(i)].x; out[2 * i + 1] = data[ECT_PAD(i)].y; }
    return 0;
}

// inverse pair: spec [km+1][4] = (reA, imA, reB, imB) records; out rows [nlon] each
int emu_ftinv_pair(int nlon, int km, const double* spec, double* outa, double* outb, int nthr, int force_blue) {
    g_ect_force_bluestein = force_blue;
    EctFftTables T; EctPairCtx c; std::vector<int> rec;
    make_ctx(T, c, nlon, km, 0, rec);
    std::vector<double2> data(ECT_PADDED_LEN(c.bluestein ? c.m : nlon));
    EctFsField fa{0, 0, 0}, fb{2, 0, 0};
    for (int t = 0; t < nthr; ++t) ftinv_load(data.data(), spec, c, fa, fb, t, nthr);
    run_stages(data.data(), c, c.bluestein != 0, nthr);
    for (int j = 0; j < nlon; ++j) { double2 x = ftinv_out(data.data(), c, j); outa[j] = x.x; outb[j] = x.y; }
    g_ect_force_bluestein = 0;
    return c.bluestein;
}

// direct pair: rows -> spec records [km+1][4]
int emu_ftdir_pair(int nlon, int km, const double* rowa, const double* rowb, double* spec, int nthr, int force_blue) {
    g_ect_force_bluestein = force_blue;
    EctFftTables T; EctPairCtx c; std::vector<int> rec;
    make_ctx(T, c, nlon, km, 1, rec);
    std::vector<double2> data(ECT_PADDED_LEN(c.bluestein ? c.m : nlon));
    for (int j = 0; j < nlon; ++j) ftdir_put(data.data(), c, j, rowa[j], rowb[j]);
    for (int t = 0; t < nthr; ++t) ftdir_zero_tail(data.data(), c, t, nthr);
    run_stages(data.data(), c, c.bluestein != 0, nthr);
    for (int t = 0; t < nthr; ++t) ftdir_store(data.data(), spec, c, 0, 2, t, nthr);
    g_ect_force_bluestein = 0;
    return c.bluestein;
}
}

// ---- chirp-z rows split over a CTA pair (fourier_cz.h): both halves run one after the other on the CPU ----
struct CzEmu {
    EctFftTables T; EctLatPlan lp; EctFftPlan ph;
    std::vector<double2> t1, t2, t1c, t2c, dat[2];
    CzCtx<double2> cx[2]; EctTw qth;
    int setup(int nlon, int km) {
        g_ect_force_bluestein = 1;
        const int id = T.get_latplan(nlon, km);
        g_ect_force_bluestein = 0;
        lp = T.latplans[id];
        if (!lp.bluestein || lp.plan_h < 0) return -1;
        ph = T.plans[lp.plan_h];
        const int H = ph.n, M = 2 * H;
        if (M != lp.m) return -2;
        t1.assign(ECT_TW1_LEN(M), make_double2(0, 0)); t2.assign(ECT_TW2_LEN, make_double2(0, 0));
        tw_build(t1.data(), t2.data(), T.tw_pool.data() + T.plans[lp.plan].tw_off, M, 0, 1);
        qth.t1 = t1.data(); qth.t2 = t2.data(); qth.sh = 1;
        t1c.assign(T.cz_pool.begin() + lp.ctw_off, T.cz_pool.begin() + lp.ctw_off + ECT_TW1_LEN(2 * nlon));
        t2c.assign(T.cz_pool.begin() + lp.ctw_off + ECT_TW1_LEN(2 * nlon), T.cz_pool.begin() + lp.ctw_off + ECT_TW1_LEN(2 * nlon) + ECT_TW2_LEN);
        for (int h = 0; h < 2; ++h) {
            dat[h].assign(ECT_PADDED_LEN(H), make_double2(1e300, 1e300));      // poison: every element must be written
            cx[h].N = nlon; cx[h].km = km; cx[h].H = H; cx[h].half = h;
            cx[h].twm.t1 = t1.data(); cx[h].twm.t2 = t2.data(); cx[h].twm.sh = 0;
            cx[h].twc.t1 = t1c.data(); cx[h].twc.t2 = t2c.data(); cx[h].twc.sh = 0;
            cx[h].n2 = 2u * (unsigned)nlon; cx[h].magic = (unsigned)((0x100000000ull + cx[h].n2 - 1) / cx[h].n2);
        }
        return 0;
    }
    void passes(int h, int dir, int nthr) {
        double2* data = dat[h].data();
        const double2* bhat = T.cz_pool.data() + (dir == 0 ? lp.bhat_inv_eo[h] : lp.bhat_dir_eo[h]);
        const int H = ph.n;
        for (int s = ph.nst - 1; s >= 1; --s)
            for (int t = 0; t < nthr; ++t) fft_stage<true, 7, true>(data, H, ph.radix[s], ph.sublen[s], ph.lshift[s], qth, (const double2*)T.roots.data(), t, nthr);
        for (int t = 0; t < nthr; ++t) blue_middle(data, H, ph.radix[0], bhat, t, nthr);
        for (int s = 1; s < ph.nst; ++s)
            for (int t = 0; t < nthr; ++t) fft_stage<false, 7, true>(data, H, ph.radix[s], ph.sublen[s], ph.lshift[s], qth, (const double2*)T.roots.data(), t, nthr);
    }
};

extern "C" {
// radices of the half plan of length n (returns the number of stages)
int emu_half_plan(int n, int* radices) {
    std::vector<int> r;
    if (!ect_fft_factorize_half(n, r)) return -1;
    for (size_t i = 0; i < r.size(); ++i) radices[i] = r[i];
    return (int)r.size();
}

// complex sign-+ FFT of length n on a half plan (composite radices), DIT from the permuted input
int emu_fft_half(int n, const double* in, double* out) {
    EctFftTables T;
    int id = T.get_plan(n, true, true);
    if (id < 0) return -1;
    std::vector<double2> d(n);
    for (int i = 0; i < n; ++i) d[i] = make_double2(in[2 * i], in[2 * i + 1]);
    ect_fft_host(T, id, d);
    for (int i = 0; i < n; ++i) { out[2 * i] = d[i].x; out[2 * i + 1] = d[i].y; }
    return 0;
}

int emu_cz_inv_pair(int nlon, int km, const double* spec, double* outa, double* outb, int nthr) {
    CzEmu E;
    int rc = E.setup(nlon, km);
    if (rc) return rc;
    CzInvScale sc; sc.pwa = sc.pwb = 0; sc.deriva = sc.derivb = 0; sc.hasb = 1; sc.s1 = sc.s2 = 1.0; sc.rowscale = 1.0;
    for (int h = 0; h < 2; ++h) {
        for (int t = 0; t < nthr; ++t)
            cz_inv_load(E.dat[h].data(), E.cx[h], sc, [&](int k, double2& ra, double2& rb) {
                ra = make_double2(spec[4 * k], spec[4 * k + 1]); rb = make_double2(spec[4 * k + 2], spec[4 * k + 3]);
            }, t, nthr);
        E.passes(h, 0, nthr);
    }
    std::vector<int> cnt(nlon, 0);
    for (int h = 0; h < 2; ++h)
        for (int t = 0; t < nthr; ++t)
            cz_inv_out((const double2*)E.dat[h].data(), (const double2*)E.dat[h ^ 1].data(), E.cx[h],
                       [&](int j, double2 y) { outa[j] = y.x; outb[j] = y.y; cnt[j]++; }, t, nthr);
    for (int j = 0; j < nlon; ++j) if (cnt[j] != 1) return -10;      // every longitude written exactly once
    return E.ph.n;
}

int emu_cz_dir_pair(int nlon, int km, const double* rowa, const double* rowb, double* spec, int nthr) {
    CzEmu E;
    int rc = E.setup(nlon, km);
    if (rc) return rc;
    for (int h = 0; h < 2; ++h) {
        for (int t = 0; t < nthr; ++t)
            cz_dir_load(E.dat[h].data(), E.cx[h], [&](int j, double& va, double& vb) { va = rowa[j]; vb = rowb[j]; }, t, nthr);
        E.passes(h, 1, nthr);
    }
    const double sc = 0.5 / (double)nlon;
    std::vector<int> cnt(km + 1, 0);
    for (int h = 0; h < 2; ++h)
        for (int t = 0; t < nthr; ++t)
            cz_dir_out((const double2*)E.dat[h].data(), (const double2*)E.dat[h ^ 1].data(), E.cx[h],
                       [&](int k, double2 Zk, double2 Zn) {
                           spec[4 * k] = (Zk.x + Zn.x) * sc; spec[4 * k + 1] = (Zk.y - Zn.y) * sc;
                           spec[4 * k + 2] = (Zk.y + Zn.y) * sc; spec[4 * k + 3] = (Zn.x - Zk.x) * sc;
                           cnt[k]++;
                       }, t, nthr);
    for (int k = 0; k <= km; ++k) if (cnt[k] != 1) return -10;
    return E.ph.n;
}
}

#include "../../ectrans_b200/csrc/supolf.h"
extern "C" void emu_supolf(int km, int par, int kcount, int knsmax, int nlat, const double* mu, double* out) {
    EctSupolfM cm;
    ect_supolf_consts(km, cm);
    for (int i = 0; i < nlat; ++i) ect_supolf_column(km, par, kcount, knsmax, mu[i], cm, out + i, nlat);
}
