// Host plan: Gaussian grid, truncation geometry and the two-phase decomposition.
// Pure host C++ (no CUDA calls) so that it can be used and tested without a GPU.
#include "ect_internal.h"
#include "fft_plan.h"
#include <cmath>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <algorithm>

static thread_local char g_err[1024] = "";
void ect_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* ect_last_error(void) { return g_err; }

// Gaussian latitudes and weights.  Newton iteration on the cosine series of the normalised
// Legendre polynomial of degree ndgl: SUGAW (LLOLD branch) common/internal/sugaw_mod.F90:157-190,
// GAWL common/internal/gawl_mod.F90:91-111, CPLEDN common/internal/cpledn_mod.F90:94-129, series
// coefficients cpu/internal/suleg_mod.F90:249-263.
void ect_gauss_latitudes(int ndgl, std::vector<double>& mu, std::vector<double>& w) {
    const int kn = ndgl;
    std::vector<double> row(ndgl + 1, 0.0);
    double zfnn = 2.0;
    for (int j = 1; j <= ndgl; ++j) zfnn *= std::sqrt(1.0 - 0.25 / ((double)j * (double)j));
    row[ndgl] = zfnn;
    for (int j = 2; j <= ndgl; j += 2)
        row[ndgl - j] = row[ndgl - j + 2] * (double)((long long)(j - 1) * (2 * ndgl - j + 2)) /
                        (double)((long long)j * (2 * ndgl - j + 1));
    const int ins2 = ndgl / 2;
    std::vector<double> zfn(ins2 + 1);
    for (int i = 0; i <= ins2; ++i) zfn[i] = row[2 * i];
    mu.assign(ndgl, 0.0);
    w.assign(ndgl, 0.0);
    const double eps = 2.220446049250313e-16;
    const double pi = 2.0 * std::asin(1.0);
    for (int jgl = 1; jgl <= ins2; ++jgl) {
        const double z = (double)(4 * jgl - 1) * pi / (double)(4 * kn + 2);
        double zx = z + 1.0 / (std::tan(z) * (double)(8LL * kn * kn));
        int iflag = 0;
        double zw = 0.0;
        for (int it = 0; it < 21; ++it) {
            if (iflag == 0) {
                double zdlk = 0.5 * zfn[0], zdlldn = 0.0;
                for (int ik = 1; ik <= ins2; ++ik) {
                    const double jn = (double)(2 * ik);
                    zdlk += zfn[ik] * std::cos(jn * zx);
                    zdlldn -= zfn[ik] * jn * std::sin(jn * zx);
                }
                const double zmod = -zdlk / zdlldn;
                zx += zmod;
                if (std::fabs(zmod) <= eps * 1000.0) iflag = 1;
            } else {
                double zdlldn = 0.0;
                for (int ik = 1; ik <= ins2; ++ik) {
                    const double jn = (double)(2 * ik);
                    zdlldn -= zfn[ik] * jn * std::sin(jn * zx);
                }
                zw = (double)(2 * kn + 1) / (zdlldn * zdlldn);
                break;
            }
        }
        mu[jgl - 1] = std::cos(zx);
        w[jgl - 1] = zw;
    }
    for (int j = 0; j < ins2; ++j) {
        mu[ndgl - 1 - j] = -mu[j];
        w[ndgl - 1 - j] = w[j];
    }
}

// NMEN: common/internal/setup_geom_mod.F90:44-78
static void compute_nmen(EctHostPlan& P) {
    const int ndgl = P.ndgl, T = P.nsmax, ndgnh = P.ndgnh;
    P.nmen.assign(ndgl, 0);
    bool reduced = false;
    for (int j = 1; j < ndgl; ++j) reduced |= (P.nloen[j] != P.nloen[0]);
    const int lin = ndgl - 1;
    if (T >= lin || !reduced) {
        for (int j = 0; j < ndgl; ++j) P.nmen[j] = std::min(T, (P.nloen[j] - 1) / 2);
        return;
    }
    std::vector<double> zsq(ndgl);
    int sub;
    if (T >= ndgl * 2 / 3 - 1) {
        const double fac = (double)(3 * (lin - T) / ndgl);
        for (int j = 0; j < ndgl; ++j) zsq[j] = fac * P.r1mu2[j];
        sub = 0;
    } else {
        zsq = P.r1mu2;
        sub = 1;
    }
    auto val = [&](int j) { return (int)((double)(P.nloen[j] - 1) / (2.0 + zsq[j])) - sub; };
    P.nmen[0] = std::min(T, val(0));
    for (int j = 1; j < ndgnh; ++j) P.nmen[j] = std::min(T, std::max(P.nmen[j - 1], val(j)));
    P.nmen[ndgl - 1] = std::min(T, val(ndgl - 1));
    for (int j = ndgl - 2; j >= ndgnh; --j) P.nmen[j] = std::min(T, std::max(P.nmen[j + 1], val(j)));
}

// Latitude bands in Fourier space: SUMPLATB (LDSPLIT=.F.) common/internal/sumplatb_mod.F90:171-216,
// SUMPLATF common/internal/sumplatf_mod.F90:110-139
static void lat_bands(const std::vector<int>& nloen, int nproca, std::vector<int>& first, std::vector<int>& count) {
    const int ndgl = (int)nloen.size();
    long long imedia = 0;
    for (int v : nloen) imedia += v;
    long long kmediap = imedia / nproca;
    const long long krestm = imedia - kmediap * nproca;
    if (krestm > 0) kmediap += 1;
    std::vector<int> klast(nproca + 2, 0);
    long long itot_top = 0, itot_bot = 0;
    int igl_top = 1, igl_bot = ndgl;
    for (int ja = 1; ja <= (nproca - 1) / 2 + 1; ++ja) {
        if (ja != nproca / 2 + 1) {
            for (;;) {
                if (igl_top <= ndgl && itot_top + nloen[igl_top - 1] < kmediap) {
                    klast[ja] = igl_top;
                    itot_top += nloen[igl_top - 1];
                    ++igl_top;
                } else { itot_top -= kmediap; break; }
            }
            klast[nproca - ja + 1] = igl_bot;
            for (;;) {
                if (igl_bot >= 1 && itot_bot + nloen[igl_bot - 1] < kmediap) {
                    itot_bot += nloen[igl_bot - 1];
                    --igl_bot;
                } else { itot_bot -= kmediap; break; }
            }
        } else {
            klast[ja] = igl_bot;
        }
    }
    bool simple = false;
    for (int ja = 1; ja <= nproca; ++ja) simple |= (klast[ja] == 0);
    if (simple) {
        std::vector<int> ilats(nproca + 1, 0);
        int ia = 0;
        for (int j = 0; j < ndgl; ++j) { ++ia; ilats[ia]++; if (ia == nproca) ia = 0; }
        klast[1] = ilats[1];
        for (int ja = 2; ja <= nproca; ++ja) klast[ja] = klast[ja - 1] + ilats[ja];
    }
    first.assign(nproca, 0);
    count.assign(nproca, 0);
    int prev = 0;
    for (int ja = 1; ja <= nproca; ++ja) {
        int cnt = klast[ja] != 0 ? klast[ja] - prev : 0;
        if (cnt < 0) cnt = 0;
        first[ja - 1] = prev;
        count[ja - 1] = cnt;
        prev += cnt;
    }
}

int ect_build_host_plan(EctHostPlan& P, int nsmax, int ndgl, const int* nloen, int nranks, int rank, bool gp_eq, bool bands_by_points) {
    if (nsmax < 0 || ndgl < 2 || (ndgl & 1) || !nloen || nranks < 1 || rank < 0 || rank >= nranks) {
        ect_set_error("ect_setup: bad arguments (nsmax=%d ndgl=%d nranks=%d rank=%d)", nsmax, ndgl, nranks, rank);
        return ECT_ERR_BADARG;
    }
    P.nsmax = nsmax; P.ndgl = ndgl; P.ndgnh = (ndgl + 1) / 2;
    P.nranks = nranks; P.rank = rank;
    P.nloen.assign(nloen, nloen + ndgl);
    for (int j = 0; j < ndgl; ++j) {
        if (P.nloen[j] < 2) {
            ect_set_error("ect_setup: nloen(%d)=%d", j + 1, P.nloen[j]);
            return ECT_ERR_BADARG;
        }
        if (P.nloen[j] != P.nloen[ndgl - 1 - j]) {
            ect_set_error("ect_setup: grid not symmetric about the equator");
            return ECT_ERR_BADARG;
        }
    }
    ect_gauss_latitudes(ndgl, P.rmu, P.rw);
    P.r1mu2.resize(ndgl); P.racthe.resize(ndgl);
    for (int j = 0; j < ndgl; ++j) {     // cpu/internal/suleg_mod.F90:386-394
        const double zcos = std::cos(std::asin(P.rmu[j]));
        P.r1mu2[j] = zcos * zcos;
        P.racthe[j] = 1.0 / zcos / ECT_RA;
    }
    compute_nmen(P);
    P.ndglu.assign(nsmax + 1, 0);
    for (int m = 0; m <= nsmax; ++m) {
        int c = 0;
        for (int j = 0; j < P.ndgnh; ++j) c += (P.nmen[j] >= m);
        P.ndglu[m] = c;
    }
    // the Legendre stage assumes the latitudes carrying m are the ndglu(m) ones nearest the equator
    // (ISL = NDGNH - NDGLU + 1, cpu/internal/leinv_mod.F90:95-97)
    for (int j = 1; j < P.ndgnh; ++j)
        if (P.nmen[j] < P.nmen[j - 1] || P.nmen[j] != P.nmen[ndgl - 1 - j]) {
            ect_set_error("ect_setup: NMEN not monotone/symmetric; unsupported grid");
            return ECT_ERR_NOTIMPL;
        }
    // SUWAVEDI: common/internal/suwavedi_mod.F90:118-137
    P.nprocm.assign(nsmax + 1, 0);
    P.ms_of.assign(nranks, {});
    {
        int ind = 1, ik = 0;
        for (int jm = 0; jm <= nsmax; ++jm) {
            ik += ind;
            if (ik > nranks) { ik = nranks; ind = -1; }
            else if (ik < 1) { ik = 1; ind = 1; }
            P.nprocm[jm] = ik - 1;
            P.ms_of[ik - 1].push_back(jm);
        }
    }
    P.myms = P.ms_of[rank];
    P.nump = (int)P.myms.size();
    P.nasm0.assign(nsmax + 1, -1);
    {
        int pos = 0;
        for (int m : P.myms) { P.nasm0[m] = pos; pos += 2 * (nsmax - m + 1); }
        P.nspec2 = pos;
    }
    P.nspec2_g = (nsmax + 1) * (nsmax + 2);
    {
        // SUMPLATB balances the bands by grid points.  The Fourier kernels cost, per latitude: (i) the FFT work -- N log2 N
        // for a 31-smooth length, two transforms of the convolution length M for a chirp-z row --, (ii) the gather /
        // scatter of its NMEN + 1 records and (iii) a fixed price (CTAs set up per row, short rows that do not fill them).
        // A least-squares fit of the per-rank Fourier times (inverse + direct) of 8 B200s under three different
        // partitions (profiles/r02b_push_bench_n8.log, 24 band times) gives, in ms,
        //     3.98e-7 N log2 N  |  2.58e-7 * 2 M log2 M   +   1.78e-5 (NMEN + 1)   +   5.67e-3
        // to 1.2 %, and predicts the four band times of the 4-GPU run (not in the fit) to 0.5 % of each other, as
        // measured.  By default the same SUMPLATB algorithm runs on that weight (in ns); ECT_SETUP_BANDS_BY_POINTS
        // restores the reference's plain point count, ECT_BAND_PAD=x the one-constant model of the first half of round 2
        // (FFT work + x max(FFT work), x = 0.16: polar bands 5 % over-, the next ones 5 % under-estimated).
        std::vector<int> w(P.nloen);
        P.band_pad = 0;
        if (!bands_by_points && nranks > 1) {
            const char* pe = getenv("ECT_BAND_PAD");
            std::vector<double> fw(ndgl);
            std::vector<char> cz(ndgl, 0);
            double fmax = 0;
            for (int j = 0; j < ndgl; ++j) {
                const int n = P.nloen[j];
                std::vector<int> rad;
                double f;
                if (n % 2 == 0 && ect_fft_factorize(n, rad, false)) f = n * std::log2((double)std::max(n, 2));
                else { const int M = ect_fft_smooth_size(2 * P.nmen[j] + n); f = 2.0 * M * std::log2((double)M); cz[j] = 1; }
                fw[j] = f; fmax = std::max(fmax, f);
            }
            if (pe) {
                P.band_pad = (int)(atof(pe) * fmax / 16.0);
                for (int j = 0; j < ndgl; ++j) w[j] = (int)(fw[j] / 16.0) + P.band_pad;
            } else {
                P.band_pad = 5672;
                for (int j = 0; j < ndgl; ++j)
                    w[j] = (int)((cz[j] ? 0.2583 : 0.3975) * fw[j] + 17.80 * (P.nmen[j] + 1)) + P.band_pad;
            }
        }
        lat_bands(w, nranks, P.lat_first, P.lat_count);
    }
    P.lat0 = P.lat_first[rank];
    P.nlat = P.lat_count[rank];
    P.gpoff.assign(P.nlat + 1, 0);
    for (int l = 0; l < P.nlat; ++l) P.gpoff[l + 1] = P.gpoff[l] + P.nloen[P.lat0 + l];
    P.ngptot = P.gpoff[P.nlat];
    P.ngptotg = 0;
    for (int v : P.nloen) P.ngptotg += v;
    // record tables
    P.mrow0.assign(P.nump + 1, 0);
    for (int ml = 0; ml < P.nump; ++ml) P.mrow0[ml + 1] = P.mrow0[ml] + P.ndglu[P.myms[ml]];
    P.leg_rec_n.assign((size_t)P.mrow0[P.nump], -1);
    P.leg_rec_s.assign((size_t)P.mrow0[P.nump], -1);
    P.send_cnt.assign(nranks, 0); P.send_off.assign(nranks, 0);
    P.recv_cnt.assign(nranks, 0); P.recv_off.assign(nranks, 0);
    i64 rec = 0;
    for (int d = 0; d < nranks; ++d) {
        P.send_off[d] = rec;
        for (int g = P.lat_first[d]; g < P.lat_first[d] + P.lat_count[d]; ++g) {
            const int gn = g < P.ndgnh ? g : ndgl - 1 - g;
            for (int ml = 0; ml < P.nump; ++ml) {
                const int m = P.myms[ml];
                if (m > P.nmen[g]) continue;
                const int i = gn - (P.ndgnh - P.ndglu[m]);
                if (g < P.ndgnh) P.leg_rec_n[(size_t)(P.mrow0[ml] + i)] = (int)rec;
                else P.leg_rec_s[(size_t)(P.mrow0[ml] + i)] = (int)rec;
                ++rec;
            }
        }
        P.send_cnt[d] = rec - P.send_off[d];
    }
    P.nrec_leg = rec;
    P.latrow0.assign(P.nlat + 1, 0);
    for (int l = 0; l < P.nlat; ++l) P.latrow0[l + 1] = P.latrow0[l] + P.nmen[P.lat0 + l] + 1;
    P.fft_rec.assign((size_t)P.latrow0[P.nlat], -1);
    rec = 0;
    for (int s = 0; s < nranks; ++s) {
        P.recv_off[s] = rec;
        for (int l = 0; l < P.nlat; ++l) {
            const int g = P.lat0 + l;
            for (int m : P.ms_of[s]) {
                if (m > P.nmen[g]) continue;
                P.fft_rec[(size_t)(P.latrow0[l] + m)] = (int)rec;
                ++rec;
            }
        }
        P.recv_cnt[s] = rec - P.recv_off[s];
    }
    P.nrec_fft = rec;
    // Destination tables for the fused (peer-memory) transpositions: where a record produced on this rank
    // lives in the buffer of the rank that consumes it.
    //   inverse: (local m, lat) -> rank owning lat, index in ITS Fourier-side buffer  [src][lat][m of src]
    //   direct : (local lat, m) -> rank owning m,   index in ITS Legendre-side buffer [dest][lat of dest][local m]
    std::vector<int> rank_of_lat(ndgl, 0);
    for (int b = 0; b < nranks; ++b)
        for (int g = P.lat_first[b]; g < P.lat_first[b] + P.lat_count[b]; ++g) rank_of_lat[g] = b;
    P.leg_dst_rank_n.assign(P.leg_rec_n.size(), 0); P.leg_dst_rank_s.assign(P.leg_rec_n.size(), 0);
    P.leg_dst_rec_n.assign(P.leg_rec_n.size(), -1); P.leg_dst_rec_s.assign(P.leg_rec_n.size(), -1);
    {
        // position of m inside ms_of[rank] restricted to m <= nmen[g] is not closed-form: walk every rank's buffer
        std::vector<int> mloc(nsmax + 1, -1);
        for (int ml = 0; ml < P.nump; ++ml) mloc[P.myms[ml]] = ml;
        for (int b = 0; b < nranks; ++b) {
            i64 r = 0;
            for (int sr = 0; sr < nranks; ++sr)
                for (int g = P.lat_first[b]; g < P.lat_first[b] + P.lat_count[b]; ++g)
                    for (int m : P.ms_of[sr]) {
                        if (m > P.nmen[g]) continue;
                        if (sr == rank) {
                            const int ml = mloc[m];
                            const int gn = g < P.ndgnh ? g : ndgl - 1 - g;
                            const size_t at = (size_t)(P.mrow0[ml] + gn - (P.ndgnh - P.ndglu[m]));
                            if (g < P.ndgnh) { P.leg_dst_rank_n[at] = b; P.leg_dst_rec_n[at] = (int)r; }
                            else { P.leg_dst_rank_s[at] = b; P.leg_dst_rec_s[at] = (int)r; }
                        }
                        ++r;
                    }
        }
    }
    P.fft_dst_rank.assign(P.fft_rec.size(), 0);
    P.fft_dst_rec.assign(P.fft_rec.size(), -1);
    for (int r = 0; r < nranks; ++r) {          // rank r's Legendre-side buffer
        i64 q = 0;
        for (int d = 0; d < nranks; ++d)
            for (int g = P.lat_first[d]; g < P.lat_first[d] + P.lat_count[d]; ++g)
                for (int m : P.ms_of[r]) {
                    if (m > P.nmen[g]) continue;
                    if (d == rank) {
                        const size_t at = (size_t)(P.latrow0[g - P.lat0] + m);
                        P.fft_dst_rank[at] = r;
                        P.fft_dst_rec[at] = (int)q;
                    }
                    ++q;
                }
    }
    // ---- grid-point partition of the caller's arrays ----
    P.ngpband = P.ngptot;
    P.gp_eq = gp_eq && nranks > 1;
    if (P.gp_eq) {
        EctGpPartition G;
        int rc = ect_gp_partition(P.nloen, nranks, G);
        if (rc) return rc;
        P.gp_regions = G.regions;
        P.gp_all_segs = G.segs; P.gp_all_seg0 = G.seg0;
        P.gp_segs.assign(G.segs.begin() + G.seg0[rank], G.segs.begin() + G.seg0[rank + 1]);
        P.ngptot = 0;
        for (const EctGpSeg& sg : P.gp_segs) P.ngptot += sg.count;
        // band owner side (TRLTOG sends / TRGTOL receives): the pieces of every task that lie in my band
        P.xb_off.assign(nranks + 1, 0);
        P.xb_idx.clear(); P.xb_idx.reserve(P.ngpband);
        for (int p = 0; p < nranks; ++p) {
            for (int i = G.seg0[p]; i < G.seg0[p + 1]; ++i) {
                const EctGpSeg& sg = G.segs[i];
                if (sg.lat < P.lat0 || sg.lat >= P.lat0 + P.nlat) continue;
                const int b0 = P.gpoff[sg.lat - P.lat0] + sg.first;
                for (int j = 0; j < sg.count; ++j) P.xb_idx.push_back(b0 + j);
            }
            P.xb_off[p + 1] = (i64)P.xb_idx.size();
        }
        if ((int)P.xb_idx.size() != P.ngpband) { ect_set_error("ect_setup: grid-point partition does not cover the latitude band"); return ECT_ERR_GENERIC; }
        // grid-point task side: my pieces are ordered by latitude, the band owners hold ascending latitude ranges
        P.xg_off.assign(nranks + 1, 0);
        for (const EctGpSeg& sg : P.gp_segs) P.xg_off[rank_of_lat[sg.lat] + 1] += sg.count;
        for (int r = 0; r < nranks; ++r) P.xg_off[r + 1] += P.xg_off[r];
    }
    if (P.nrec_leg >= (1LL << 24) || P.nrec_fft >= (1LL << 24) || nranks > 127) {
        ect_set_error("ect_setup: record count overflows the packed (rank, record) table (2^24 records per rank)");
        return ECT_ERR_NOTIMPL;
    }
    return ECT_SUCCESS;
}
