// Grid-point decomposition of the reference's default setup (LDEQ_REGIONS=T, LDSPLIT=T; ectrans-benchmark.F90:385):
//   eq_regions   common/internal/eq_regions_mod.F90:76-230   bands x regions per band, equal-area partition
//   SUMPLATBEQ   common/internal/sumplatbeq_mod.F90:84-150   points per band, (split) first / last latitude
//   SUMPLAT      common/internal/sumplat_mod.F90:138-150
//   SUSTAONL     common/internal/sustaonl_mod.F90:118-196    first point / count per (latitude, region)
//   PE2SET       common/internal/pe2set_mod.F90:70-84        task = regions of the bands before + region
// Host only.  SUSTAONL hands out the points of a band one at a time, always from the latitude whose next point lies
// furthest west; the reference scans all latitudes of the band per point (O(points x latitudes)), here a heap keyed
// by (angle, latitude) yields the same order in O(points log latitudes).
#include "ect_internal.h"
#include <cmath>
#include <queue>
#include <numeric>

static long long f_nint(double x) { return x >= 0 ? (long long)std::floor(x + 0.5) : -(long long)std::floor(-x + 0.5); }

// Gamma(x) by the polynomial the reference uses for the area of the sphere (eq_regions_mod.F90:282-330).  For odd task
// counts the ideal collar sizes come in exact-tie pairs (r1 + r2 = integer + 1/2 with r2 = k + 1/2), so which collar gets
// the extra region hangs on the last bits of the area: std::tgamma would flip e.g. eq_regions(31) from the reference's
// [1 6 8 9 6 1] to [1 6 9 8 6 1].
static double eq_gamma(double x) {
    static const double c[14] = {0.999999999999999990e+00, -0.422784335098466784e+00, -0.233093736421782878e+00,
                                 0.191091101387638410e+00, -0.024552490005641278e+00, -0.017645244547851414e+00,
                                 0.008023273027855346e+00, -0.000804329819255744e+00, -0.000360837876648255e+00,
                                 0.000145596568617526e+00, -0.000017545539395205e+00, -0.000002591225267689e+00,
                                 0.000001337767384067e+00, -0.000000199542863674e+00};
    const int n = (int)f_nint(x - 2.0);
    double w = x - (double)(n + 2);
    double y = c[13];
    for (int i = 12; i >= 0; --i) y = y * w + c[i];
    if (n > 0) {
        w = x - 1.0;
        for (int k = 2; k <= n; ++k) w *= x - (double)k;
    } else {
        w = 1.0;
        for (int k = 0; k < -n; ++k) y *= x + (double)k;
    }
    return w / y;
}

int ect_eq_regions(int n, std::vector<int>& regions) {
    regions.clear();
    if (n < 1) return ECT_ERR_BADARG;
    if (n == 1) { regions.push_back(1); return ECT_SUCCESS; }
    const double pi = 2.0 * std::asin(1.0);
    const double area = (2.0 * std::pow(pi, 1.5) / eq_gamma(1.5)) / (double)n;      // area of the unit sphere / n
    // area of a cap: (4 pi) * sin^2, associated as the reference's 4*pi*sin(s/2)**2 -- the ties described above make
    // even this last bit matter
    auto cap = [&](double s) { const double h = std::sin(s / 2.0); return (4.0 * pi) * (h * h); };
    const double polar = n == 2 ? 0.5 * pi : 2.0 * std::asin(0.5 * std::sqrt(area / pi));
    const double ideal = std::sqrt(area);
    int collars = 0;
    if (n > 2 && ideal > 0) collars = (int)std::max<long long>(1, f_nint((pi - 2.0 * polar) / ideal));
    std::vector<double> r;
    r.push_back(1.0);
    if (collars > 0) {
        const double fit = (pi - 2.0 * polar) / (double)collars;
        for (int c = 1; c <= collars; ++c) r.push_back((cap(polar + c * fit) - cap(polar + (c - 1) * fit)) / area);
    }
    r.push_back(1.0);
    double carry = 0.0;
    for (double v : r) {
        const int k = (int)f_nint(v + carry);
        carry += v - (double)k;
        regions.push_back(k);
    }
    if (std::accumulate(regions.begin(), regions.end(), 0) != n) { ect_set_error("eq_regions: N /= sum(n_regions)"); return ECT_ERR_GENERIC; }
    return ECT_SUCCESS;
}

int ect_gp_partition(const std::vector<int>& nloen, int nproc, EctGpPartition& G) {
    const int ndgl = (int)nloen.size();
    int rc = ect_eq_regions(nproc, G.regions);
    if (rc) return rc;
    const int nbands = (int)G.regions.size();
    G.seg0.assign(nproc + 1, 0);
    G.segs.clear();
    G.band_first.assign(nbands, 0); G.band_last.assign(nbands, 0); G.band_points.assign(nbands, 0);
    if (nproc == 1) {
        for (int j = 0; j < ndgl; ++j) G.segs.push_back({j, 0, nloen[j]});
        G.seg0[1] = ndgl; G.band_last[0] = ndgl - 1;
        G.band_points[0] = std::accumulate(nloen.begin(), nloen.end(), 0LL);
        return ECT_SUCCESS;
    }
    const long long total = std::accumulate(nloen.begin(), nloen.end(), 0LL);
    long long share = total / nproc;
    const long long extra = total - share * nproc;          // KRESTM: that many tasks carry one point more
    if (extra > 0) ++share;
    if (share < (nloen[0] - 1) / (G.regions[0] + G.regions[1]) + 1) {
        ect_set_error("SUMPLATBEQ: NPROC TOO BIG FOR THIS RESOLUTION, LDSPLIT=T");
        return ECT_ERR_BADARG;
    }
    // bands: points per band; a band ends inside latitude `cut` when the running count overshoots (the rest of that
    // latitude is carried into the next band as left_over)
    std::vector<int> last(nbands, ndgl - 1), cut(nbands, -1);
    {
        long long left_over = 0;
        int lat = -1, task = 0;
        for (int a = 0; a < nbands; ++a) {
            long long want = 0;
            for (int b = 0; b < G.regions[a]; ++b) { ++task; want += (task <= extra || extra == 0) ? share : share - 1; }
            G.band_points[a] = want;
            long long have = left_over;
            // a band inside what is left of one latitude cannot be described by SUMPLATBEQ -- except the last band, which
            // may be exactly the rest of the last latitude
            if (a > 0 && have >= want && !(a == nbands - 1 && have == want && lat == ndgl - 1)) {
                ect_set_error("SUMPLATBEQ: NPROC TOO BIG FOR THIS RESOLUTION, LDSPLIT=T (a band of %lld points inside one latitude)", want);
                return ECT_ERR_BADARG;
            }
            while (++lat < ndgl) {
                if (have + nloen[lat] < want) { have += nloen[lat]; continue; }
                last[a] = lat;
                if (have + nloen[lat] == want) { left_over = 0; cut[a] = -1; }
                else { left_over = nloen[lat] - (want - have); cut[a] = lat; }
                break;
            }
        }
    }
    G.band_first[0] = 0; G.band_last[nbands - 1] = ndgl - 1;
    for (int a = 0; a + 1 < nbands; ++a) {
        if (cut[a] < 0) { G.band_first[a + 1] = last[a] + 1; G.band_last[a] = last[a]; }
        else { G.band_first[a + 1] = cut[a]; G.band_last[a] = cut[a]; }
    }
    // regions of every band
    std::vector<long long> before(ndgl + 1, 0);
    for (int j = 0; j < ndgl; ++j) before[j + 1] = before[j] + nloen[j];
    long long done = 0;
    int pe = 0;
    for (int a = 0; a < nbands; ++a) {
        const int l0 = G.band_first[a], nl = G.band_last[a] - l0 + 1, nb = G.regions[a];
        const long long pts = G.band_points[a];
        std::vector<int> next(nl, 1), end(nl);                // 1-based next point / last point of the band per latitude
        for (int k = 0; k < nl; ++k) end[k] = nloen[l0 + k];
        next[0] = (int)(done - before[l0]) + 1;
        long long span = nloen[l0] - next[0] + 1;
        for (int k = 1; k < nl; ++k) span += nloen[l0 + k];
        end[nl - 1] = (int)(nloen[l0 + nl - 1] - span + pts);
        if (nl < 1 || next[0] < 1 || next[0] > nloen[l0] || end[nl - 1] < 0 || end[nl - 1] > nloen[l0 + nl - 1]) {
            ect_set_error("SUSTAONL: inconsistent partitioning (band %d)", a);
            return ECT_ERR_GENERIC;
        }
        typedef std::pair<long long, int> Key;                 // (angle of the next point in 1/1000 degree, latitude)
        std::priority_queue<Key, std::vector<Key>, std::greater<Key>> heap;
        auto angle = [&](int k) { return f_nint((double)(next[k] - 1) * (360000.0 / (double)nloen[l0 + k])); };
        for (int k = 0; k < nl; ++k) if (next[k] <= end[k]) heap.push(Key(angle(k), k));
        const long long base = pts / nb, more = pts - base * nb;
        for (int b = 0; b < nb; ++b, ++pe) {
            std::vector<int> sta(nl, 0), cnt(nl, 0);
            for (long long i = 0, n = base + (b < more ? 1 : 0); i < n; ++i) {
                if (heap.empty()) { ect_set_error("SUSTAONL: inconsistent partitioning"); return ECT_ERR_GENERIC; }
                const int k = heap.top().second;
                heap.pop();
                if (cnt[k] == 0) sta[k] = next[k];
                ++cnt[k]; ++next[k];
                if (next[k] <= end[k]) heap.push(Key(angle(k), k));
            }
            for (int k = 0; k < nl; ++k) if (cnt[k] > 0) G.segs.push_back({l0 + k, sta[k] - 1, cnt[k]});
            G.seg0[pe + 1] = (int)G.segs.size();
        }
        done += pts;
    }
    return ECT_SUCCESS;
}

// C ABI (host only, no handle): the decomposition for nproc tasks.  regions: int[nproc] (first *nbands valid);
// seg0: int[nproc + 1]; segs: int[3 * capacity] (latitude, first point, count -- 0-based), in every task's local order.
extern "C" int ect_gridpoint_partition(int ndgl, const int* nloen, int nproc, int* nbands, int* regions, int* seg0,
                                       int* segs, long long capacity_segs, long long* nsegs) {
    if (!nloen || ndgl < 1 || nproc < 1) return ECT_ERR_BADARG;
    EctGpPartition G;
    int rc = ect_gp_partition(std::vector<int>(nloen, nloen + ndgl), nproc, G);
    if (rc) return rc;
    if (nbands) *nbands = (int)G.regions.size();
    if (regions) for (size_t i = 0; i < G.regions.size(); ++i) regions[i] = G.regions[i];
    if (seg0) for (int i = 0; i <= nproc; ++i) seg0[i] = G.seg0[i];
    if (nsegs) *nsegs = (long long)G.segs.size();
    if (segs) {
        if ((long long)G.segs.size() > capacity_segs) return ECT_ERR_BADARG;
        for (size_t i = 0; i < G.segs.size(); ++i) { segs[3 * i] = G.segs[i].lat; segs[3 * i + 1] = G.segs[i].first; segs[3 * i + 2] = G.segs[i].count; }
    }
    return ECT_SUCCESS;
}
