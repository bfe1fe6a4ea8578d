// transi C face (reference src/transi/transi.h, transi.c, transi_module.F90) over the ect_* C ABI.
#include "../../include/transi_b200.h"
#include "../../include/ectrans_b200.h"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

extern "C" {

const char* trans_error_msg(int code) {      // transi.c:33-58
    switch (code) {
        case TRANS_SUCCESS: return "Trans: No error";
        case TRANS_ERROR: return "Trans: Error";
        case TRANS_NOTIMPL: return "Trans: Not (yet) implemented";
        case TRANS_MISSING_ARG: return "Trans: Required member of the argument structure is missing or not allocated";
        case TRANS_UNRECOGNIZED_ARG: return "Trans: Unrecognized argument";
        case TRANS_STALE_ARG: return "Trans: Passed argument was already used in a previous call";
        default: return ect_strerror(code);
    }
}

static int g_use_mpi = 0;
int trans_use_mpi(_bool v) { if (v) return TRANS_NOTIMPL; g_use_mpi = 0; return TRANS_SUCCESS; }
int trans_init(void) { return TRANS_SUCCESS; }

// src/transi/version.h: this library mirrors ecTrans 1.7.0 (the reference tree's VERSION)
const char* ectrans_version(void) { return "1.7.0"; }
unsigned int ectrans_version_int(void) { return 10700u; }
const char* ectrans_version_str(void) { return "1.7.0 (ectrans_b200: B200-native hot path)"; }
const char* ectrans_git_sha1(void) { return "not available"; }
const char* ectrans_git_sha1_abbrev(unsigned int) { return "not available"; }

// transi.h:121-194.  What cannot be honoured is refused, never silently ignored.
int trans_set_handles_limit(int limit) { return limit > 0 ? TRANS_SUCCESS : TRANS_ERROR; }       // handles are not limited here
int trans_set_radius(double radius) { return radius == 6371229.0 ? TRANS_SUCCESS : TRANS_NOTIMPL; }
int trans_set_nprtrv(int nprtrv) { return nprtrv == 1 ? TRANS_SUCCESS : TRANS_NOTIMPL; }
int trans_set_nprgpew(int nprgpew) { return nprgpew == 1 ? TRANS_SUCCESS : TRANS_NOTIMPL; }
int trans_set_leq_regions(_bool) { return TRANS_SUCCESS; }

int trans_new(struct Trans_t* t) {          // transi.c:60-82 defaults; every other member cleared
    if (!t) return TRANS_MISSING_ARG;
    memset(t, 0, sizeof(*t));
    t->handle = 0; t->llatlon = 0; t->lsplit = 1; t->flt = -1; t->fft = TRANS_FFTW;
    t->nsmax = -1; t->nmsmax = -1; t->ndgl = -1; t->nlon = -1;
    t->llam = 0; t->ndgux = -1; t->pexwn = 1.; t->peywn = 1.;
    t->myproc = 1; t->nproc = 1; t->nprtrw = 1;
    return TRANS_SUCCESS;
}

int trans_set_resol(struct Trans_t* t, int ndgl, const int* nloen) {
    if (!t || !nloen || ndgl <= 0) return TRANS_MISSING_ARG;
    t->ndgl = ndgl;
    free(t->nloen);
    t->nloen = (int*)malloc(sizeof(int) * ndgl);
    memcpy(t->nloen, nloen, sizeof(int) * ndgl);
    return TRANS_SUCCESS;
}
// lon-lat and LAM resolutions are recorded like the reference does (transi.c:93-125) and refused by trans_setup
int trans_set_resol_lonlat(struct Trans_t* t, int nlon, int nlat) {
    if (!t) return TRANS_MISSING_ARG;
    t->ndgl = (nlat % 2 == 0) ? nlat : nlat - 1; t->nlon = nlon; t->llatlon = (nlat % 2 == 0) ? 2 : 1;
    return TRANS_SUCCESS;
}
int trans_set_resol_lam(struct Trans_t* t, int nx, int ny, double dx, double dy) {
    if (!t) return TRANS_MISSING_ARG;
    t->ndgl = ny; t->nlon = nx; t->llam = 1;
    t->pexwn = 2. * 3.14159265358979323846 / ((double)nx * dx); t->peywn = 2. * 3.14159265358979323846 / ((double)ny * dy);
    return TRANS_SUCCESS;
}

int trans_set_trunc(struct Trans_t* t, int nsmax) {
    if (!t) return TRANS_MISSING_ARG;
    t->nsmax = nsmax;
    return TRANS_SUCCESS;
}
int trans_set_trunc_lam(struct Trans_t* t, int trunc_x, int trunc_y) {
    if (!t) return TRANS_MISSING_ARG;
    t->llam = 1; t->nmsmax = trunc_x; t->nsmax = trunc_y;
    return TRANS_SUCCESS;
}
static int set_path(char*& dst, const char* path) {
    if (!path) return TRANS_MISSING_ARG;
    free(dst);
    dst = (char*)malloc(1024);
    strncpy(dst, path, 1023); dst[1023] = 0;
    return TRANS_SUCCESS;
}
int trans_set_read(struct Trans_t* t, const char* filepath) { return t ? set_path(t->readfp, filepath) : TRANS_MISSING_ARG; }
int trans_set_write(struct Trans_t* t, const char* filepath) { return t ? set_path(t->writefp, filepath) : TRANS_MISSING_ARG; }
int trans_set_cache(struct Trans_t* t, const void* cache, size_t size) {
    if (!t) return TRANS_MISSING_ARG;
    t->cache = cache; t->cachesize = size;          // refused by trans_setup (in-memory FLT cache: out of scope)
    return TRANS_SUCCESS;
}

static int* alloc_int(long long n) { return (int*)calloc((size_t)(n > 0 ? n : 1), sizeof(int)); }

static int fill_inquire(struct Trans_t* t) {
    ect_info inf;
    int rc = ect_inquire(t->handle, &inf);
    if (rc) return rc;
    t->nspec2 = inf.nspec2; t->nspec = inf.nspec2 / 2; t->nspec2g = inf.nspec2g; t->nspec2mx = inf.nspec2;
    t->nump = inf.nump; t->ngptot = inf.ngptot; t->ngptotg = inf.ngptotg; t->ngptotmx = inf.ngptot;
    t->myproc = inf.rank + 1; t->nproc = inf.nranks; t->nprtrw = inf.nranks; t->nprtrns = inf.nranks;
    // grid-point decomposition of this face: the Fourier latitude bands (one task: one region)
    t->n_regions_NS = inf.nranks; t->n_regions_EW = 1; t->my_region_NS = inf.rank + 1; t->my_region_EW = 1;
    t->nfrstloff = inf.lat0; t->nptrfloff = inf.lat0;
    t->nlei3 = inf.ndgnh;
    int nspolegl = 0;
    rc = ect_inquire_rpnm(t->handle, nullptr, 0, &nspolegl, nullptr);
    if (rc) return rc;
    t->nspolegl = nspolegl;
    return TRANS_SUCCESS;
}

int trans_setup(struct Trans_t* t) {       // transi_module.F90 trans_setup
    if (!t) return TRANS_MISSING_ARG;
    if (t->ndgl <= 0) return TRANS_MISSING_ARG;
    if (t->llam || t->llatlon || t->flt > 0 || t->cache) return TRANS_NOTIMPL;        // LAM, lon-lat, FLT, in-memory cache: out of scope
    if (t->nsmax < 0) return TRANS_NOTIMPL;                                          // grid-only resolutions (LDGRIDONLY)
    std::vector<int> reg;
    const int* nloen = t->nloen;
    if (!nloen) {
        if (t->nlon <= 0) t->nlon = 2 * t->ndgl;           // transi_module.F90:698-716: regular grid of 2 NDGL points
        reg.assign(t->ndgl, t->nlon);
        nloen = reg.data();
    }
    ect_setup_opts o;
    memset(&o, 0, sizeof(o));
    o.nsmax = t->nsmax; o.ndgl = t->ndgl; o.nloen = nloen; o.nranks = 1; o.rank = 0; o.device = -1; o.precision = ECT_PREC_DP;
    if (t->readfp) o.flags |= ECT_SETUP_LEGPOL_DEFER;      // CDIO_LEGPOL = 'readf': the table comes from the file
    int h = 0;
    int rc = ect_setup(&o, &h);
    if (rc) return rc;
    t->handle = h;
    if (t->readfp && (rc = ect_read_legpol(h, t->readfp))) { ect_release(h); t->handle = 0; return rc; }
    if (t->writefp && (rc = ect_write_legpol(h, t->writefp))) return rc;
    return fill_inquire(t);
}

int trans_inquire(struct Trans_t* t, const char* varlist) {   // transi_module.F90 trans_inquire
    if (!t || !varlist) return TRANS_MISSING_ARG;
    ect_info inf;
    int rc = ect_inquire(t->handle, &inf);
    if (rc) return rc;
    std::string s(varlist);
    size_t pos = 0;
    while (pos <= s.size()) {
        size_t e = s.find(',', pos);
        if (e == std::string::npos) e = s.size();
        std::string v = s.substr(pos, e - pos);
        while (!v.empty() && v.front() == ' ') v.erase(v.begin());
        while (!v.empty() && v.back() == ' ') v.pop_back();
        pos = e + 1;
        if (v.empty()) continue;
        if (v == "nspec" || v == "nspec2" || v == "nspec2g" || v == "nspec2mx" || v == "nump" || v == "ngptot" ||
            v == "ngptotg" || v == "ngptotmx" || v == "nprtrw" || v == "myproc" || v == "nproc" || v == "nprtrns" ||
            v == "n_regions_NS" || v == "n_regions_EW" || v == "my_region_NS" || v == "my_region_EW" || v == "nfrstloff" ||
            v == "nptrfloff" || v == "nlei3" || v == "nspolegl") {
            rc = fill_inquire(t);
        } else if (v == "nmyms") {
            free(t->nmyms); t->nmyms = alloc_int(inf.nump);
            rc = ect_inquire_array(t->handle, ECT_ARR_MYMS, t->nmyms, inf.nump);
        } else if (v == "nasm0") {
            free(t->nasm0); t->nasm0 = alloc_int(inf.nsmax + 1);
            rc = ect_inquire_array(t->handle, ECT_ARR_NASM0, t->nasm0, inf.nsmax + 1);
            for (int m = 0; m <= inf.nsmax; ++m) t->nasm0[m] = t->nasm0[m] >= 0 ? t->nasm0[m] + 1 : -99;   // 1-based (suwavedi_mod.F90:112,131)
        } else if (v == "nnmeng") {
            free(t->nnmeng); t->nnmeng = alloc_int(inf.ndgl);
            rc = ect_inquire_array(t->handle, ECT_ARR_NMEN, t->nnmeng, inf.ndgl);
        } else if (v == "rmu" || v == "rgw") {
            double*& p = (v == "rmu") ? t->rmu : t->rgw;
            free(p); p = (double*)calloc(inf.ndgl, sizeof(double));
            rc = ect_inquire_array(t->handle, v == "rmu" ? ECT_ARR_RMU : ECT_ARR_RGW, p, inf.ndgl);
        } else if (v == "ngptotl") {
            free(t->ngptotl); t->ngptotl = alloc_int(inf.nranks); t->ngptotl[inf.rank] = inf.ngptot;
        } else if (v == "numpp") {
            free(t->numpp); t->numpp = alloc_int(inf.nranks);
            std::vector<int> pm(inf.nsmax + 1);
            rc = ect_inquire_array(t->handle, ECT_ARR_NPROCM, pm.data(), inf.nsmax + 1);
            for (int m = 0; m <= inf.nsmax; ++m) t->numpp[pm[m]]++;
        } else if (v == "nallms" || v == "nptrms") {
            std::vector<int> pm(inf.nsmax + 1);
            rc = ect_inquire_array(t->handle, ECT_ARR_NPROCM, pm.data(), inf.nsmax + 1);
            free(t->nallms); free(t->nptrms);
            t->nallms = alloc_int(inf.nsmax + 1); t->nptrms = alloc_int(inf.nranks);
            int k = 0;
            for (int r = 0; r < inf.nranks; ++r) {
                t->nptrms[r] = k + 1;
                for (int m = 0; m <= inf.nsmax; ++m) if (pm[m] == r) t->nallms[k++] = m;
            }
        } else if (v == "nvalue") {
            std::vector<int> ms(inf.nump);
            rc = ect_inquire_array(t->handle, ECT_ARR_MYMS, ms.data(), inf.nump);
            free(t->nvalue); t->nvalue = alloc_int(inf.nspec2);
            int k = 0;
            for (int m : ms) for (int n = m; n <= inf.nsmax; ++n) { t->nvalue[k++] = n; t->nvalue[k++] = n; }
        } else if (v == "nultpp" || v == "nptrls") {
            free(t->nultpp); free(t->nptrls);
            t->nultpp = alloc_int(inf.nranks); t->nptrls = alloc_int(inf.nranks);
            rc = ect_inquire_array(t->handle, ECT_ARR_LATCOUNT, t->nultpp, inf.nranks);
            if (!rc) rc = ect_inquire_array(t->handle, ECT_ARR_LATFIRST, t->nptrls, inf.nranks);
            for (int r = 0; r < inf.nranks; ++r) t->nptrls[r] += 1;
        } else if (v == "npossp") {                // start of each W-set's coefficients in the global spectral array, 1-based
            std::vector<int> pm(inf.nsmax + 1);
            rc = ect_inquire_array(t->handle, ECT_ARR_NPROCM, pm.data(), inf.nsmax + 1);
            free(t->npossp); t->npossp = alloc_int(inf.nranks + 1);
            std::vector<long long> cnt(inf.nranks, 0);
            for (int m = 0; m <= inf.nsmax; ++m) cnt[pm[m]] += 2 * (inf.nsmax - m + 1);
            t->npossp[0] = 1;
            for (int r = 0; r < inf.nranks; ++r) t->npossp[r + 1] = t->npossp[r] + (int)cnt[r];
        } else if (v == "ndim0g") {                // start of wavenumber m in the global array ordered by W-set (suwavedi_mod.F90:150-160)
            std::vector<int> pm(inf.nsmax + 1);
            rc = ect_inquire_array(t->handle, ECT_ARR_NPROCM, pm.data(), inf.nsmax + 1);
            free(t->ndim0g); t->ndim0g = alloc_int(inf.nsmax + 1);
            int pos = 1;
            for (int r = 0; r < inf.nranks; ++r)
                for (int m = 0; m <= inf.nsmax; ++m) if (pm[m] == r) { t->ndim0g[m] = pos; pos += 2 * (inf.nsmax - m + 1); }
        } else if (v == "n_regions") {
            free(t->n_regions); t->n_regions = alloc_int(inf.nranks);
            for (int r = 0; r < inf.nranks; ++r) t->n_regions[r] = 1;
        } else if (v == "nfrstlat" || v == "nlstlat" || v == "nptrfrstlat" || v == "nptrlstlat") {
            // latitude bands are never split on this face: the pointers into NSTA / NONL are the latitudes themselves
            std::vector<int> f(inf.nranks), c(inf.nranks);
            rc = ect_inquire_array(t->handle, ECT_ARR_LATFIRST, f.data(), inf.nranks);
            if (!rc) rc = ect_inquire_array(t->handle, ECT_ARR_LATCOUNT, c.data(), inf.nranks);
            int*& p = v == "nfrstlat" ? t->nfrstlat : v == "nlstlat" ? t->nlstlat : v == "nptrfrstlat" ? t->nptrfrstlat : t->nptrlstlat;
            free(p); p = alloc_int(inf.nranks);
            const bool last = (v == "nlstlat" || v == "nptrlstlat");
            for (int r = 0; r < inf.nranks; ++r) p[r] = last ? f[r] + c[r] : f[r] + 1;
        } else if (v == "nptrlat") {
            free(t->nptrlat); t->nptrlat = alloc_int(inf.ndgl);
            for (int j = 0; j < inf.ndgl; ++j) t->nptrlat[j] = j + 1;
        } else if (v == "nsta" || v == "nonl") {   // (ndgl + n_regions_NS - 1) x n_regions_EW; whole latitudes: first point 1, NLOEN points
            std::vector<int> nl(inf.ndgl);
            rc = ect_inquire_array(t->handle, ECT_ARR_NLOEN, nl.data(), inf.ndgl);
            int*& p = v == "nsta" ? t->nsta : t->nonl;
            free(p); p = alloc_int(inf.ndgl + inf.nranks - 1);
            for (int j = 0; j < inf.ndgl; ++j) p[j] = v == "nsta" ? 1 : nl[j];
        } else if (v == "ldsplitlat") {
            free(t->ldsplitlat); t->ldsplitlat = alloc_int(inf.ndgl);
        } else if (v == "ndglu") {
            free(t->ndglu); t->ndglu = alloc_int(inf.nsmax + 1);
            rc = ect_inquire_array(t->handle, ECT_ARR_NDGLU, t->ndglu, inf.nsmax + 1);
        } else if (v == "npms") {                  // 1-based column of wavenumber m in RPNM (trans_inq.F90), -1: not on this task
            free(t->npms); t->npms = alloc_int(inf.nsmax + 1);
            int ns = 0;
            rc = ect_inquire_rpnm(t->handle, nullptr, 0, &ns, t->npms);
            for (int m = 0; m <= inf.nsmax; ++m) if (t->npms[m] >= 0) t->npms[m] += 1;
        } else if (v == "rpnm") {
            int ns = 0;
            rc = ect_inquire_rpnm(t->handle, nullptr, 0, &ns, nullptr);
            if (!rc) {
                free(t->rpnm); t->rpnm = (double*)calloc((size_t)std::max(1, ns) * inf.ndgnh, sizeof(double));
                t->nlei3 = inf.ndgnh; t->nspolegl = ns;
                rc = ect_inquire_rpnm(t->handle, t->rpnm, (long long)ns * inf.ndgnh, &ns, nullptr);
            }
        } else if (v == "rlapin") {                // RLAPIN(-1:nsmax+2) = -a^2 / (n (n + 1)), 0 for n <= 0 (suleg_mod.F90)
            free(t->rlapin); t->rlapin = (double*)calloc(inf.nsmax + 4, sizeof(double));
            for (int n = 1; n <= inf.nsmax + 2; ++n) t->rlapin[n + 1] = -(6371229.0 * 6371229.0) / ((double)n * (double)(n + 1));
        } else {
            return TRANS_UNRECOGNIZED_ARG;
        }
        if (rc) return rc;
    }
    return TRANS_SUCCESS;
}

struct InvTrans_t new_invtrans(struct Trans_t* t) {
    struct InvTrans_t a;
    memset(&a, 0, sizeof(a));
    a.nproma = t ? t->ngptot : 0; a.ngpblks = 1; a.trans = t;
    return a;
}
struct DirTrans_t new_dirtrans(struct Trans_t* t) {
    struct DirTrans_t a;
    memset(&a, 0, sizeof(a));
    a.nproma = t ? t->ngptot : 0; a.ngpblks = 1; a.trans = t;
    return a;
}
struct SpecNorm_t new_specnorm(struct Trans_t* t) {
    struct SpecNorm_t a;
    memset(&a, 0, sizeof(a));
    a.nmaster = 1; a.nfld = 0; a.trans = t;
    return a;
}

int trans_invtrans(struct InvTrans_t* a) {      // transi_module.F90:2032-2061 argument checks
    if (!a || !a->trans) return TRANS_MISSING_ARG;
    if (a->count > 0) return TRANS_STALE_ARG;
    a->count++;
    if (a->rmeanu || a->rmeanv) return TRANS_NOTIMPL;
    if (a->nscalar > 0 && !a->rspscalar) return TRANS_MISSING_ARG;
    if (a->nvordiv > 0 && (!a->rspvor || !a->rspdiv)) return TRANS_MISSING_ARG;
    if ((a->nscalar > 0 || a->nvordiv > 0) && !a->rgp) return TRANS_MISSING_ARG;
    ect_inv_args e;
    memset(&e, 0, sizeof(e));
    e.memspace = ECT_MEM_HOST;
    e.nproma = a->lglobal ? a->trans->ngptot : a->nproma;     // one rank: global field == one block of ngptot
    if (a->lglobal && a->trans->nproc != 1) return TRANS_NOTIMPL;
    e.scders = a->lscalarders; e.vorgp = a->lvordivgp; e.divgp = a->lvordivgp; e.uvder = a->luvder_EW;
    e.spvor = a->rspvor; e.spdiv = a->rspdiv; e.nuv = a->nvordiv;
    e.spscalar = a->rspscalar; e.nscalar = a->nscalar;
    e.gp = a->rgp;
    return ect_inv_trans(a->trans->handle, &e);
}

int trans_dirtrans(struct DirTrans_t* a) {
    if (!a || !a->trans) return TRANS_MISSING_ARG;
    if (a->count > 0) return TRANS_STALE_ARG;
    a->count++;
    if (a->rmeanu || a->rmeanv) return TRANS_NOTIMPL;
    if (!a->rgp) return TRANS_MISSING_ARG;
    if (a->nscalar > 0 && !a->rspscalar) return TRANS_MISSING_ARG;
    if (a->nvordiv > 0 && (!a->rspvor || !a->rspdiv)) return TRANS_MISSING_ARG;
    if (a->lglobal && a->trans->nproc != 1) return TRANS_NOTIMPL;
    ect_dir_args e;
    memset(&e, 0, sizeof(e));
    e.memspace = ECT_MEM_HOST;
    e.nproma = a->lglobal ? a->trans->ngptot : a->nproma;
    e.nuv = a->nvordiv; e.nscalar = a->nscalar;
    e.gp = a->rgp;
    e.spvor = a->rspvor; e.spdiv = a->rspdiv; e.spscalar = a->rspscalar;
    return ect_dir_trans(a->trans->handle, &e);
}

struct DirTransAdj_t new_dirtrans_adj(struct Trans_t* t) {
    struct DirTransAdj_t a;
    memset(&a, 0, sizeof(a));
    a.nproma = t ? t->ngptot : 0; a.ngpblks = 1; a.trans = t;
    return a;
}
int trans_dirtrans_adj(struct DirTransAdj_t* a) {
    if (!a || !a->trans) return TRANS_MISSING_ARG;
    if (a->count > 0) return TRANS_STALE_ARG;
    a->count++;
    if (a->rmeanu || a->rmeanv) return TRANS_NOTIMPL;
    if (!a->rgp) return TRANS_MISSING_ARG;
    if (a->nscalar > 0 && !a->rspscalar) return TRANS_MISSING_ARG;
    if (a->nvordiv > 0 && (!a->rspvor || !a->rspdiv)) return TRANS_MISSING_ARG;
    if (a->lglobal && a->trans->nproc != 1) return TRANS_NOTIMPL;
    ect_dir_args e;
    memset(&e, 0, sizeof(e));
    e.memspace = ECT_MEM_HOST;
    e.nproma = a->lglobal ? a->trans->ngptot : a->nproma;
    e.nuv = a->nvordiv; e.nscalar = a->nscalar;
    e.gp = a->rgp;
    e.spvor = (double*)a->rspvor; e.spdiv = (double*)a->rspdiv; e.spscalar = (double*)a->rspscalar;
    return ect_dir_transad(a->trans->handle, &e);
}

struct InvTransAdj_t new_invtrans_adj(struct Trans_t* t) {
    struct InvTransAdj_t a;
    memset(&a, 0, sizeof(a));
    a.nproma = t ? t->ngptot : 0; a.ngpblks = 1; a.trans = t;
    return a;
}
int trans_invtrans_adj(struct InvTransAdj_t* a) {
    if (!a || !a->trans) return TRANS_MISSING_ARG;
    if (a->count > 0) return TRANS_STALE_ARG;
    a->count++;
    if (a->rmeanu || a->rmeanv) return TRANS_NOTIMPL;
    if (a->lscalarders || a->luvder_EW || a->lvordivgp) return TRANS_NOTIMPL;
    if (!a->rgp) return TRANS_MISSING_ARG;
    if (a->nscalar > 0 && !a->rspscalar) return TRANS_MISSING_ARG;
    if (a->nvordiv > 0 && (!a->rspvor || !a->rspdiv)) return TRANS_MISSING_ARG;
    if (a->lglobal && a->trans->nproc != 1) return TRANS_NOTIMPL;
    ect_inv_args e;
    memset(&e, 0, sizeof(e));
    e.memspace = ECT_MEM_HOST;
    e.nproma = a->lglobal ? a->trans->ngptot : a->nproma;
    e.nuv = a->nvordiv; e.nscalar = a->nscalar;
    e.gp = (double*)a->rgp;
    e.spvor = a->rspvor; e.spdiv = a->rspdiv; e.spscalar = a->rspscalar;
    return ect_inv_transad(a->trans->handle, &e);
}

// owners: the reference counts tasks from 1
static int owners0(const int* own1, int nfld, int nproc, std::vector<int>& out) {
    if (nfld > 0 && !own1) return TRANS_MISSING_ARG;
    out.resize(nfld > 0 ? nfld : 0);
    for (int f = 0; f < nfld; ++f) {
        if (own1[f] < 1 || own1[f] > nproc) return TRANS_ERROR;
        out[f] = own1[f] - 1;
    }
    return TRANS_SUCCESS;
}
struct DistGrid_t new_distgrid(struct Trans_t* t) { struct DistGrid_t a; memset(&a, 0, sizeof(a)); a.nproma = t ? t->ngptot : 0; a.ngpblks = 1; a.trans = t; return a; }
struct GathGrid_t new_gathgrid(struct Trans_t* t) { struct GathGrid_t a; memset(&a, 0, sizeof(a)); a.nproma = t ? t->ngptot : 0; a.ngpblks = 1; a.trans = t; return a; }
struct DistSpec_t new_distspec(struct Trans_t* t) { struct DistSpec_t a; memset(&a, 0, sizeof(a)); a.trans = t; return a; }
struct GathSpec_t new_gathspec(struct Trans_t* t) { struct GathSpec_t a; memset(&a, 0, sizeof(a)); a.trans = t; return a; }
int trans_distgrid(struct DistGrid_t* a) {
    if (!a || !a->trans || !a->rgp) return TRANS_MISSING_ARG;
    if (a->count > 0) return TRANS_STALE_ARG;
    a->count++;
    std::vector<int> own; int rc = owners0(a->nfrom, a->nfld, a->trans->nproc, own);
    return rc ? rc : ect_dist_grid(a->trans->handle, a->rgpg, a->nfld, a->nproma, own.data(), a->rgp);
}
int trans_gathgrid(struct GathGrid_t* a) {
    if (!a || !a->trans || !a->rgp) return TRANS_MISSING_ARG;
    if (a->count > 0) return TRANS_STALE_ARG;
    a->count++;
    std::vector<int> own; int rc = owners0(a->nto, a->nfld, a->trans->nproc, own);
    return rc ? rc : ect_gath_grid(a->trans->handle, a->rgp, a->nfld, a->nproma, own.data(), a->rgpg);
}
int trans_distspec(struct DistSpec_t* a) {
    if (!a || !a->trans || !a->rspec) return TRANS_MISSING_ARG;
    if (a->count > 0) return TRANS_STALE_ARG;
    a->count++;
    std::vector<int> own; int rc = owners0(a->nfrom, a->nfld, a->trans->nproc, own);
    return rc ? rc : ect_dist_spec(a->trans->handle, a->rspecg, a->nfld, own.data(), a->rspec);
}
int trans_gathspec(struct GathSpec_t* a) {
    if (!a || !a->trans || !a->rspec) return TRANS_MISSING_ARG;
    if (a->count > 0) return TRANS_STALE_ARG;
    a->count++;
    std::vector<int> own; int rc = owners0(a->nto, a->nfld, a->trans->nproc, own);
    return rc ? rc : ect_gath_spec(a->trans->handle, a->rspec, a->nfld, own.data(), a->rspecg);
}

struct VorDivToUV_t new_vordiv_to_UV(void) { struct VorDivToUV_t a; memset(&a, 0, sizeof(a)); return a; }
// transi_module.F90:2663-2743
int trans_vordiv_to_UV(struct VorDivToUV_t* a) {
    if (!a) return TRANS_MISSING_ARG;
    if (a->count > 0) return TRANS_STALE_ARG;
    a->count = 1;
    if (a->ncoeff == 0 || a->nsmax == 0) return TRANS_MISSING_ARG;
    if (!a->rspvor || !a->rspdiv || !a->rspu || !a->rspv) return TRANS_MISSING_ARG;
    if (a->ncoeff != (a->nsmax + 1) * (a->nsmax + 2)) return TRANS_ERROR;      // distributed ncoeff needs a Trans_t: use ect_vordiv_to_uv
    const int rc = ect_vordiv_to_uv(0, a->nsmax, a->rspvor, a->rspdiv, a->rspu, a->rspv, a->nfld, ECT_MEM_HOST);
    return rc == ECT_SUCCESS ? TRANS_SUCCESS : (rc >= -5 ? rc : TRANS_ERROR);
}

int trans_specnorm(struct SpecNorm_t* a) {
    if (!a || !a->trans || !a->rspec || !a->rnorm || a->nfld <= 0) return TRANS_MISSING_ARG;
    if (a->count > 0) return TRANS_STALE_ARG;
    a->count++;
    return ect_specnorm_met(a->trans->handle, a->rspec, a->nfld, ECT_MEM_HOST, a->rmet, a->rnorm);
}

int trans_delete(struct Trans_t* t) {
    if (!t) return TRANS_MISSING_ARG;
    int rc = t->handle ? ect_release(t->handle) : 0;
    int** ip[] = {&t->nloen, &t->ngptotl, &t->nmyms, &t->nasm0, &t->numpp, &t->npossp, &t->nallms, &t->nptrms, &t->ndim0g,
                  &t->nvalue, &t->n_regions, &t->nfrstlat, &t->nlstlat, &t->nptrlat, &t->nptrfrstlat, &t->nptrlstlat, &t->nsta,
                  &t->nonl, &t->ldsplitlat, &t->nultpp, &t->nptrls, &t->nnmeng, &t->npms, &t->ndglu, &t->mvalue};
    for (int** p : ip) { free(*p); *p = nullptr; }
    double** dp[] = {&t->rmu, &t->rgw, &t->rpnm, &t->rlapin, &t->pweight};
    for (double** p : dp) { free(*p); *p = nullptr; }
    free(t->readfp); t->readfp = nullptr;
    free(t->writefp); t->writefp = nullptr;
    t->handle = 0;
    return rc;
}

int trans_finalize(void) { return ect_finalize(); }

}  // extern "C"
