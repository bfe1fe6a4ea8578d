// transi C face (reference src/transi/transi.h, transi.c, transi_module.F90) over the ect_* C ABI.
#include "../../include/transi_b200.h"
#include "../../include/ectrans_b200.h"
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
int trans_use_mpi(int v) { if (v) return TRANS_NOTIMPL; g_use_mpi = 0; return TRANS_SUCCESS; }
int trans_init(void) { return TRANS_SUCCESS; }

int trans_new(struct Trans_t* t) {          // transi.c: defaults
    if (!t) return TRANS_MISSING_ARG;
    memset(t, 0, sizeof(*t));
    t->nsmax = -1; t->nlon = -1; t->flt = -1; t->lsplit = 1; t->handle = 0;
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

int trans_set_trunc(struct Trans_t* t, int nsmax) {
    if (!t) return TRANS_MISSING_ARG;
    t->nsmax = nsmax;
    return TRANS_SUCCESS;
}

static int* alloc_int(long long n) { return (int*)calloc((size_t)(n > 0 ? n : 1), sizeof(int)); }

static int fill_inquire(struct Trans_t* t) {
    ect_info inf;
    int rc = ect_inquire(t->handle, &inf);
    if (rc) return rc;
    t->nspec2 = inf.nspec2; t->nspec = inf.nspec2 / 2; t->nspec2g = inf.nspec2g; t->nspec2mx = inf.nspec2;
    t->nump = inf.nump; t->ngptot = inf.ngptot; t->ngptotg = inf.ngptotg; t->ngptotmx = inf.ngptot;
    t->myproc = inf.rank + 1; t->nproc = inf.nranks; t->nprtrw = inf.nranks;
    return TRANS_SUCCESS;
}

int trans_setup(struct Trans_t* t) {
    if (!t) return TRANS_MISSING_ARG;
    if (t->ndgl <= 0) return TRANS_MISSING_ARG;
    if (t->llatlon || t->flt > 0) return TRANS_NOTIMPL;
    std::vector<int> reg;
    const int* nloen = t->nloen;
    if (!nloen) {
        if (t->nlon <= 0) return TRANS_MISSING_ARG;
        reg.assign(t->ndgl, t->nlon);
        nloen = reg.data();
    }
    if (t->nsmax < 0) t->nsmax = (2 * t->ndgl - 1) / 2;      // linear-grid default as in transi_module.F90
    ect_setup_opts o;
    memset(&o, 0, sizeof(o));
    o.nsmax = t->nsmax; o.ndgl = t->ndgl; o.nloen = nloen; o.nranks = 1; o.rank = 0; o.device = -1; o.precision = ECT_PREC_DP;
    int h = 0;
    int rc = ect_setup(&o, &h);
    if (rc) return rc;
    t->handle = h;
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
            v == "ngptotg" || v == "ngptotmx" || v == "nprtrw" || v == "myproc" || v == "nproc") {
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
    int** ip[] = {&t->nloen, &t->ngptotl, &t->nmyms, &t->nasm0, &t->numpp, &t->nallms, &t->nptrms, &t->nvalue,
                  &t->nultpp, &t->nptrls, &t->nnmeng};
    for (int** p : ip) { free(*p); *p = nullptr; }
    free(t->rmu); t->rmu = nullptr;
    free(t->rgw); t->rgw = nullptr;
    t->handle = 0;
    return rc;
}

int trans_finalize(void) { return ect_finalize(); }

}  // extern "C"
