/*
 * transi_b200.h -- the transi C face of the reference (src/transi/transi.h) on top of the B200 library.
 *
 * Subset: the functions and struct members the hot path needs (transi.h:110-233, 293, 422, 667-685;
 * structs transi.h:701-830, 905-921, 986-1010, 1190-1200).  Same names, argument meaning, layouts
 * (rspscalar[nspec2][nscalar], rgp[ngpblks][nfld][nproma], lglobal: rgp[nfld][ngptotg]) and the same
 * return codes (TRANS_SUCCESS 0, -1 error, -2 not implemented, -3 missing arg, -4 unrecognised arg,
 * -5 stale arg; transi.c:33-58).  Argument structs are single use (count guard,
 * transi_module.F90:1949-1954).  Also trans_dirtrans_adj / trans_invtrans_adj (transi.h:354, 491) and
 * trans_distgrid / gathgrid / distspec / gathspec (transi.h:520-616).  Out of scope: vordiv_to_UV, LAM, I/O cache.
 * A caller of the reference includes this header instead of "ectrans/transi.h" and links
 * libectrans_b200.so instead of libtransi_dp.so.
 */
#ifndef TRANSI_B200_H
#define TRANSI_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define TRANS_SUCCESS         0
#define TRANS_ERROR          -1
#define TRANS_NOTIMPL        -2
#define TRANS_MISSING_ARG    -3
#define TRANS_UNRECOGNIZED_ARG -4
#define TRANS_STALE_ARG      -5

struct Trans_t {
  /* input */
  int    ndgl;        /* number of latitudes                                   */
  int*   nloen;       /* points per latitude [ndgl] (owned by the library after trans_set_resol) */
  int    nlon;        /* regular grids: points per latitude                    */
  int    nsmax;       /* spectral truncation                                   */
  int    lsplit;      /* accepted, ignored (latitude bands are never split)    */
  int    llatlon;     /* must be 0                                             */
  int    flt;         /* must be 0 / -1 (no fast Legendre transform)           */
  int    fft;         /* ignored                                               */
  /* parallel */
  int    myproc;      /* 1-based task                                          */
  int    nproc;
  int    handle;      /* resolution tag                                        */
  /* TRANS_INQ results (filled by trans_setup / trans_inquire) */
  int    nspec, nspec2, nspec2g, nspec2mx, nump, ngptot, ngptotg, ngptotmx;
  int*   ngptotl;     /* [nproc]                                               */
  int*   nmyms;       /* [nump]                                                */
  int*   nasm0;       /* [nsmax+1], 1-based offsets as in the reference, -99 when not local */
  int    nprtrw;
  int*   numpp;       /* [nprtrw]                                              */
  int*   nallms;      /* [nsmax+1]                                             */
  int*   nptrms;      /* [nprtrw] 1-based                                      */
  int*   nvalue;      /* [nspec2] total wavenumber n of each coefficient       */
  int*   nultpp;      /* [nprtrw] latitudes per rank in Fourier space          */
  int*   nptrls;      /* [nprtrw] 1-based first latitude per rank              */
  int*   nnmeng;      /* [ndgl] cut-off zonal wavenumber per latitude          */
  double* rmu;        /* [ndgl]                                                */
  double* rgw;        /* [ndgl]                                                */
};

struct InvTrans_t {
  const double* rspscalar; const double* rspvor; const double* rspdiv;
  const double* rmeanu; const double* rmeanv;     /* LAM only: must be NULL */
  double* rgp;
  int nproma, nscalar, nvordiv, lscalarders, luvder_EW, lvordivgp, ngpblks, lglobal;
  struct Trans_t* trans;
  int count;
};

struct DirTrans_t {
  const double* rgp;
  double* rspscalar; double* rspvor; double* rspdiv;
  const double* rmeanu; const double* rmeanv;
  int nproma, nscalar, nvordiv, ngpblks, lglobal;
  struct Trans_t* trans;
  int count;
};

struct DirTransAdj_t {    /* transi.h:945-972: rgp is the OUTPUT of the adjoint, the spectral arrays its input */
  double* rgp;
  const double* rspscalar; const double* rspvor; const double* rspdiv;
  const double* rmeanu; const double* rmeanv;
  int nproma, nscalar, nvordiv, ngpblks, lglobal;
  struct Trans_t* trans;
  int count;
};

struct InvTransAdj_t {    /* transi.h:1032-1066: rgp is the INPUT, results are added to the spectral arrays */
  double* rspscalar; double* rspvor; double* rspdiv;
  const double* rmeanu; const double* rmeanv;
  const double* rgp;
  int nproma, nscalar, nvordiv, lscalarders, luvder_EW, lvordivgp, ngpblks, lglobal;
  struct Trans_t* trans;
  int count;
};

struct DistGrid_t { const double* rgpg; double* rgp; const int* nfrom; int nproma, nfld, ngpblks; struct Trans_t* trans; int count; };
struct GathGrid_t { double* rgpg; const double* rgp; const int* nto; int nproma, nfld, ngpblks; struct Trans_t* trans; int count; };
struct DistSpec_t { const double* rspecg; double* rspec; const int* nfrom; int nfld; struct Trans_t* trans; int count; };
struct GathSpec_t { double* rspecg; const double* rspec; const int* nto; int nfld; struct Trans_t* trans; int count; };

/* transi.h:1193-1217; handle-free: a spectral-only resolution of truncation nsmax, ncoeff = (nsmax+1)(nsmax+2) */
struct VorDivToUV_t {
  const double* rspvor;  /* [ncoeff][nfld] */
  const double* rspdiv;
  double* rspu;          /* U cos(theta)   */
  double* rspv;
  int nfld;
  int nsmax;
  int ncoeff;
  int count;
};

struct SpecNorm_t {
  const double* rspec;   /* [nspec2][nfld] */
  int nmaster;
  const double* rmet;    /* metric (0:nsmax), optional */
  double* rnorm;         /* [nfld]         */
  int nfld;
  struct Trans_t* trans;
  int count;
};

const char* trans_error_msg(int errcode);
int trans_use_mpi(int);                 /* 0: serial; 1 is refused (multi-GPU goes through ect_setup + NCCL) */
int trans_init(void);
int trans_new(struct Trans_t*);
int trans_set_resol(struct Trans_t*, int ndgl, const int* nloen);
int trans_set_trunc(struct Trans_t*, int nsmax);
int trans_setup(struct Trans_t*);
int trans_inquire(struct Trans_t*, const char* varlist);
struct InvTrans_t new_invtrans(struct Trans_t*);
int trans_invtrans(struct InvTrans_t*);
struct DirTrans_t new_dirtrans(struct Trans_t*);
int trans_dirtrans(struct DirTrans_t*);
struct DirTransAdj_t new_dirtrans_adj(struct Trans_t*);
int trans_dirtrans_adj(struct DirTransAdj_t*);
struct InvTransAdj_t new_invtrans_adj(struct Trans_t*);
int trans_invtrans_adj(struct InvTransAdj_t*);
/* nfrom / nto: 1-based tasks as in the reference */
struct DistGrid_t new_distgrid(struct Trans_t*);
int trans_distgrid(struct DistGrid_t*);
struct GathGrid_t new_gathgrid(struct Trans_t*);
int trans_gathgrid(struct GathGrid_t*);
struct DistSpec_t new_distspec(struct Trans_t*);
int trans_distspec(struct DistSpec_t*);
struct GathSpec_t new_gathspec(struct Trans_t*);
int trans_gathspec(struct GathSpec_t*);
struct VorDivToUV_t new_vordiv_to_UV(void);
int trans_vordiv_to_UV(struct VorDivToUV_t*);
struct SpecNorm_t new_specnorm(struct Trans_t*);
int trans_specnorm(struct SpecNorm_t*);
int trans_delete(struct Trans_t*);
int trans_finalize(void);

#ifdef __cplusplus
}
#endif
#endif
