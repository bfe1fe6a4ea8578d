/*
 * transi_b200.h -- the transi C face of the reference (src/transi/transi.h) on top of the B200 library.
 *
 * Subset: the functions and struct members the hot path needs (transi.h:110-233, 293, 422, 667-685;
 * structs transi.h:701-830, 905-921, 986-1010, 1190-1200).  Same names, argument meaning, layouts
 * (rspscalar[nspec2][nscalar], rgp[ngpblks][nfld][nproma], lglobal: rgp[nfld][ngptotg]) and the same
 * return codes (TRANS_SUCCESS 0, -1 error, -2 not implemented, -3 missing arg, -4 unrecognised arg,
 * -5 stale arg; transi.c:33-58).  Argument structs are single use (count guard,
 * transi_module.F90:1949-1954).  Also trans_dirtrans_adj / trans_invtrans_adj (transi.h:354, 491) and
 * trans_distgrid / gathgrid / distspec / gathspec (transi.h:520-616), trans_vordiv_to_UV, trans_specnorm, the setters
 * trans_set_* (transi.h:121-194) and every member of struct Trans_t (transi.h:701-850).  Out of scope: LAM, lon-lat
 * grids, the in-memory cache (those calls exist and return TRANS_NOTIMPL).  The reference's own test program
 * tests/transi/transi_test_program.c compiles against this header unchanged (include/ectrans/transi.h forwards here).
 * A caller of the reference includes this header instead of "ectrans/transi.h" and links
 * libectrans_b200.so instead of libtransi_dp.so.
 */
#ifndef TRANSI_B200_H
#define TRANSI_B200_H
#ifndef ectrans_transi_h
#define ectrans_transi_h        /* the reference's include guard: "ectrans/transi.h" and this header exclude each other */
#endif
#include <stddef.h>

typedef int _bool;              /* transi.h:81 */

#ifdef __cplusplus
extern "C" {
#endif

/* src/transi/version.h:21-29 */
const char* ectrans_version(void);
unsigned int ectrans_version_int(void);
const char* ectrans_version_str(void);
const char* ectrans_git_sha1(void);
const char* ectrans_git_sha1_abbrev(unsigned int length);

#define TRANS_FFT992 1
#define TRANS_FFTW   2

#define TRANS_SUCCESS         0
#define TRANS_ERROR          -1
#define TRANS_NOTIMPL        -2
#define TRANS_MISSING_ARG    -3
#define TRANS_UNRECOGNIZED_ARG -4
#define TRANS_STALE_ARG      -5

/* Member for member the reference's struct (transi.h:701-850).  Scalars are defined by trans_setup(), arrays are
 * allocated and filled by trans_inquire() and freed by trans_delete(). */
struct Trans_t {
  /* INPUT */
  int    ndgl;        /* number of latitudes                                                     */
  int*   nloen;       /* points per latitude [ndgl]                                              */
  int    nlon;        /* regular grids: points per latitude                                      */
  int    nsmax;       /* spectral truncation                                                     */
  _bool  llam;        /* LAM resolution: not supported (trans_setup returns TRANS_NOTIMPL)        */
  _bool  lsplit;      /* latitudes may be split between grid-point tasks (one task: irrelevant)  */
  int    llatlon;     /* must be 0                                                               */
  int    flt;         /* must be 0 / -1 (no fast Legendre transform)                             */
  int    fft;         /* TRANS_FFT992 / TRANS_FFTW: accepted, ignored (hand-written FFT kernels) */
  char*  readfp;      /* trans_set_read : Legendre polynomials are read from this file (reference format) */
  char*  writefp;     /* trans_set_write: ... and written to this one                            */
  const void* cache;  /* trans_set_cache: not supported                                          */
  size_t cachesize;
  /* PARALLELISATION */
  int    myproc;      /* 1-based task                                                            */
  int    nproc;
  /* MULTI-TRANSFORMS-MANAGEMENT */
  int    handle;      /* resolution tag                                                          */
  /* SPECTRAL SPACE */
  int    nspec, nspec2, nspec2g, nspec2mx, nump, ngptot, ngptotg, ngptotmx;
  int*   ngptotl;     /* [n_regions_NS][n_regions_EW]                                            */
  int*   nmyms;       /* [nump]                                                                  */
  int*   nasm0;       /* [nsmax+1], 1-based offsets as in the reference, -99 when not local       */
  int    nprtrw;
  int*   numpp;       /* [nprtrw]                                                                */
  int*   npossp;      /* [nprtrw+1] 1-based start of each W-set in the global spectral array      */
  int*   nptrms;      /* [nprtrw] 1-based                                                        */
  int*   nallms;      /* [nsmax+1]                                                               */
  int*   ndim0g;      /* [nsmax+1] 1-based start of wavenumber m in the global spectral array     */
  int*   nvalue;      /* [nspec2] total wavenumber n of each coefficient                         */
  /* GRIDPOINT SPACE */
  int    n_regions_NS, n_regions_EW, my_region_NS, my_region_EW;
  int*   n_regions;   /* [n_regions_NS]                                                          */
  int*   nfrstlat;    /* [n_regions_NS] 1-based                                                  */
  int*   nlstlat;     /* [n_regions_NS]                                                          */
  int    nfrstloff;
  int*   nptrlat;     /* [ndgl]                                                                  */
  int*   nptrfrstlat; /* [n_regions_NS]                                                          */
  int*   nptrlstlat;  /* [n_regions_NS]                                                          */
  int    nptrfloff;
  int*   nsta;        /* [n_regions_EW][ndgl+n_regions_NS-1]                                     */
  int*   nonl;        /* [n_regions_EW][ndgl+n_regions_NS-1]                                     */
  _bool* ldsplitlat;  /* [ndgl]                                                                  */
  /* FOURIER SPACE */
  int    nprtrns;
  int*   nultpp;      /* [nprtrns] latitudes per rank in Fourier space                           */
  int*   nptrls;      /* [nprtrns] 1-based first latitude per rank                               */
  int*   nnmeng;      /* [ndgl] cut-off zonal wavenumber per latitude                            */
  /* LEGENDRE */
  double* rmu;        /* [ndgl]                                                                  */
  double* rgw;        /* [ndgl]                                                                  */
  double* rpnm;       /* [nspolegl][nlei3] (Fortran RPNM(nlei3, nspolegl))                        */
  int     nlei3;
  int     nspolegl;
  int*    npms;       /* [nsmax+1]                                                               */
  double* rlapin;     /* [nsmax+4] = RLAPIN(-1:nsmax+2)                                          */
  int*    ndglu;      /* [nsmax+1]                                                               */
  /* LAM (never used here; members exist for source compatibility) */
  double  pexwn, peywn;
  double* pweight;
  int     ndgux;
  int     nmsmax;
  int*    mvalue;
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
/* transi.h:121-194.  Values the hot path cannot honour are refused with TRANS_NOTIMPL instead of being ignored. */
int trans_set_handles_limit(int limit);
int trans_set_radius(double radius);           /* 6371229 m only (setup_trans0.F90:129 default; tables are built for it) */
int trans_set_nprtrv(int nprtrv);              /* 1 only on this face (V-sets: ect_inv_trans_vset)                      */
int trans_set_nprgpew(int nprgpew);            /* 1 only                                                                */
int trans_set_leq_regions(_bool ldeq_regions); /* accepted (one task: no effect)                                        */
int trans_use_mpi(_bool);               /* 0: serial; 1 is refused (multi-GPU goes through ect_setup + NCCL) */
int trans_init(void);
int trans_set_read(struct Trans_t*, const char* filepath);
int trans_set_write(struct Trans_t*, const char* filepath);
int trans_set_cache(struct Trans_t*, const void*, size_t);
int trans_new(struct Trans_t*);
int trans_set_resol(struct Trans_t*, int ndgl, const int* nloen);
int trans_set_resol_lonlat(struct Trans_t*, int nlon, int nlat);               /* TRANS_NOTIMPL at trans_setup (LDLL out of scope) */
int trans_set_resol_lam(struct Trans_t*, int nx, int ny, double dx, double dy); /* TRANS_NOTIMPL at trans_setup (LAM out of scope)  */
int trans_set_trunc(struct Trans_t*, int nsmax);
int trans_set_trunc_lam(struct Trans_t*, int trunc_x, int trunc_y);
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
