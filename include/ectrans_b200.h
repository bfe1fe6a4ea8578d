/*
 * ectrans_b200.h -- C ABI of the B200-native global spectral transform.
 *
 * This is the drop-in boundary underneath the three faces of the reference
 * (SURVEY.md 8(b)).  Every entry point cites the reference interface it replaces
 * (paths relative to /root/reference):
 *
 *   ect_setup      <- SETUP_TRANS0 + SETUP_TRANS   src/trans/include/ectrans/setup_trans0.h:12-89,
 *                                                  setup_trans.h:12-115; transi trans_setup() src/transi/transi.h
 *   ect_inquire*   <- TRANS_INQ                    src/trans/include/ectrans/trans_inq.h:12-21; transi trans_inquire()
 *   ect_inv_trans  <- INV_TRANS                    src/trans/include/ectrans/inv_trans.h:12-161; transi trans_invtrans()
 *   ect_dir_trans  <- DIR_TRANS                    src/trans/include/ectrans/dir_trans.h:12-140; transi trans_dirtrans()
 *   ect_specnorm   <- SPECNORM                     src/trans/include/ectrans/specnorm.h
 *   ect_inv_transad / ect_dir_transad, ect_gath_* / ect_dist_*, ect_gpnorm_trans, ect_vordiv_to_uv, ect_inquire_rpnm,
 *   ect_trans_pnm, ect_write_legpol / ect_read_legpol, ect_gridpoint_partition, ect_*_trans_vset, ect_specnorm_vset:
 *                  the interfaces they replace are cited at their declarations below
 *   ect_release    <- TRANS_RELEASE / trans_delete src/trans/include/ectrans/trans_release.h
 *   ect_finalize   <- TRANS_END / trans_finalize   src/trans/include/ectrans/trans_end.h
 *
 * Conventions
 *   - plain C types only; all arrays are raw pointers with explicit extents.
 *   - memory layouts are the reference's Fortran layouts (column major), i.e. for the C
 *     reader: spectral arrays are [nspec2][nfld] (field fastest), grid-point arrays are
 *     [ngpblks][nfld][nproma] (point-in-block fastest).  These equal transi's
 *     rspscalar[nspec2][nscalar] and rgp[ngpblks][nfld][nproma] (transi.h:986-1010).
 *   - memspace ECT_MEM_HOST: the call copies inputs host->device and results
 *     device->host inside the call, like the reference GPU backend
 *     (gpu/internal/ltinv_mod.F90:333-339, trltog_mod.F90:950-990).
 *     ECT_MEM_DEVICE: pointers are device pointers, no copies.
 *   - every function returns ECT_SUCCESS (0) or a negative code; nothing aborts.
 *     The codes -1..-5 equal transi's (transi.c:33-58).
 *   - one handle = one resolution on one GPU/rank; calls on a handle are serialised by
 *     the caller (same contract as the reference: SET_RESOL global state, not re-entrant).
 *   - there is no CPU fallback: without a CUDA device ect_setup() fails with
 *     ECT_ERR_CUDA unless ECT_SETUP_HOST_ONLY is given (plan/inquire only).
 */
#ifndef ECTRANS_B200_H
#define ECTRANS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define ECT_SUCCESS        0
#define ECT_ERR_GENERIC   -1
#define ECT_ERR_NOTIMPL   -2
#define ECT_ERR_MISSING   -3
#define ECT_ERR_BADARG    -4
#define ECT_ERR_STALE     -5
#define ECT_ERR_CUDA      -6
#define ECT_ERR_NCCL      -7
#define ECT_ERR_HANDLE    -8

#define ECT_MEM_HOST   0
#define ECT_MEM_DEVICE 1

#define ECT_SETUP_HOST_ONLY 1   /* build geometry + decomposition only, no CUDA (inquire works) */
#define ECT_SETUP_STREAM_GIVEN 2 /* opts.stream is valid even if it is 0 (the legacy default stream) */
#define ECT_SETUP_GP_EQ_REGIONS 8 /* grid-point arrays follow the reference's default decomposition (LDEQ_REGIONS=T, LDSPLIT=T:
                                     eq_regions bands x regions, split latitudes) instead of this library's native one (= the
                                     Fourier latitude bands, TRLTOG / TRGTOL local); TRLTOG / TRGTOL then are NCCL all-to-alls */
#define ECT_SETUP_BANDS_BY_POINTS 16 /* Fourier latitude bands balanced by grid points exactly as SUMPLATB does (sumplatb_mod.F90:171-216);
                                     default with more than one rank: the same algorithm on a cost weight per latitude (FFT work of
                                     the row + a per-row constant, csrc/host_plan.cu), which balances the Fourier kernels */
#define ECT_SETUP_LEGPOL_DEFER 4 /* do not compute the Legendre table: ect_read_legpol() fills it (CDIO_LEGPOL='readf') */

#define ECT_NCCL_UID_BYTES 128

/* Precision of the caller's arrays (the reference builds one library per precision, src/trans/CMakeLists.txt:9-39).
   ECT_PREC_SP: every field array in ect_inv_args / ect_dir_args / ect_specnorm is float although the members are
   typed double*.  The Fourier stage runs in float; the Legendre contraction runs as 3xTF32 on the tcgen05 tensor
   cores with fp32 accumulation (csrc/legendre_tc.cu; the reference's GPU sp build: hicblas_cutlass.cuda.h:41-72), m = 0
   in double precision as in the reference (ledir_mod.F90:133-171); ECT_SP_TC=0 runs every m on the FP64 kernels. */
#define ECT_PREC_DP 0
#define ECT_PREC_SP 1

typedef struct ect_setup_opts {
    int nsmax;            /* KSMAX: spectral truncation                                  */
    int ndgl;             /* KDGL : number of Gaussian latitudes (even)                  */
    const int* nloen;     /* KLOEN(ndgl): points per latitude, north -> south            */
    int nranks;           /* number of tasks NPROC = NPRTRW * NPRTRV (NPRTRV = 1 unless ECT_SETUP_NPRTRV is in flags); 1 = LDMPOFF */
    int rank;             /* 0-based task id (MYPROC-1); W-set = rank / NPRTRV, V-set = rank % NPRTRV */
    int flags;            /* ECT_SETUP_*                                                 */
    int device;           /* CUDA device ordinal, -1 = current                           */
    void* stream;         /* cudaStream_t to run on, NULL = library-owned stream         */
    const void* nccl_uid; /* ECT_NCCL_UID_BYTES from ect_nccl_unique_id() of rank 0; NULL if nranks == 1 */
    int precision;        /* ECT_PREC_DP (JPRB = real64, trans_*_dp) or ECT_PREC_SP (JPRB = real32, trans_*_sp) */
} ect_setup_opts;

typedef struct ect_info {
    int nsmax, ndgl, ndgnh;
    int nranks, rank;
    int nspec2, nspec2g;     /* local / global number of real spectral coefficients (2 per (m,n)) */
    int ngptot, ngptotg;     /* local / global number of grid points                              */
    int nump;                /* number of local zonal wavenumbers                                 */
    int lat0, nlat;          /* local latitude band (0-based first, count) in Fourier/grid space  */
    long long table_bytes;   /* Legendre table bytes resident in HBM                              */
} ect_info;

/* arrays for ect_inquire_array() */
#define ECT_ARR_NLOEN   1   /* int[ndgl]                                                    */
#define ECT_ARR_NMEN    2   /* int[ndgl]     G%NMEN                                          */
#define ECT_ARR_NDGLU   3   /* int[nsmax+1]  G%NDGLU                                         */
#define ECT_ARR_MYMS    4   /* int[nump]     D%MYMS                                          */
#define ECT_ARR_NASM0   5   /* int[nsmax+1]  D%NASM0, 0-based offsets, -1 = not local        */
#define ECT_ARR_NPROCM  6   /* int[nsmax+1]  D%NPROCM, 0-based rank                          */
#define ECT_ARR_RMU     7   /* double[ndgl]  F%RMU                                           */
#define ECT_ARR_RGW     8   /* double[ndgl]  F%RW                                            */
#define ECT_ARR_LATFIRST 9  /* int[nranks]   first latitude of each rank's band              */
#define ECT_ARR_LATCOUNT 10 /* int[nranks]                                                   */
#define ECT_ARR_SENDCNT 11  /* long long[nranks] TRMTOL send counts in (lat,m) records       */
#define ECT_ARR_RECVCNT 12  /* long long[nranks]                                             */
#define ECT_ARR_RACTHE  13  /* double[ndgl]  F%RACTHE                                        */
/* Fourier-buffer record tables (what TRMTOL/TRLTOM move; NLTSGTB/NSTAGT0B/NPNTGTB1 in the reference) */
#define ECT_ARR_MROW0   14  /* long long[nump+1]  start of local m in LEGRECN/LEGRECS          */
#define ECT_ARR_LEGRECN 15  /* int[mrow0[nump]]   record of (local m, northern latitude i)     */
#define ECT_ARR_LEGRECS 16  /* int[mrow0[nump]]   record of its southern mirror latitude       */
#define ECT_ARR_LATROW0 17  /* long long[nlat+1]  start of local latitude in FFTREC            */
#define ECT_ARR_FFTREC  18  /* int[latrow0[nlat]] record of (local latitude, m), -1 if m > NMEN */
#define ECT_ARR_SENDOFF 19  /* long long[nranks]                                              */
#define ECT_ARR_RECVOFF 20  /* long long[nranks]                                              */
/* fused (peer-memory) transposition: destination rank / record of what this rank produces */
#define ECT_ARR_LEGDSTRANKN 21 /* int[mrow0[nump]]   rank owning the northern latitude                 */
#define ECT_ARR_LEGDSTRECN  22 /* int[mrow0[nump]]   record in that rank's Fourier-side buffer          */
#define ECT_ARR_LEGDSTRANKS 23
#define ECT_ARR_LEGDSTRECS  24
#define ECT_ARR_FFTDSTRANK  25 /* int[latrow0[nlat]] rank owning m                                      */
#define ECT_ARR_FFTDSTREC   26 /* int[latrow0[nlat]] record in that rank's Legendre-side buffer         */
/* grid-point partition of the caller's arrays (TRANS_INQ KSTA / KONL / N_REGIONS; 0-based) */
#define ECT_ARR_GPSEGS      27 /* int[3 * npieces]: (latitude, first point, count) of this task's pieces, local order;
                                  call with capacity 0 ... the count is ect_inquire_array(..., ECT_ARR_NGPSEGS) */
#define ECT_ARR_NGPSEGS     28 /* int[1]  number of pieces                                            */
/* TRLTOG / TRGTOL message tables of the eq_regions partition (test access): the message between band owner r and
   task p carries the band points XBIDX[XBOFF[p] .. XBOFF[p+1]) of r = the local points [XGOFF[r], XGOFF[r+1]) of p */
#define ECT_ARR_XBIDX       30 /* int[points of the latitude band]  */
#define ECT_ARR_XBOFF       31 /* long long[nranks + 1]             */
#define ECT_ARR_XGOFF       32 /* long long[nranks + 1]             */
#define ECT_ARR_NREGIONS    29 /* int[nranks] N_REGIONS(band), zero padded (ECT_SETUP_GP_EQ_REGIONS only) */

typedef struct ect_inv_args {
    int memspace;                 /* ECT_MEM_HOST / ECT_MEM_DEVICE                          */
    int nproma;                   /* KPROMA, <= 0 means ngptot (one block)                  */
    int scders, vorgp, divgp, uvder;   /* LDSCDERS, LDVORGP, LDDIVGP, LDUVDER               */
    /* spectral input, call mode 1 */
    const double* spvor;          /* PSPVOR(nuv, nspec2)                                    */
    const double* spdiv;          /* PSPDIV(nuv, nspec2)                                    */
    int nuv;
    const double* spscalar;       /* PSPSCALAR(nscalar, nspec2)                             */
    int nscalar;
    /* spectral input, call mode 2 (used when spscalar == NULL) */
    const double* spsc2;  int nsc2;                         /* PSPSC2(nsc2, nspec2)          */
    const double* spsc3a; int nsc3a_lev, nsc3a_fld;         /* PSPSC3A(lev, nspec2, fld)     */
    const double* spsc3b; int nsc3b_lev, nsc3b_fld;         /* PSPSC3B(lev, nspec2, fld)     */
    /* grid-point output, call mode 1: PGP(nproma, nfld_gp, ngpblks), field order
       [vor][div] u v scalars [N-S ders] [E-W du dv] [E-W ders]  (inv_trans.h:66-76)        */
    double* gp;
    /* grid-point output, call mode 2 (used when gp == NULL) */
    double* gpuv;                 /* PGPUV(nproma, nuv, nvar, ngpblks), var: [vor][div] u v [du dv] */
    double* gp2;                  /* PGP2 (nproma, nsc2*(1|3), ngpblks)                      */
    double* gp3a;                 /* PGP3A(nproma, lev, fld*(1|3), ngpblks)                  */
    double* gp3b;
} ect_inv_args;

typedef struct ect_dir_args {
    int memspace;
    int nproma;
    int nuv, nscalar;             /* counts for call mode 1                                  */
    /* grid-point input, call mode 1: PGP(nproma, 2*nuv+nscalar, ngpblks), order u v scalars (dir_trans.h:59-61) */
    const double* gp;
    /* call mode 2 */
    const double* gpuv;           /* PGPUV(nproma, nuv, 2, ngpblks): u, v                    */
    const double* gp2;  int nsc2;
    const double* gp3a; int nsc3a_lev, nsc3a_fld;
    const double* gp3b; int nsc3b_lev, nsc3b_fld;
    /* spectral output */
    double* spvor; double* spdiv; double* spscalar;
    double* spsc2; double* spsc3a; double* spsc3b;
} ect_dir_args;

/* timings of the last call on this handle, milliseconds, CUDA events on the handle's stream */
typedef struct ect_timings {
    float h2d, prologue, legendre, transpose, fourier, epilogue, d2h, total;
    long long launches;     /* kernels launched by the last call */
} ect_timings;

int ect_setup(const ect_setup_opts* opts, int* handle);
int ect_inquire(int handle, ect_info* info);
int ect_inquire_array(int handle, int which, void* out, long long capacity_elems);
int ect_inv_trans(int handle, const ect_inv_args* args);
int ect_dir_trans(int handle, const ect_dir_args* args);
int ect_specnorm(int handle, const double* spec, int nfld, int memspace, double* norms /* host, nfld */);
/* SPECNORM with the optional metric PMET(0:NSMAX) (host, may be NULL).  The n-sums of every wavenumber and then the
 * wavenumbers are added in the reference's order (spnormd_mod.F90:36-51, spnorm_ctl_mod.F90:56-57): the norms are bit
 * identical across decompositions. */
int ect_specnorm_met(int handle, const double* spec, int nfld, int memspace, const double* pmet, double* norms);
/* SPECNORM with KVSET (NPRTRV > 1, specnorm.h): spec holds the nfld fields of this task's V-set, kvset[nfld_g] the 1-based
 * V-set of every global field; norms[nfld_g] arrive on every task. */
int ect_specnorm_vset(int handle, const double* spec, int nfld, int memspace, const double* pmet, const int* kvset,
                      int nfld_g, double* norms);
int ect_get_timings(int handle, ect_timings* t);
/* INV_TRANSAD / DIR_TRANSAD: adjoints for the inner products of the reference's adjoint tests (grid: plain sum;
 * spectral: weight 2 for m > 0, 1 for the real parts of m = 0).  Replace src/trans/include/ectrans/inv_transad.h,
 * dir_transad.h.  Same argument structs as the forward calls with the roles of the arrays exchanged:
 *   ect_inv_transad: gp* arrays are INPUT (u, v, scalars; no scders / vorgp / divgp / uvder -> ECT_ERR_NOTIMPL),
 *                    the results are ADDED to the spectral arrays (as the reference does);
 *   ect_dir_transad: spectral arrays are INPUT (left untouched; the reference zeroes them), gp* arrays OUTPUT. */
int ect_inv_transad(int handle, const ect_inv_args* args);
int ect_dir_transad(int handle, const ect_dir_args* args);

/* GATH_GRID / DIST_GRID / GATH_SPEC / DIST_SPEC: host arrays, every rank of the handle calls (collective).
 * Replaces src/trans/include/ectrans/gath_grid.h:12-60, dist_grid.h:12-68, gath_spec.h:12-70, dist_spec.h:12-71.
 *   gp_local  PGP(nproma, nfld, ngpblks)       gp_global PGPG(ngptotg, nfld_owned)
 *   sp_local  PSPEC(nfld, nspec2)              sp_global PSPECG(nfld_owned, nspec2g), m ascending then n ascending
 * kto[f] / kfrom[f] = rank (0-based, the reference's KTO/KFROM minus 1) that holds field f of the global array; a
 * rank's global array carries the fields it holds, in order (may be NULL on ranks that hold none).
 * GATH_SPEC zeroes the imaginary parts of the zonal (m = 0) coefficients (LDZA0IP default). */
int ect_gath_grid(int handle, const void* gp_local, int nfld, int nproma, const int* kto, void* gp_global);
int ect_dist_grid(int handle, const void* gp_global, int nfld, int nproma, const int* kfrom, void* gp_local);
int ect_gath_spec(int handle, const void* sp_local, int nfld, const int* kto, void* sp_global);
int ect_dist_spec(int handle, const void* sp_global, int nfld, const int* kfrom, void* sp_local);

/* GPNORM_TRANS (src/trans/include/ectrans/gpnorm_trans.h:12-47): global average (Gaussian weight / points of the
 * latitude, latitudes added in global order), minimum and maximum of nfld grid-point fields PGP(nproma, nfld,
 * ngpblks).  ave_only != 0 (LDAVE_ONLY): pmin / pmax hold the task-local extrema on entry.  Collective; the results
 * arrive on every rank (the reference defines them on the first task only).  ave, pmin, pmax: host, double[nfld]. */
int ect_gpnorm_trans(int handle, const void* gp, int nfld, int nproma, int memspace, double* ave, double* pmin,
                     double* pmax, int ave_only);
/* VORDIV_TO_UV (src/trans/include/ectrans/vordiv_to_uv.h; transi trans_vordiv_to_UV, transi.h:1193-1217): spectral
 * vorticity / divergence (nfld, nspec2) -> spectral U cos(theta), V cos(theta), rows n <= nsmax.  handle > 0: that
 * handle's wavenumbers, precision and stream; handle == 0: one task holding every m (nspec2 = (nsmax+1)(nsmax+2)),
 * double precision -- the reference builds a temporary spectral-only resolution for this call. */
int ect_vordiv_to_uv(int handle, int nsmax, const void* spvor, const void* spdiv, void* spu, void* spv, int nfld,
                     int memspace);
/* Legendre polynomials in the reference's layouts, read back from the table in HBM.
 *   ect_inquire_rpnm <- TRANS_INQ(PRPNM, KSPOLEGL, KPMS) src/trans/include/ectrans/trans_inq.h:93-101: rpnm is
 *       PRPNM(ndgnh, nspolegl) column major (may be NULL to query the sizes), npms[m] = NPMS(m) or -1;
 *   ect_trans_pnm    <- TRANS_PNM src/trans/include/ectrans/trans_pnm.h: one wavenumber, PRPNM(ld, ncols),
 *       ld >= ndgnh, ncols >= nsmax - m + 2; column p (1-based) holds n = nsmax + 2 - p. */
int ect_inquire_rpnm(int handle, double* rpnm, long long capacity_elems, int* nspolegl, int* npms);
int ect_trans_pnm(int handle, int m, double* rpnm, int ld, int ncols);

/* Legendre polynomial cache in the reference's file format (SETUP_TRANS CDIO_LEGPOL='writef' / 'readf' with
 * CDLEGPOLFNAME, src/trans/include/ectrans/setup_trans.h:60-66; write_legpol_mod.F90, read_legpol_mod.F90).  The
 * reader checks label, truncation, latitude count, NLOEN and NMEN like the reference and fails with ECT_ERR_BADARG
 * ("READ_LEGPOL: WRONG ..." in ect_last_error()).  One file per task: it holds the task's wavenumbers (MYMS). */
int ect_write_legpol(int handle, const char* path);
int ect_read_legpol(int handle, const char* path);

/* The reference's default grid-point decomposition (SETUP_TRANS0 LDEQ_REGIONS=T, SETUP_TRANS LDSPLIT=T:
 * eq_regions_mod.F90, sumplatbeq_mod.F90, sumplat_mod.F90, sustaonl_mod.F90, pe2set_mod.F90) for nproc tasks; host only,
 * no handle.  regions: int[nproc], the first *nbands entries are N_REGIONS(band); task (0-based) = regions of the bands
 * before + region.  seg0: int[nproc + 1]; segs: int[3 * capacity_segs] = (latitude, first point = NSTA - 1, NONL) of
 * the pieces of every task in its local point order; *nsegs = number of pieces (call with segs = NULL to size). */
int ect_gridpoint_partition(int ndgl, const int* nloen, int nproc, int* nbands, int* regions, int* seg0, int* segs,
                            long long capacity_segs, long long* nsegs);

/* V-sets (NPRTRV > 1).  ect_setup with flags |= ECT_SETUP_NPRTRV(V): opts.nranks is then the total number of tasks
 * NPROC = NPRTRW * V and task pe (0-based) is W-set pe / V, V-set pe % V (pe2set_mod.F90:81-82).  Spectral arrays hold the
 * levels / fields of the task's V-set, the grid-point arrays ALL of them on the task's eq_regions points (as in the
 * reference, inv_trans.h:36-58, :100-160); the kvset* arrays (1-based V-set per GLOBAL level / field) say which is which.
 * In ect_inv_args / ect_dir_args the spectral counts (nuv, nscalar, nsc2, nsc3a_lev, nsc3b_lev) are the LOCAL ones;
 * nsc3a_fld / nsc3b_fld are not distributed.  Calls are synchronous. */
#define ECT_SETUP_NPRTRV(v) (((v) & 0xff) << 8)
typedef struct ect_vset_args {
    const int* kvsetuv;   int nuv_g;         /* KVSETUV(:)   vor / div / u / v levels                   */
    const int* kvsetsc;   int nscalar_g;     /* KVSETSC(:)   scalars of call mode 1                      */
    const int* kvsetsc2;  int nsc2_g;        /* KVSETSC2(:)                                              */
    const int* kvsetsc3a; int nsc3a_lev_g;   /* KVSETSC3A(:) levels (every 3-D field of PSPSC3A alike)   */
    const int* kvsetsc3b; int nsc3b_lev_g;
} ect_vset_args;
int ect_inv_trans_vset(int handle, const ect_inv_args* args, const ect_vset_args* vs);
int ect_dir_trans_vset(int handle, const ect_dir_args* args, const ect_vset_args* vs);

int ect_synchronize(int handle);   /* wait for asynchronous ECT_MEM_DEVICE calls on this handle */
/* How TRMTOL / TRLTOM (trmtol_mod.F90:101-141) run on this handle: *peer_memory = 1 when the producing kernels store
 * straight into the consumer rank's buffer over NVLink (decided collectively at setup: one host, one process per GPU,
 * peer access between every pair; otherwise 0 = NCCL all-to-all-v); *entry_barriers = consumer-done barriers issued so
 * far (one before every transform that follows a transform of the same direction). */
int ect_comm_info(int handle, int* peer_memory, long long* entry_barriers);
int ect_release(int handle);
int ect_finalize(void);
const char* ect_strerror(int code);
const char* ect_last_error(void);

/* multi-GPU plumbing: rank 0 creates the id, the caller broadcasts it (MPI_Bcast in a
   Fortran/MPI host, torch.distributed in the Python harness) and passes it to ect_setup. */
int ect_nccl_unique_id(void* out_bytes /* ECT_NCCL_UID_BYTES */);

/* pinned host allocation for callers that want fast H2D/D2H (the reference benchmark's
   ectrans_memory.c / --no-pinning switch, src/programs/util/ectrans_memory.c) */
int ect_host_alloc(void** ptr, long long bytes);
int ect_host_free(void* ptr);

/* test access: copies the Legendre table of local wavenumber index ml (par 0: n-m even, 1: odd),
   laid out [k][ndglu], to host memory */
int ect_debug_get_table(int handle, int ml, int par, double* out, long long capacity_elems);

/* FP64 peak microbenchmarks used for the roofline denominators (not part of the reference API) */
int ect_measure_fp64_peak(int which /*0 DMMA m8n8k4, 1 DFMA, 2 DMMA and DFMA warps mixed, 3 mma.sync TF32 m16n8k8, 4 FFMA*/, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
