// Internal handle and plan structures of the B200 spectral-transform library.
#pragma once
#include <vector>
#include <string>
#include <cstdint>
#include <cuda_runtime.h>
#include "fft_plan.h"
#include "../../include/ectrans_b200.h"

#define ECT_RA 6371229.0          // common/external/setup_trans0.F90:129
#define ECT_LAT_PAD 64            // latitude pitch of the Legendre tables (multiple of the GEMM tile)
#define ECT_CPAD 16               // record pitch granularity (doubles)

typedef long long i64;

// ---------------------------------------------------------------------------------------
// Host plan: geometry + decomposition (no CUDA).  Mirrors what SETUP_TRANS leaves in the
// reference's R/G/D/F module globals (common/internal/tpm_dim.F90, tpm_geometry.F90,
// tpm_distr.F90, tpm_fields.F90).
// ---------------------------------------------------------------------------------------
struct EctGpSeg { int lat, first, count; };       // piece of a latitude: 0-based latitude, first point on it, number of points

struct EctHostPlan {
    int nsmax = 0, ndgl = 0, ndgnh = 0;
    int nranks = 1, rank = 0;
    std::vector<int> nloen, nmen, ndglu;
    std::vector<double> rmu, rw, r1mu2, racthe;
    // spectral space: zonal wavenumbers (SUWAVEDI)
    std::vector<int> nprocm;                 // m -> owning rank
    std::vector<std::vector<int>> ms_of;     // rank -> its m's in local order (MYMS)
    std::vector<int> myms;
    std::vector<int> nasm0;                  // m -> 0-based offset in the local spectral array, -1 if not mine
    int nump = 0, nspec2 = 0, nspec2_g = 0;
    // Fourier / grid-point space: latitude bands (SUMPLATF)
    std::vector<int> lat_first, lat_count;   // per rank
    int lat0 = 0, nlat = 0;
    int band_pad = 0;                        // per-latitude constant of the band cost model (host_plan.cu), 0 when balanced by points
    std::vector<int> gpoff;                  // local latitude -> first local grid point
    int ngptot = 0, ngptotg = 0;
    // Fourier-buffer records.  A record = one (latitude, m) pair, m <= NMEN(lat).
    //   leg side (this rank's m, all latitudes): [dest rank][lat of dest][local m]
    //   fft side (this rank's latitudes, all m): [src rank][local lat][m of src]
    std::vector<i64> mrow0;                  // local m -> start in leg_rec_n/s (length ndglu(m))
    std::vector<int> leg_rec_n, leg_rec_s;   // record of (m, northern lat i) / its southern mirror
    std::vector<i64> latrow0;                // local lat -> start in fft_rec (length nmen+1)
    std::vector<int> fft_rec;
    std::vector<i64> send_cnt, send_off, recv_cnt, recv_off;   // records, per peer (leg -> fft direction)
    // fused transposition: destination rank and record index in the destination's buffer
    std::vector<int> leg_dst_rank_n, leg_dst_rank_s, leg_dst_rec_n, leg_dst_rec_s;   // per (local m, northern lat i)
    std::vector<int> fft_dst_rank, fft_dst_rec;                                      // per (local lat, m)
    i64 nrec_leg = 0, nrec_fft = 0;
    // Grid-point partition of the caller's arrays.  Default: the Fourier latitude bands themselves (TRLTOG / TRGTOL
    // are local).  gp_eq: the reference's eq_regions / LDSPLIT decomposition; then the caller's ngptot points are the
    // pieces gp_segs and TRLTOG / TRGTOL become an all-to-all between band owners and grid-point tasks.
    bool gp_eq = false;
    int ngpband = 0;                         // points of this rank's latitude band (what the Fourier stage works on)
    std::vector<int> gp_regions;             // N_REGIONS(band)
    std::vector<EctGpSeg> gp_segs;           // my pieces, local point order
    std::vector<EctGpSeg> gp_all_segs; std::vector<int> gp_all_seg0;    // every task's pieces (GATH_GRID / DIST_GRID)
    // band owner side: my band points ordered by (grid-point task, its local order); xb_off[p] .. xb_off[p+1] go to task p
    std::vector<int> xb_idx; std::vector<i64> xb_off;
    // grid-point task side: my local points [xg_off[r], xg_off[r+1]) lie on latitudes of band owner r
    std::vector<i64> xg_off;
    std::string err;
};

int ect_build_host_plan(EctHostPlan& P, int nsmax, int ndgl, const int* nloen, int nranks, int rank, bool gp_eq = false,
                        bool bands_by_points = false);

// Grid-point decomposition LDEQ_REGIONS=T, LDSPLIT=T (gp_partition.cu)
struct EctGpPartition {
    std::vector<int> regions;                     // N_REGIONS(band)
    std::vector<int> band_first, band_last;       // NFRSTLAT / NLSTLAT (0-based; a split latitude belongs to both bands)
    std::vector<long long> band_points;           // KPROCAGP
    std::vector<int> seg0;                        // task -> first entry of segs (size ntasks + 1)
    std::vector<EctGpSeg> segs;                   // per task, in its local point order (latitude, then longitude)
};
int ect_eq_regions(int n, std::vector<int>& regions);
int ect_gp_partition(const std::vector<int>& nloen, int nproc, EctGpPartition& G);
void ect_gauss_latitudes(int ndgl, std::vector<double>& mu, std::vector<double>& w);

// ---------------------------------------------------------------------------------------
// Device state
// ---------------------------------------------------------------------------------------
struct EctLegM {           // per local m, device-resident descriptor
    int m;
    int ndglu;             // latitudes carrying m (northern hemisphere)
    int ldp;               // latitude pitch of the table rows (multiple of ECT_LAT_PAD)
    int ils, ila;          // rows: n-m even (symmetric), odd (antisymmetric)
    i64 ps_off, pa_off;    // offsets (doubles) into the polynomial table: P[k][lat]
    i64 xrow0;             // first row of this m in the spectral work array X / POA (rows n = m .. T+1)
    i64 rec0;              // = mrow0: start in leg_rec_n / leg_rec_s
    int isl;               // 0-based global index of the first northern latitude (ndgnh - ndglu)
    int pad;
};

struct EctFieldCfg {       // field bookkeeping of one call (INV_TRANS inv_trans.F90:212-387)
    int kf_uv = 0, kf_sc = 0;
    int scders = 0, vorgp = 0, divgp = 0, uvder = 0;
    int nleg = 0;          // Legendre fields (inverse: KF_OUT_LT, direct: KF_FS)
    int nfs = 0;           // Fourier / grid-point fields
    int cp = 0;            // record pitch in doubles = roundup(2*nleg, ECT_CPAD)
    int npairs = 0;        // field pairs of the Fourier stage
    int fp32 = 0;          // caller arrays are float
    int adj = 0;           // adjoint call (INV_TRANSAD runs the direct pipeline, DIR_TRANSAD the inverse one, with other scalings)
};

struct EctTcState;         // tcgen05 contraction of sp handles (legendre_tc.cu)
struct EctDevice {
    int dev = 0;
    EctTcState* tc = nullptr;
    int n_inv_tiles_m0 = 0, n_dir_tiles_m0 = 0;      // leading entries of inv_tiles / dir_tiles that belong to m = 0
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // geometry
    double *rw = nullptr, *racthe = nullptr;      // indexed by global latitude
    double* racthe_loc = nullptr;                 // indexed by local latitude
    double* rw_loc = nullptr;
    int *nloen = nullptr, *nmen = nullptr, *gpoff = nullptr;
    // Legendre
    double* ptab = nullptr;  i64 ptab_elems = 0;
    EctLegM* legm = nullptr;
    std::vector<EctLegM> h_legm;
    int *leg_rec_n = nullptr, *leg_rec_s = nullptr;
    int* nasm0 = nullptr;                 // per local m index
    i64 xrows = 0;                        // total rows of the spectral work array
    // tile schedules
    int2* inv_tiles = nullptr; int n_inv_tiles = 0;     // (local m, latitude tile)
    int2* dir_tiles = nullptr; int n_dir_tiles = 0;     // (local m, n tile)
    // Fourier
    EctFftTables fft;
    EctFftPlan* plans = nullptr;
    EctLatPlan* latplans = nullptr;
    uint16_t* perm_pool = nullptr;
    double2 *tw_pool = nullptr, *cz_pool = nullptr, *roots = nullptr;
    float2* cz_pool_f = nullptr;                        // sp handles: chirp tables in float (cz_pool stays null)
    int* lat_plan = nullptr;              // local lat -> latplan id
    i64* latrow0 = nullptr;
    int* fft_rec = nullptr;
    int* lat_aff = nullptr;               // per local latitude: record tables affine in the wavenumber (fourier.cu FtArgs::lat_aff)
    std::vector<int> h_lat_plan;
    // per smem-class work lists (latitudes sorted by cost)
    struct Bucket { int smem; int threads; int threads_inv = 0; int maxr; int nostage = 0; int cz = 0;   // cz: chirp-z rows on CTA pairs (k_fourier_cz)
                    std::vector<int> lats; int* d_lats = nullptr;
                    int push_sps = 0; i64 push_slot = 0; };   // direct stage, push mode: scratch slots per SM, double2 per slot
    std::vector<Bucket> buckets;
    // workspaces (grow only)
    double* xwork = nullptr; i64 xwork_elems = 0;       // X (inverse input) / POA (direct output)
    double* fbuf_leg = nullptr; i64 fbuf_leg_elems = 0; // Fourier buffer, Legendre side
    double* fbuf_fft = nullptr; i64 fbuf_fft_elems = 0; // Fourier buffer, FFT side (== leg side when 1 rank)
    // staging for host-pointer calls
    double* stage_sp = nullptr; i64 stage_sp_elems = 0;
    double* stage_gp = nullptr; i64 stage_gp_elems = 0;
    // gp_eq: band buffer the Fourier stage works on, send / receive buffers of TRLTOG / TRGTOL, exchange tables
    double* gpband = nullptr; i64 gpband_elems = 0;
    double* gpsend = nullptr; i64 gpsend_elems = 0;
    double* gprecv = nullptr; i64 gprecv_elems = 0;
    int* xb_idx = nullptr; i64* xb_off = nullptr; i64* xg_off = nullptr;
    // chunked host path: copy streams and per-slot events
    cudaStream_t cin = nullptr, cout = nullptr;
    cudaEvent_t ev_in_ready[2] = {}, ev_cmp_done[2] = {}, ev_out_done[2] = {}, ev_sp[4] = {}, ev_c0 = nullptr, ev_c1 = nullptr;
    float chunked_ms = -1.f;
    // per-call small tables
    // per-call small tables: ring of slots, each (pinned host, device) pair guarded by an event recorded
    // after the last kernel of the call that used it (calls are asynchronous in ECT_MEM_DEVICE mode)
    static const int kSlots = 4;
    void* callbuf = nullptr; void* h_callbuf = nullptr; size_t callbuf_bytes = 0;
    void* ring_d[kSlots] = {}; void* ring_h[kSlots] = {}; size_t ring_bytes[kSlots] = {};
    cudaEvent_t ring_ev[kSlots] = {}; bool ring_used[kSlots] = {}; int ring_next = 0; int ring_cur = 0;
    double* normbuf = nullptr; int normbuf_n = 0;
    // NCCL + fused (peer-memory) transposition
    void* comm = nullptr;
    void* comm_world = nullptr;           // V-sets: communicator of all W * V tasks (comm is then the W-group's)
    bool p2p = false;                     // kernels write records straight into the consumer rank's buffer
    int *leg_dst_rank_n = nullptr, *leg_dst_rank_s = nullptr, *leg_dst_rec_n = nullptr, *leg_dst_rec_s = nullptr;
    int *fft_dst_rank = nullptr, *fft_dst_rec = nullptr;
    double** peer_fft = nullptr;          // device array [nranks]: Fourier-side buffer of every rank
    double** peer_leg = nullptr;          // device array [nranks]: Legendre-side buffer of every rank
    std::vector<void*> ipc_open;          // mappings to close on reallocation / release
    int cp_alloc = 0;                     // record pitch the Fourier buffers were sized for
    int* barrier_buf = nullptr;
    int p2p_last = -1;                    // pipeline of the previous peer-mode transform (1: inverse / TRMTOL, 0: direct / TRLTOM)
    i64 entry_barriers = 0;               // consumer-done barriers issued (ect_transpose_enter)
    // direct Fourier stage in peer mode: a CTA collects the records of its field-pair chunk in a local (L2 resident)
    // slot and pushes them to the consumer rank as 256-byte runs instead of 16-byte pieces (fourier.cu, FtArgs::push_scr)
    void* push_scr = nullptr; unsigned* push_mask = nullptr; size_t push_bytes = 0; int push_nsm = 0;
    // side streams: the shared-memory classes of the Fourier stage run concurrently so that small classes fill the
    // tails of large ones
    static const int kSide = 3;
    cudaStream_t side[kSide] = {}; cudaEvent_t ev_fork = nullptr; cudaEvent_t ev_join[kSide] = {};
    // timing
    cudaEvent_t ev[16] = {};
    int last_dir = 0; bool timed = false;
    i64 launches = 0;
    int launch_error = 0;                 // a stage launcher failed (its ect_last_error text is set); checked after the stages of a call
};

// V-sets (NPRTRV > 1): tasks form a W x V grid (PE2SET: w = pe / V, v = pe % V).  hp describes the W-group of this task
// (wavenumbers, Fourier latitude band: shared by the V tasks of the group, each working on its own fields); the
// caller's grid-point arrays follow the eq_regions decomposition over all W * V tasks and carry ALL fields, so
// TRLTOG / TRGTOL redistribute points and fields at once.  The exchange tables live in hp (xb_* over the world tasks,
// xg_off over the W band owners).
struct EctVsets {
    int V = 1, v = 0, world = 1, wrank = 0;
    int ngptot = 0;                       // my grid points
};

struct EctHandle {
    EctVsets vs;
    EctHostPlan hp;
    EctDevice* d = nullptr;
    int precision = 0;
    bool defer_table = false;     // ECT_SETUP_LEGPOL_DEFER: the table is allocated but filled by ect_read_legpol
};

// setup (device)
int ect_device_setup(EctHandle* h, cudaStream_t stream, bool use_given_stream, int device, const void* nccl_uid);
void ect_device_free(EctHandle* h);

// stage launchers (all asynchronous on d->stream)
// d_vor/d_div/d_sc: device arrays of EctSpecField {base, stride} per field
void ect_launch_ltinv_prologue(EctHandle* h, const EctFieldCfg& f, const void* d_vor, const void* d_div,
                               const void* d_sc);
void ect_launch_leinv(EctHandle* h, const EctFieldCfg& f);
void ect_launch_ftinv(EctHandle* h, const EctFieldCfg& f, double* const* d_gp_base, const i64* d_gp_blkstride,
                      const void* d_fsfields, const void* d_pairs, int nproma);
void ect_launch_ftdir(EctHandle* h, const EctFieldCfg& f, double* const* d_gp_base, const i64* d_gp_blkstride,
                      const void* d_pairs, int nproma);
void ect_launch_ledir(EctHandle* h, const EctFieldCfg& f);
void ect_launch_ltdir_epilogue(EctHandle* h, const EctFieldCfg& f, void* d_vor, void* d_div, void* d_sc);
int ect_legendre_setup(EctHandle* h);
// sp handles: LEINV / LEDIR of the wavenumbers m > 0 as 3xTF32 on tcgen05 (legendre_tc.cu); m = 0 stays on the FP64 kernels
bool ect_tc_enabled(EctHandle* h);
int ect_tc_launch_leinv(EctHandle* h, const EctFieldCfg& f);
int ect_tc_operands(EctHandle* h, int cp, float** xh, float** xl);     // inverse operand rows for the prologue to fill
int ect_tc_launch_ledir(EctHandle* h, const EctFieldCfg& f);
void ect_tc_invalidate(EctHandle* h);
void ect_tc_free(EctDevice* d);
int ect_fourier_setup(EctHandle* h);
int ect_fourier_set_affine(EctHandle* h);       // after the transposition mode is decided
int ect_legendre_get_table(EctHandle* h, int ml, int par, double* out, long long cap);
int ect_legendre_set_table(EctHandle* h, int ml, int par, const double* in);      // [k][ndglu], host
int ect_transpose(EctHandle* h, const EctFieldCfg& f, int to_fft);   // TRMTOL (1) / TRLTOM (0)
int ect_transpose_enter(EctHandle* h, int to_fft);                   // peer mode: consumer-done barrier before the producing kernel

const char* ect_cuda_err(cudaError_t e);
#define ECT_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
    ect_set_error("%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); return ECT_ERR_CUDA; } } while (0)
void ect_set_error(const char* fmt, ...);
