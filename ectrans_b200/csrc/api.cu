// C ABI of the B200 spectral transform: handle registry, field bookkeeping and stage sequencing.
// Sequencing follows the reference control routines
//   INV_TRANS  cpu/external/inv_trans.F90:182-609 -> inv_trans_ctl_mod.F90:282-291
//              (LTINV -> TRMTOL -> FOURIER_IN/FSC/FTINV -> TRLTOG)
//   DIR_TRANS  cpu/external/dir_trans.F90 -> dir_trans_ctl_mod.F90
//              (TRGTOL -> FTDIR/FOURIER_OUT -> TRLTOM -> LTDIR)
#include "ect_internal.h"
#include <unistd.h>
#include "fourier_phases.h"
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>      // header-only NVTX 3: ranges cost nothing unless a profiler injects its library
#include <dlfcn.h>
#include <mutex>
#include <map>
#include <cstring>
#include <cstdio>
#include <algorithm>
#include <cstdlib>

struct EctSpecFieldH { const double* base; long long stride; };

// NVTX ranges with the reference's GSTATS labels (ectrans-benchmark.F90:1681-1697, tpm_stats.F90:33-62 turns the
// same labels into NVTX ranges on the reference's GPU branch): they bracket where a stage is ENQUEUED on the host.
struct EctRange {
    explicit EctRange(const char* name) { nvtxRangePushA(name); }
    ~EctRange() { nvtxRangePop(); }
};

static std::mutex g_mu;
static std::map<int, EctHandle*> g_handles;
static int g_next = 1;

static EctHandle* get_handle(int id) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_handles.find(id);
    return it == g_handles.end() ? nullptr : it->second;
}

// debug / tests: state of the direct stage's record push {enabled, SM id range, scratch bytes, largest slots per SM}
extern "C" int ect_debug_push_info(int handle, long long* out4) {
    EctHandle* h = get_handle(handle);
    if (!h || !h->d || !out4) return ECT_ERR_BADARG;
    int sps = 0;
    for (auto& b : h->d->buckets) sps = std::max(sps, b.push_sps);
    out4[0] = h->d->push_scr != nullptr && sps > 0; out4[1] = h->d->push_nsm; out4[2] = (long long)h->d->push_bytes; out4[3] = sps;
    return ECT_SUCCESS;
}

extern "C" const char* ect_strerror(int code) {
    switch (code) {
        case ECT_SUCCESS: return "success";
        case ECT_ERR_GENERIC: return "error";
        case ECT_ERR_NOTIMPL: return "not implemented";
        case ECT_ERR_MISSING: return "missing argument";
        case ECT_ERR_BADARG: return "unrecognised or inconsistent argument";
        case ECT_ERR_STALE: return "stale argument";
        case ECT_ERR_CUDA: return "CUDA error";
        case ECT_ERR_NCCL: return "NCCL error";
        case ECT_ERR_HANDLE: return "invalid handle";
        default: return "unknown error code";
    }
}

// NCCL is bound at run time (dlopen) and only when more than one rank is used: a DT_NEEDED libnccl.so.2 would
// pin whichever copy the loader finds first and break hosts that ship their own (PyTorch bundles a newer one
// under the same soname).  An already loaded libnccl.so.2 is reused.
struct EctNccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*) = nullptr;     // optional (NCCL >= 2.18): V-sets only
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static EctNccl g_nccl;
static int nccl_load() {
    if (g_nccl.lib) return ECT_SUCCESS;
    void* L = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!L) L = dlopen("libnccl.so.2", RTLD_NOW);
    if (!L) L = dlopen("libnccl.so", RTLD_NOW);
    if (!L) { ect_set_error("NCCL: cannot load libnccl.so.2 (%s)", dlerror()); return ECT_ERR_NCCL; }
#define ECT_SYM(field, name) do { *(void**)(&g_nccl.field) = dlsym(L, name); \
        if (!g_nccl.field) { ect_set_error("NCCL: symbol %s not found", name); return ECT_ERR_NCCL; } } while (0)
    ECT_SYM(GetUniqueId, "ncclGetUniqueId"); ECT_SYM(CommInitRank, "ncclCommInitRank"); ECT_SYM(CommDestroy, "ncclCommDestroy");
    ECT_SYM(GroupStart, "ncclGroupStart"); ECT_SYM(GroupEnd, "ncclGroupEnd"); ECT_SYM(Send, "ncclSend"); ECT_SYM(Recv, "ncclRecv");
    ECT_SYM(AllReduce, "ncclAllReduce"); ECT_SYM(AllGather, "ncclAllGather"); ECT_SYM(GetErrorString, "ncclGetErrorString");
#undef ECT_SYM
    *(void**)(&g_nccl.CommSplit) = dlsym(L, "ncclCommSplit");
    g_nccl.lib = L;
    return ECT_SUCCESS;
}
#define ncclGetUniqueId g_nccl.GetUniqueId
#define ncclCommInitRank g_nccl.CommInitRank
#define ncclCommDestroy g_nccl.CommDestroy
#define ncclGroupStart g_nccl.GroupStart
#define ncclGroupEnd g_nccl.GroupEnd
#define ncclSend g_nccl.Send
#define ncclRecv g_nccl.Recv
#define ncclAllReduce g_nccl.AllReduce
#define ncclAllGather g_nccl.AllGather
#define ncclGetErrorString g_nccl.GetErrorString

#define ECT_NCCL(call) do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) { \
    ect_set_error("%s:%d: NCCL: %s", __FILE__, __LINE__, ncclGetErrorString(r__)); return ECT_ERR_NCCL; } } while (0)

extern "C" int ect_nccl_unique_id(void* out_bytes) {
    if (!out_bytes) return ECT_ERR_MISSING;
    static_assert(sizeof(ncclUniqueId) <= ECT_NCCL_UID_BYTES, "uid size");
    ncclUniqueId id;
    { int lrc = nccl_load(); if (lrc) return lrc; }
    ECT_NCCL(ncclGetUniqueId(&id));
    memset(out_bytes, 0, ECT_NCCL_UID_BYTES);
    memcpy(out_bytes, &id, sizeof(id));
    return ECT_SUCCESS;
}

extern "C" int ect_host_alloc(void** ptr, long long bytes) {
    if (!ptr || bytes < 0) return ECT_ERR_BADARG;
    ECT_CUDA(cudaHostAlloc(ptr, (size_t)std::max<long long>(bytes, 1), cudaHostAllocDefault));
    return ECT_SUCCESS;
}
extern "C" int ect_host_free(void* ptr) {
    if (ptr) ECT_CUDA(cudaFreeHost(ptr));
    return ECT_SUCCESS;
}

// Peer-memory transposition is a collective decision taken once per handle: every pair of ranks of the W-group must
// sit on the same host, in different processes (CUDA IPC cannot map a handle of its own process), on devices that
// can address each other.  Anything else (several nodes, no NVLink / PCIe peer access, two ranks in one process)
// uses the NCCL all-to-all-v -- the reference's own TRMTOL / TRLTOM structure (trmtol_mod.F90:101-141).
struct EctPeerId { char host[64]; long long pid; char busid[32]; };
static int p2p_capable(EctHandle* h, bool wanted, bool* out) {
    const EctHostPlan& P = h->hp;
    EctDevice* d = h->d;
    ncclComm_t comm = (ncclComm_t)d->comm;
    EctPeerId me;
    memset(&me, 0, sizeof(me));
    if (gethostname(me.host, sizeof(me.host) - 1) != 0) strcpy(me.host, "?");
    {   // same hostname in two containers: the boot id tells hosts apart
        FILE* fp = fopen("/proc/sys/kernel/random/boot_id", "r");
        if (fp) { char b[40] = {0}; if (fgets(b, sizeof(b), fp)) { size_t l = strlen(me.host); strncpy(me.host + l, b, sizeof(me.host) - 1 - l); } fclose(fp); }
    }
    me.pid = (long long)getpid();
    ECT_CUDA(cudaDeviceGetPCIBusId(me.busid, sizeof(me.busid), d->dev));
    char *dsend = nullptr, *drecv = nullptr;
    ECT_CUDA(cudaMalloc(&dsend, sizeof(me)));
    ECT_CUDA(cudaMalloc(&drecv, sizeof(me) * P.nranks));
    ECT_CUDA(cudaMemcpyAsync(dsend, &me, sizeof(me), cudaMemcpyHostToDevice, d->stream));
    ECT_NCCL(ncclAllGather(dsend, drecv, sizeof(me), ncclChar, comm, d->stream));
    std::vector<EctPeerId> all(P.nranks);
    ECT_CUDA(cudaMemcpyAsync(all.data(), drecv, sizeof(me) * P.nranks, cudaMemcpyDeviceToHost, d->stream));
    ECT_CUDA(cudaStreamSynchronize(d->stream));
    int ok = wanted ? 1 : 0;
    for (int r = 0; r < P.nranks && ok; ++r) {
        if (r == P.rank) continue;
        if (memcmp(all[r].host, me.host, sizeof(me.host)) != 0 || all[r].pid == me.pid) { ok = 0; break; }
        int pd = -1, can = 0;
        if (cudaDeviceGetByPCIBusId(&pd, all[r].busid) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }   // not visible to this process
        if (pd != d->dev && (cudaDeviceCanAccessPeer(&can, d->dev, pd) != cudaSuccess || !can)) { cudaGetLastError(); ok = 0; break; }
    }
    int* dok = (int*)dsend;
    ECT_CUDA(cudaMemcpyAsync(dok, &ok, sizeof(int), cudaMemcpyHostToDevice, d->stream));
    ECT_NCCL(ncclAllReduce(dok, dok, 1, ncclInt, ncclMin, comm, d->stream));
    ECT_CUDA(cudaMemcpyAsync(&ok, dok, sizeof(int), cudaMemcpyDeviceToHost, d->stream));
    ECT_CUDA(cudaStreamSynchronize(d->stream));
    cudaFree(dsend); cudaFree(drecv);
    *out = ok != 0;
    return ECT_SUCCESS;
}

// ---------------------------------------------------------------------------------------
// setup / release
// ---------------------------------------------------------------------------------------
int ect_device_setup(EctHandle* h, cudaStream_t stream, bool use_given_stream, int device, const void* uid) {
    EctHostPlan& P = h->hp;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        ect_set_error("ect_setup: no CUDA device available (%s); this library has no CPU fallback",
                      e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return ECT_ERR_CUDA;
    }
    EctDevice* d = new EctDevice();
    h->d = d;
    if (device >= 0) ECT_CUDA(cudaSetDevice(device));
    ECT_CUDA(cudaGetDevice(&d->dev));
    if (stream || use_given_stream) { d->stream = stream; d->own_stream = false; }
    else { ECT_CUDA(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking)); d->own_stream = true; }
    for (auto& ev : d->ev) ECT_CUDA(cudaEventCreate(&ev));
    ECT_CUDA(cudaMalloc(&d->rw, P.ndgl * sizeof(double)));
    ECT_CUDA(cudaMalloc(&d->racthe, P.ndgl * sizeof(double)));
    ECT_CUDA(cudaMemcpy(d->rw, P.rw.data(), P.ndgl * sizeof(double), cudaMemcpyHostToDevice));
    ECT_CUDA(cudaMemcpy(d->racthe, P.racthe.data(), P.ndgl * sizeof(double), cudaMemcpyHostToDevice));
    int rc;
    if ((rc = ect_legendre_setup(h))) return rc;
    if ((rc = ect_fourier_setup(h))) return rc;
    if (h->vs.V > 1) {       // V-sets: one communicator over all tasks, split into the W-groups (same v)
        if (!uid) { ect_set_error("ect_setup: nranks > 1 needs nccl_uid"); return ECT_ERR_MISSING; }
        if ((rc = nccl_load())) return rc;
        ncclUniqueId id;
        memcpy(&id, uid, sizeof(id));
        ncclComm_t world;
        ECT_NCCL(ncclCommInitRank(&world, h->vs.world, id, h->vs.wrank));
        d->comm_world = world;
        if (P.nranks > 1) {
            if (!g_nccl.CommSplit) { ect_set_error("ect_setup: NPRTRV > 1 needs ncclCommSplit (NCCL >= 2.18)"); return ECT_ERR_NCCL; }
            ncclComm_t sub;
            ECT_NCCL(g_nccl.CommSplit(world, h->vs.v, P.rank, &sub, nullptr));
            d->comm = sub;
        }
    }
    if (P.nranks > 1) {
        if (h->vs.V == 1) {
            if (!uid) { ect_set_error("ect_setup: nranks > 1 needs nccl_uid"); return ECT_ERR_MISSING; }
            if ((rc = nccl_load())) return rc;
            ncclUniqueId id;
            memcpy(&id, uid, sizeof(id));
            ncclComm_t comm;
            ECT_NCCL(ncclCommInitRank(&comm, P.nranks, id, P.rank));
            d->comm = comm;
        }
        ECT_CUDA(cudaMalloc(&d->barrier_buf, 256));
        ECT_CUDA(cudaMemset(d->barrier_buf, 0, 256));
        const char* nop2p = getenv("ECT_NO_P2P");
        if ((rc = p2p_capable(h, !(nop2p && atoi(nop2p) != 0), &d->p2p))) return rc;
    }
    // destination tables of the transposition-fused stores.  One rank or NCCL mode: everything stays local
    // (rank 0 of a one-entry pointer table, record = local record); peer mode: consumer rank + its record.
    {
        const bool fused = d->p2p;
        const size_t nl = P.leg_rec_n.size(), nf = P.fft_rec.size();
        std::vector<int> zl(nl, 0), zf(nf, 0);
        auto up = [&](int*& dst, const std::vector<int>& v) -> int {
            ECT_CUDA(cudaMalloc(&dst, std::max<size_t>(v.size(), 1) * sizeof(int)));
            if (!v.empty()) ECT_CUDA(cudaMemcpy(dst, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice));
            return ECT_SUCCESS;
        };
        if ((rc = up(d->leg_dst_rank_n, fused ? P.leg_dst_rank_n : zl))) return rc;
        if ((rc = up(d->leg_dst_rank_s, fused ? P.leg_dst_rank_s : zl))) return rc;
        if ((rc = up(d->leg_dst_rec_n, fused ? P.leg_dst_rec_n : P.leg_rec_n))) return rc;
        if ((rc = up(d->leg_dst_rec_s, fused ? P.leg_dst_rec_s : P.leg_rec_s))) return rc;
        if ((rc = up(d->fft_dst_rank, fused ? P.fft_dst_rank : zf))) return rc;
        if ((rc = up(d->fft_dst_rec, fused ? P.fft_dst_rec : P.fft_rec))) return rc;
        ECT_CUDA(cudaMalloc(&d->peer_fft, P.nranks * sizeof(double*)));
        ECT_CUDA(cudaMalloc(&d->peer_leg, P.nranks * sizeof(double*)));
        if ((rc = ect_fourier_set_affine(h))) return rc;
    }
    return ECT_SUCCESS;
}

void ect_device_free(EctHandle* h) {
    EctDevice* d = h->d;
    if (!d) return;
    cudaStreamSynchronize(d->stream);
    for (void* m : d->ipc_open) cudaIpcCloseMemHandle(m);
    ect_tc_free(d);
    if (d->comm) ncclCommDestroy((ncclComm_t)d->comm);
    if (d->comm_world) ncclCommDestroy((ncclComm_t)d->comm_world);
    void* ptrs[] = {d->rw, d->racthe, d->racthe_loc, d->rw_loc, d->nloen, d->gpoff, d->ptab, d->legm, d->leg_rec_n, d->leg_rec_s,
                    d->nasm0, d->inv_tiles, d->dir_tiles, d->plans, d->latplans, d->perm_pool, d->tw_pool,
                    d->cz_pool, d->cz_pool_f, d->roots, d->lat_plan, d->latrow0, d->fft_rec, d->lat_aff, d->xwork, d->fbuf_leg,
                    d->stage_sp, d->stage_gp, d->normbuf, d->leg_dst_rank_n, d->leg_dst_rank_s, d->leg_dst_rec_n,
                    d->leg_dst_rec_s, d->fft_dst_rank, d->fft_dst_rec, d->peer_fft, d->peer_leg, d->barrier_buf, d->push_scr, d->push_mask,
                    d->gpband, d->gpsend, d->gprecv, d->xb_idx, d->xb_off, d->xg_off};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (int i = 0; i < EctDevice::kSlots; ++i) {
        if (d->ring_d[i]) cudaFree(d->ring_d[i]);
        if (d->ring_h[i]) cudaFreeHost(d->ring_h[i]);
        if (d->ring_ev[i]) cudaEventDestroy(d->ring_ev[i]);
    }
    if (d->fbuf_fft && d->fbuf_fft != d->fbuf_leg) cudaFree(d->fbuf_fft);
    for (auto& b : d->buckets) if (b.d_lats) cudaFree(b.d_lats);
    if (d->cin) { cudaStreamDestroy(d->cin); cudaStreamDestroy(d->cout); cudaEventDestroy(d->ev_c0); cudaEventDestroy(d->ev_c1); }
    for (int i = 0; i < 2; ++i) for (cudaEvent_t e : {d->ev_in_ready[i], d->ev_cmp_done[i], d->ev_out_done[i]}) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : d->ev_sp) if (e) cudaEventDestroy(e);
    for (auto& ev : d->ev) if (ev) cudaEventDestroy(ev);
    if (d->ev_fork) cudaEventDestroy(d->ev_fork);
    for (int k = 0; k < EctDevice::kSide; ++k) { if (d->ev_join[k]) cudaEventDestroy(d->ev_join[k]); if (d->side[k]) cudaStreamDestroy(d->side[k]); }
    if (d->own_stream && d->stream) cudaStreamDestroy(d->stream);
    delete d;
    h->d = nullptr;
}

extern "C" int ect_setup(const ect_setup_opts* o, int* handle) {
    EctRange r_setup("SETUP_TRANS    - Setup ecTrans handle");
    if (!o || !handle) { ect_set_error("ect_setup: null argument"); return ECT_ERR_MISSING; }
    EctHandle* h = new EctHandle();
    const int world = o->nranks < 1 ? 1 : o->nranks, V = std::max(1, (o->flags >> 8) & 0xff);
    int rc;
    if (V > 1) {
        // NPRTRV > 1: W = world / V groups of wavenumbers / latitude bands, fields spread over the V tasks of a group
        if (world % V != 0 || o->rank < 0 || o->rank >= world) { delete h; ect_set_error("ect_setup: NPROC = %d inconsistent with NPRTRV = %d", world, V); return ECT_ERR_BADARG; }
        h->vs.V = V; h->vs.world = world; h->vs.wrank = o->rank; h->vs.v = o->rank % V;
        rc = ect_build_host_plan(h->hp, o->nsmax, o->ndgl, o->nloen, world / V, o->rank / V, false, (o->flags & ECT_SETUP_BANDS_BY_POINTS) != 0);
        if (!rc) {
            EctHostPlan& P = h->hp;
            EctGpPartition G;
            rc = ect_gp_partition(P.nloen, world, G);
            if (!rc) {
                P.gp_regions = G.regions;
                P.gp_segs.assign(G.segs.begin() + G.seg0[o->rank], G.segs.begin() + G.seg0[o->rank + 1]);
                for (const EctGpSeg& sg : P.gp_segs) h->vs.ngptot += sg.count;
                P.xb_off.assign(world + 1, 0);
                for (int p = 0; p < world; ++p) {
                    for (int i = G.seg0[p]; i < G.seg0[p + 1]; ++i) {
                        const EctGpSeg& sg = G.segs[i];
                        if (sg.lat < P.lat0 || sg.lat >= P.lat0 + P.nlat) continue;
                        for (int j = 0; j < sg.count; ++j) P.xb_idx.push_back(P.gpoff[sg.lat - P.lat0] + sg.first + j);
                    }
                    P.xb_off[p + 1] = (i64)P.xb_idx.size();
                }
                P.xg_off.assign(P.nranks + 1, 0);
                for (const EctGpSeg& sg : P.gp_segs)
                    for (int r = 0; r < P.nranks; ++r)
                        if (sg.lat >= P.lat_first[r] && sg.lat < P.lat_first[r] + P.lat_count[r]) P.xg_off[r + 1] += sg.count;
                for (int r = 0; r < P.nranks; ++r) P.xg_off[r + 1] += P.xg_off[r];
            }
        }
    } else
        rc = ect_build_host_plan(h->hp, o->nsmax, o->ndgl, o->nloen, world, o->rank, (o->flags & ECT_SETUP_GP_EQ_REGIONS) != 0,
                                 (o->flags & ECT_SETUP_BANDS_BY_POINTS) != 0);
    if (rc) { delete h; return rc; }
    if (o->precision != ECT_PREC_DP && o->precision != ECT_PREC_SP) { delete h; ect_set_error("ect_setup: unknown precision %d", o->precision); return ECT_ERR_BADARG; }
    h->precision = o->precision;
    h->defer_table = (o->flags & ECT_SETUP_LEGPOL_DEFER) != 0;
    if (!(o->flags & ECT_SETUP_HOST_ONLY)) {
        rc = ect_device_setup(h, (cudaStream_t)o->stream, (o->flags & ECT_SETUP_STREAM_GIVEN) != 0, o->device, o->nccl_uid);
        if (rc) { ect_device_free(h); delete h; return rc; }
    }
    std::lock_guard<std::mutex> lk(g_mu);
    *handle = g_next++;
    g_handles[*handle] = h;
    return ECT_SUCCESS;
}

extern "C" int ect_release(int handle) {
    EctHandle* h;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_handles.find(handle);
        if (it == g_handles.end()) { ect_set_error("ect_release: invalid handle %d", handle); return ECT_ERR_HANDLE; }
        h = it->second;
        g_handles.erase(it);
    }
    ect_device_free(h);
    delete h;
    return ECT_SUCCESS;
}

extern "C" int ect_finalize(void) {
    std::vector<int> ids;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        for (auto& kv : g_handles) ids.push_back(kv.first);
    }
    for (int id : ids) ect_release(id);
    return ECT_SUCCESS;
}

extern "C" int ect_inquire(int handle, ect_info* info) {
    EctHandle* h = get_handle(handle);
    if (!h) { ect_set_error("ect_inquire: invalid handle %d", handle); return ECT_ERR_HANDLE; }
    if (!info) return ECT_ERR_MISSING;
    const EctHostPlan& P = h->hp;
    info->nsmax = P.nsmax; info->ndgl = P.ndgl; info->ndgnh = P.ndgnh;
    info->nranks = P.nranks; info->rank = P.rank;
    info->nspec2 = P.nspec2; info->nspec2g = P.nspec2_g;
    info->ngptot = h->vs.V > 1 ? h->vs.ngptot : P.ngptot; info->ngptotg = P.ngptotg;
    info->nump = P.nump; info->lat0 = P.lat0; info->nlat = P.nlat;
    info->table_bytes = h->d ? h->d->ptab_elems * (long long)sizeof(double) : 0;
    return ECT_SUCCESS;
}

template <typename T, typename S>
static int copy_out(void* out, long long cap, const std::vector<S>& v) {
    if ((long long)v.size() > cap) { ect_set_error("ect_inquire_array: capacity %lld < %zu", cap, v.size()); return ECT_ERR_BADARG; }
    T* o = (T*)out;
    for (size_t i = 0; i < v.size(); ++i) o[i] = (T)v[i];
    return ECT_SUCCESS;
}

extern "C" int ect_inquire_array(int handle, int which, void* out, long long cap) {
    EctHandle* h = get_handle(handle);
    if (!h) { ect_set_error("ect_inquire_array: invalid handle %d", handle); return ECT_ERR_HANDLE; }
    if (!out) return ECT_ERR_MISSING;
    const EctHostPlan& P = h->hp;
    switch (which) {
        case ECT_ARR_NLOEN: return copy_out<int>(out, cap, P.nloen);
        case ECT_ARR_NMEN: return copy_out<int>(out, cap, P.nmen);
        case ECT_ARR_NDGLU: return copy_out<int>(out, cap, P.ndglu);
        case ECT_ARR_MYMS: return copy_out<int>(out, cap, P.myms);
        case ECT_ARR_NASM0: return copy_out<int>(out, cap, P.nasm0);
        case ECT_ARR_NPROCM: return copy_out<int>(out, cap, P.nprocm);
        case ECT_ARR_RMU: return copy_out<double>(out, cap, P.rmu);
        case ECT_ARR_RGW: return copy_out<double>(out, cap, P.rw);
        case ECT_ARR_RACTHE: return copy_out<double>(out, cap, P.racthe);
        case ECT_ARR_LATFIRST: return copy_out<int>(out, cap, P.lat_first);
        case ECT_ARR_LATCOUNT: return copy_out<int>(out, cap, P.lat_count);
        case ECT_ARR_SENDCNT: return copy_out<long long>(out, cap, P.send_cnt);
        case ECT_ARR_RECVCNT: return copy_out<long long>(out, cap, P.recv_cnt);
        case ECT_ARR_SENDOFF: return copy_out<long long>(out, cap, P.send_off);
        case ECT_ARR_RECVOFF: return copy_out<long long>(out, cap, P.recv_off);
        case ECT_ARR_MROW0: return copy_out<long long>(out, cap, P.mrow0);
        case ECT_ARR_LEGRECN: return copy_out<int>(out, cap, P.leg_rec_n);
        case ECT_ARR_LEGRECS: return copy_out<int>(out, cap, P.leg_rec_s);
        case ECT_ARR_LATROW0: return copy_out<long long>(out, cap, P.latrow0);
        case ECT_ARR_FFTREC: return copy_out<int>(out, cap, P.fft_rec);
        case ECT_ARR_LEGDSTRANKN: return copy_out<int>(out, cap, P.leg_dst_rank_n);
        case ECT_ARR_LEGDSTRECN: return copy_out<int>(out, cap, P.leg_dst_rec_n);
        case ECT_ARR_LEGDSTRANKS: return copy_out<int>(out, cap, P.leg_dst_rank_s);
        case ECT_ARR_LEGDSTRECS: return copy_out<int>(out, cap, P.leg_dst_rec_s);
        case ECT_ARR_FFTDSTRANK: return copy_out<int>(out, cap, P.fft_dst_rank);
        case ECT_ARR_FFTDSTREC: return copy_out<int>(out, cap, P.fft_dst_rec);
        case ECT_ARR_GPSEGS: case ECT_ARR_NGPSEGS: {
            std::vector<int> v;
            if (P.gp_eq || h->vs.V > 1) for (const EctGpSeg& sg : P.gp_segs) { v.push_back(sg.lat); v.push_back(sg.first); v.push_back(sg.count); }
            else for (int l = 0; l < P.nlat; ++l) { v.push_back(P.lat0 + l); v.push_back(0); v.push_back(P.nloen[P.lat0 + l]); }
            if (which == ECT_ARR_NGPSEGS) return copy_out<int>(out, cap, std::vector<int>(1, (int)v.size() / 3));
            return copy_out<int>(out, cap, v);
        }
        case ECT_ARR_XBIDX: return copy_out<int>(out, cap, P.xb_idx);
        case ECT_ARR_XBOFF: return copy_out<long long>(out, cap, P.xb_off);
        case ECT_ARR_XGOFF: return copy_out<long long>(out, cap, P.xg_off);
        case ECT_ARR_NREGIONS: {
            std::vector<int> v(P.nranks, 0);
            for (size_t i = 0; i < P.gp_regions.size(); ++i) v[i] = P.gp_regions[i];
            return copy_out<int>(out, cap, v);
        }
        default: ect_set_error("ect_inquire_array: unknown array id %d", which); return ECT_ERR_BADARG;
    }
}

// ---------------------------------------------------------------------------------------
// workspaces
// ---------------------------------------------------------------------------------------
static int ensure(double*& p, i64& have, i64 need, cudaStream_t s, bool zero) {
    if (need <= have && p) return ECT_SUCCESS;
    if (p) { ECT_CUDA(cudaStreamSynchronize(s)); ECT_CUDA(cudaFree(p)); p = nullptr; have = 0; }
    need = std::max<i64>(need, 16);
    ECT_CUDA(cudaMalloc(&p, (size_t)need * sizeof(double)));
    if (zero) ECT_CUDA(cudaMemsetAsync(p, 0, (size_t)need * sizeof(double), s));
    have = need;
    return ECT_SUCCESS;
}

// Publishes the (re)allocated Fourier buffers: pointer tables for the transposition-fused stores.  In peer mode
// the buffers of all ranks are mapped through CUDA IPC; the handles travel through an NCCL all-gather.
static int publish_buffers(EctHandle* h) {
    const EctHostPlan& P = h->hp;
    EctDevice* d = h->d;
    std::vector<double*> pf(P.nranks, nullptr), pl(P.nranks, nullptr);
    if (!d->p2p) {
        // inverse stores go to the local Legendre-side buffer (then NCCL moves them), direct stores to the local
        // Fourier-side buffer; with one rank both are the same memory
        for (int r = 0; r < P.nranks; ++r) { pf[r] = d->fbuf_leg; pl[r] = d->fbuf_fft; }
    } else {
        for (void* m : d->ipc_open) ECT_CUDA(cudaIpcCloseMemHandle(m));
        d->ipc_open.clear();
        cudaIpcMemHandle_t mine[2];
        ECT_CUDA(cudaIpcGetMemHandle(&mine[0], d->fbuf_fft));
        ECT_CUDA(cudaIpcGetMemHandle(&mine[1], d->fbuf_leg));
        const size_t hb = sizeof(mine);
        char *dsend = nullptr, *drecv = nullptr;
        ECT_CUDA(cudaMalloc(&dsend, hb));
        ECT_CUDA(cudaMalloc(&drecv, hb * P.nranks));
        ECT_CUDA(cudaMemcpyAsync(dsend, mine, hb, cudaMemcpyHostToDevice, d->stream));
        ECT_NCCL(ncclAllGather(dsend, drecv, hb, ncclChar, (ncclComm_t)d->comm, d->stream));
        std::vector<cudaIpcMemHandle_t> all(2 * P.nranks);
        ECT_CUDA(cudaMemcpyAsync(all.data(), drecv, hb * P.nranks, cudaMemcpyDeviceToHost, d->stream));
        ECT_CUDA(cudaStreamSynchronize(d->stream));
        cudaFree(dsend); cudaFree(drecv);
        for (int r = 0; r < P.nranks; ++r) {
            if (r == P.rank) { pf[r] = d->fbuf_fft; pl[r] = d->fbuf_leg; continue; }
            void *a = nullptr, *b = nullptr;
            ECT_CUDA(cudaIpcOpenMemHandle(&a, all[2 * r], cudaIpcMemLazyEnablePeerAccess));
            ECT_CUDA(cudaIpcOpenMemHandle(&b, all[2 * r + 1], cudaIpcMemLazyEnablePeerAccess));
            d->ipc_open.push_back(a); d->ipc_open.push_back(b);
            pf[r] = (double*)a; pl[r] = (double*)b;
        }
    }
    ECT_CUDA(cudaMemcpyAsync(d->peer_fft, pf.data(), P.nranks * sizeof(double*), cudaMemcpyHostToDevice, d->stream));
    ECT_CUDA(cudaMemcpyAsync(d->peer_leg, pl.data(), P.nranks * sizeof(double*), cudaMemcpyHostToDevice, d->stream));
    ECT_CUDA(cudaStreamSynchronize(d->stream));
    return ECT_SUCCESS;
}

static int ensure_work(EctHandle* h, const EctFieldCfg& f) {
    EctDevice* d = h->d;
    const EctHostPlan& P = h->hp;
    int rc;
    if ((rc = ensure(d->xwork, d->xwork_elems, (d->xrows + 2) * (i64)f.cp, d->stream, true))) return rc;
    // the Fourier buffers grow by record pitch only (same decision on every rank: the IPC exchange is collective)
    if (f.cp > d->cp_alloc) {
        if (d->p2p) {
            // peers hold IPC mappings of the buffers about to be freed and may still be storing into them: wait until
            // every rank has finished all work queued so far (collective, like the exchange of the new handles below)
            ECT_NCCL(ncclAllReduce(d->barrier_buf + 16, d->barrier_buf + 24, 1, ncclInt, ncclSum, (ncclComm_t)d->comm, d->stream));
            ECT_CUDA(cudaStreamSynchronize(d->stream));
            for (void* m : d->ipc_open) ECT_CUDA(cudaIpcCloseMemHandle(m));
            d->ipc_open.clear();
            d->p2p_last = -1;
        }
        if ((rc = ensure(d->fbuf_leg, d->fbuf_leg_elems, (P.nrec_leg + 1) * (i64)f.cp, d->stream, true))) return rc;
        if (P.nranks > 1) {
            if ((rc = ensure(d->fbuf_fft, d->fbuf_fft_elems, (P.nrec_fft + 1) * (i64)f.cp, d->stream, true))) return rc;
        } else {
            d->fbuf_fft = d->fbuf_leg;
            d->fbuf_fft_elems = d->fbuf_leg_elems;
        }
        d->cp_alloc = f.cp;
        if ((rc = publish_buffers(h))) return rc;
    }
    return ECT_SUCCESS;
}

// Per-call tables go to the device through a tiny kernel that reads the pinned (mapped) host buffer, not through
// cudaMemcpyAsync: a DMA copy would queue behind the multi-GB field copies of the chunked host path in the same
// copy engine and stall the compute stream until they finish (measured: 16 ms of kernels took 60 ms per chunk).
__global__ void k_upload_tables(int4* __restrict__ dst, const int4* __restrict__ src, int n16) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) dst[i] = src[i];
}
static int upload_callbuf(EctDevice* d, size_t bytes) {
    const int n16 = (int)((bytes + 15) / 16);
    void* hdev = nullptr;
    ECT_CUDA(cudaHostGetDevicePointer(&hdev, d->h_callbuf, 0));
    k_upload_tables<<<std::min(64, (n16 + 255) / 256), 256, 0, d->stream>>>((int4*)d->callbuf, (const int4*)hdev, n16);
    d->launches++;
    return ECT_SUCCESS;
}

static int ensure_callbuf(EctDevice* d, size_t bytes) {
    const int i = d->ring_next;
    d->ring_next = (i + 1) % EctDevice::kSlots;
    d->ring_cur = i;
    if (!d->ring_ev[i]) ECT_CUDA(cudaEventCreateWithFlags(&d->ring_ev[i], cudaEventDisableTiming));
    if (d->ring_used[i]) ECT_CUDA(cudaEventSynchronize(d->ring_ev[i]));     // previous user of this slot is done
    if (bytes > d->ring_bytes[i]) {
        if (d->ring_d[i]) { ECT_CUDA(cudaFree(d->ring_d[i])); ECT_CUDA(cudaFreeHost(d->ring_h[i])); }
        bytes = (bytes + 4095) / 4096 * 4096;
        ECT_CUDA(cudaMalloc(&d->ring_d[i], bytes));
        ECT_CUDA(cudaHostAlloc(&d->ring_h[i], bytes, cudaHostAllocMapped));
        d->ring_bytes[i] = bytes;
    }
    d->callbuf = d->ring_d[i];
    d->h_callbuf = d->ring_h[i];
    return ECT_SUCCESS;
}
static int release_callbuf(EctDevice* d) {
    ECT_CUDA(cudaEventRecord(d->ring_ev[d->ring_cur], d->stream));
    d->ring_used[d->ring_cur] = true;
    return ECT_SUCCESS;
}

__global__ void k_debug_delay(long long cycles) {
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) { }
}

// Peer mode, before the producing kernel of a transform: its stores land in buffers that the CONSUMER kernel of the
// previous transform may still be reading on another rank (rank A's k_leinv of call c+1 writes rank B's Fourier-side
// buffer while B's k_fourier<inverse> of call c reads it; the same for k_fourier<direct> -> k_ledir).  One barrier
// per transform, after the producer, does not order that.  Alternating inverse / direct calls are ordered by the
// other direction's barrier (different buffer pair); two transforms on the same pipeline in a row -- inv, inv or
// dir, dir: the chunked host path, several INV_TRANS per time step -- need this second, consumer-done barrier.
int ect_transpose_enter(EctHandle* h, int to_fft) {
    const EctHostPlan& P = h->hp;
    EctDevice* d = h->d;
    if (P.nranks == 1 || !d->p2p) return ECT_SUCCESS;
    static const char* mode = getenv("ECT_P2P_ENTRY_BARRIER");      // 0: never (reproduces the round-1 hazard), 2: always
    const int m = mode ? atoi(mode) : 1;
    if ((m == 1 && d->p2p_last == to_fft) || m == 2) {
        ECT_NCCL(ncclAllReduce(d->barrier_buf + 16, d->barrier_buf + 24, 1, ncclInt, ncclSum, (ncclComm_t)d->comm, d->stream));
        d->entry_barriers++;
    }
    d->p2p_last = to_fft;
    return ECT_SUCCESS;
}

// TRMTOL (to_fft = 1) / TRLTOM (to_fft = 0): all-to-all-v of (lat, m) records over the W-set.
// Reference cpu/internal/trmtol_mod.F90:101-141, trltom_mod.F90; single task = same buffer.
int ect_transpose(EctHandle* h, const EctFieldCfg& f, int to_fft) {
    const EctHostPlan& P = h->hp;
    EctDevice* d = h->d;
    if (P.nranks == 1) return ECT_SUCCESS;
    ncclComm_t comm = (ncclComm_t)d->comm;
    if (d->p2p) {
        // the producing kernels already wrote every record into its consumer's buffer over NVLink: all that is
        // left of TRMTOL / TRLTOM is "everybody has finished writing"
        ECT_NCCL(ncclAllReduce(d->barrier_buf, d->barrier_buf + 8, 1, ncclInt, ncclSum, comm, d->stream));
        // test hook (tools/selfcheck_run.py): the last rank's consumer kernel starts late, so that its peers reach the
        // producer of their next transform while it still reads -- the schedule the consumer-done barrier exists for
        static const char* delay = getenv("ECT_DEBUG_DELAY_CONSUMER_MS");
        if (delay && atof(delay) > 0 && P.rank == P.nranks - 1) {
            k_debug_delay<<<1, 1, 0, d->stream>>>((long long)(atof(delay) * 1.9e6));
            d->launches++;
        }
        return ECT_SUCCESS;
    }
    const i64 cp = f.cp;
    ECT_NCCL(ncclGroupStart());
    for (int p = 0; p < P.nranks; ++p) {
        double* lb = d->fbuf_leg + P.send_off[p] * cp;
        double* fb = d->fbuf_fft + P.recv_off[p] * cp;
        const size_t nl = (size_t)(P.send_cnt[p] * cp), nf = (size_t)(P.recv_cnt[p] * cp);
        if (to_fft) {
            if (nl) ECT_NCCL(ncclSend(lb, nl, ncclDouble, p, comm, d->stream));
            if (nf) ECT_NCCL(ncclRecv(fb, nf, ncclDouble, p, comm, d->stream));
        } else {
            if (nf) ECT_NCCL(ncclSend(fb, nf, ncclDouble, p, comm, d->stream));
            if (nl) ECT_NCCL(ncclRecv(lb, nl, ncclDouble, p, comm, d->stream));
        }
    }
    ECT_NCCL(ncclGroupEnd());
    return ECT_SUCCESS;
}

// ---------------------------------------------------------------------------------------
// TRLTOG / TRGTOL for the eq_regions grid-point partition (cpu/internal/trltog_mod.F90:213-271, trgtol_mod.F90): an
// all-to-all between the owners of the Fourier latitude bands and the grid-point tasks.  The Fourier stage works on a
// band buffer B(point of the band, field); messages are [field][points of the pair (band owner, task)], points in
// the task's local order, so both ends address them with one offset table each:
//   band owner:  message to task p   = points xb_idx[xb_off[p] .. xb_off[p+1]) of B
//   task:        message from owner r = local points [xg_off[r], xg_off[r+1])  of the caller's PGP arrays
// ---------------------------------------------------------------------------------------
template <typename T, bool TO_MSG>
__global__ void k_gp_band_msg(T* __restrict__ band, T* __restrict__ msg, const int* __restrict__ idx,
                              const i64* __restrict__ off, int nranks, int nband, int nf) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nband) return;
    int lo = 0, hi = nranks;                       // task p with off[p] <= k < off[p+1]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid] <= k) lo = mid; else hi = mid; }
    const i64 o = off[lo], cnt = off[lo + 1] - o;
    const int src = idx[k];
    for (int f = blockIdx.y; f < nf; f += gridDim.y) {
        T* m = msg + o * nf + (i64)f * cnt + (k - o);
        T* b = band + (i64)f * nband + src;
        if (TO_MSG) *m = *b; else *b = *m;
    }
}
template <typename T, bool TO_MSG>
__global__ void k_gp_user_msg(double* const* __restrict__ base, const i64* __restrict__ blkstride, T* __restrict__ msg,
                              const i64* __restrict__ off, int nranks, int ngptot, int nproma, int nf) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngptot) return;
    int lo = 0, hi = nranks;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid] <= g) lo = mid; else hi = mid; }
    const i64 o = off[lo], cnt = off[lo + 1] - o;
    const int blk = g / nproma, in = g - blk * nproma;
    for (int f = blockIdx.y; f < nf; f += gridDim.y) {
        T* m = msg + o * nf + (i64)f * cnt + (g - o);
        T* u = reinterpret_cast<T*>(base[f]) + (i64)blk * blkstride[f] + in;
        if (TO_MSG) *m = *u; else *u = *m;
    }
}

// nf_band: fields of the band buffer (this task's V-set); nf_user: fields of the caller's arrays (all V-sets)
static int gp_exchange_setup(EctHandle* h, int nf_band, int nf_user) {
    const EctHostPlan& P = h->hp;
    EctDevice* d = h->d;
    const int ntasks = h->vs.V > 1 ? h->vs.world : P.nranks;
    const int ngp_user = h->vs.V > 1 ? h->vs.ngptot : P.ngptot;
    int rc;
    if (!d->xb_idx) {
        ECT_CUDA(cudaMalloc(&d->xb_idx, std::max<size_t>(P.xb_idx.size(), 1) * sizeof(int)));
        ECT_CUDA(cudaMalloc(&d->xb_off, (ntasks + 1) * sizeof(i64)));
        ECT_CUDA(cudaMalloc(&d->xg_off, (P.nranks + 1) * sizeof(i64)));
        ECT_CUDA(cudaMemcpy(d->xb_idx, P.xb_idx.data(), P.xb_idx.size() * sizeof(int), cudaMemcpyHostToDevice));
        ECT_CUDA(cudaMemcpy(d->xb_off, P.xb_off.data(), (ntasks + 1) * sizeof(i64), cudaMemcpyHostToDevice));
        ECT_CUDA(cudaMemcpy(d->xg_off, P.xg_off.data(), (P.nranks + 1) * sizeof(i64), cudaMemcpyHostToDevice));
    }
    const i64 nb = (i64)std::max(P.ngpband, 1) * std::max(nf_band, 1), ng = (i64)std::max(ngp_user, 1) * std::max(nf_user, 1);
    if ((rc = ensure(d->gpband, d->gpband_elems, nb, d->stream, false))) return rc;
    if ((rc = ensure(d->gpsend, d->gpsend_elems, std::max(nb, ng), d->stream, false))) return rc;
    if ((rc = ensure(d->gprecv, d->gprecv_elems, std::max(nb, ng), d->stream, false))) return rc;
    return ECT_SUCCESS;
}

// to_grid = 1: TRLTOG (band buffer -> caller's arrays), 0: TRGTOL.  d_gpb / d_gps: the caller-array field table on the
// device, in MESSAGE order: the fields of V-set 0 first, then V-set 1, ... (nfl[v] of them; one V-set: all fields).
static int gp_exchange(EctHandle* h, int nf_band, int nf_user, const std::vector<int>& nfl, int es, int to_grid,
                       double* const* d_gpb, const i64* d_gps, int nproma) {
    const EctHostPlan& P = h->hp;
    EctDevice* d = h->d;
    const int V = h->vs.V, W = P.nranks;
    const int ntasks = V > 1 ? h->vs.world : P.nranks;
    const int ngp_user = V > 1 ? h->vs.ngptot : P.ngptot;
    ncclComm_t comm = (ncclComm_t)(V > 1 ? d->comm_world : d->comm);
    const dim3 gb((P.ngpband + 255) / 256, std::max(1, std::min(nf_band, 64))), gu((ngp_user + 255) / 256, std::max(1, std::min(nf_user, 64)));
#define ECT_GPX(T) do { \
        if (to_grid) { if (P.ngpband && nf_band) k_gp_band_msg<T, true><<<gb, 256, 0, d->stream>>>((T*)d->gpband, (T*)d->gpsend, d->xb_idx, d->xb_off, ntasks, P.ngpband, nf_band); } \
        else if (ngp_user && nf_user) k_gp_user_msg<T, true><<<gu, 256, 0, d->stream>>>(d_gpb, d_gps, (T*)d->gpsend, d->xg_off, W, ngp_user, nproma, nf_user); \
    } while (0)
    if (es == 4) ECT_GPX(float); else ECT_GPX(double);
#undef ECT_GPX
    // band side: one message per task p, nf_band fields; task side: one message per (band owner w, V-set v), nfl[v] fields,
    // stored inside the block of w ([field slot][points of w]) behind the fields of the V-sets before v
    char* bandbuf = (char*)(to_grid ? d->gpsend : d->gprecv);
    char* userbuf = (char*)(to_grid ? d->gprecv : d->gpsend);
    ECT_NCCL(ncclGroupStart());
    for (int p = 0; p < ntasks; ++p) {
        const size_t nb = (size_t)(P.xb_off[p + 1] - P.xb_off[p]) * nf_band * es;
        if (!nb) continue;
        char* at = bandbuf + (size_t)P.xb_off[p] * nf_band * es;
        if (to_grid) ECT_NCCL(ncclSend(at, nb, ncclChar, p, comm, d->stream));
        else ECT_NCCL(ncclRecv(at, nb, ncclChar, p, comm, d->stream));
    }
    for (int w = 0; w < W; ++w) {
        const size_t cnt = (size_t)(P.xg_off[w + 1] - P.xg_off[w]);
        size_t pre = 0;
        for (int v = 0; v < V; ++v) {
            const size_t nu = cnt * nfl[v] * es;
            char* at = userbuf + ((size_t)P.xg_off[w] * nf_user + cnt * pre) * es;
            pre += nfl[v];
            if (!nu) continue;
            if (to_grid) ECT_NCCL(ncclRecv(at, nu, ncclChar, w * V + v, comm, d->stream));
            else ECT_NCCL(ncclSend(at, nu, ncclChar, w * V + v, comm, d->stream));
        }
    }
    ECT_NCCL(ncclGroupEnd());
#define ECT_GPX(T) do { \
        if (to_grid) { if (ngp_user && nf_user) k_gp_user_msg<T, false><<<gu, 256, 0, d->stream>>>(d_gpb, d_gps, (T*)d->gprecv, d->xg_off, W, ngp_user, nproma, nf_user); } \
        else if (P.ngpband && nf_band) k_gp_band_msg<T, false><<<gb, 256, 0, d->stream>>>((T*)d->gpband, (T*)d->gprecv, d->xb_idx, d->xb_off, ntasks, P.ngpband, nf_band); \
    } while (0)
    if (es == 4) ECT_GPX(float); else ECT_GPX(double);
#undef ECT_GPX
    ECT_CUDA(cudaGetLastError());
    d->launches += 2;
    return ECT_SUCCESS;
}

// ---------------------------------------------------------------------------------------
// field bookkeeping shared by both directions
// ---------------------------------------------------------------------------------------
struct ScalarSeg { const double* host; long long nfld_in_array; int lev, fld; int kind; };  // kind 0: (nf,nspec2) 1: 3-D

struct CallLayout {
    int kf_uv = 0, kf_sc = 0, nsc2 = 0, n3a_lev = 0, n3a_fld = 0, n3b_lev = 0, n3b_fld = 0;
    bool mode2_sp = false;
};

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

// pointer arithmetic on caller arrays in elements of the handle's precision (members are typed double*)
template <typename T> static T* adv(T* p, i64 n, int es) { return (T*)((char*)p + n * es); }
template <typename T> static const T* adv(const T* p, i64 n, int es) { return (const T*)((const char*)p + n * es); }

static std::vector<int2> make_pairs(const std::vector<int>& groups) {
    std::vector<int2> pairs;
    int f0 = 0;
    for (int g : groups) {
        for (int j = 0; j < g; j += 2) pairs.push_back(make_int2(f0 + j, j + 1 < g ? f0 + j + 1 : -1));
        f0 += g;
    }
    return pairs;
}


// ---------------------------------------------------------------------------------------
// Host-pointer calls on large field sets: the call is split into field chunks and pipelined -- chunk c+1's
// inputs travel host->device and chunk c-1's results device->host (each on its own copy stream) while chunk c
// is transformed, so that the PCIe time of the larger side hides everything else.  A chunk is an ordinary
// device-pointer call on dense staging arrays; chunks never mix scalar fields of different caller arrays
// (fields of one array share complex transforms, see make_pairs).  The chunk plan depends on global sizes only,
// so every rank issues the same sequence of (collective) sub-calls.
// ---------------------------------------------------------------------------------------
// sub-call view: the chunk's spectral fields live inside larger staged arrays (row pitch in fields)
struct SubView { int uv_stride, sc_stride; };
static int inv_trans_impl(int handle, const ect_inv_args* a, const SubView* view, int adj);
static int dir_trans_impl(int handle, const ect_dir_args* a, const SubView* view, int adj);
struct HostChunk { int j0, nj; int seg, s0, ns; };   // uv levels [j0, j0+nj), scalars [s0, s0+ns) (flattened index) of segment seg
// one caller array slice: fields [start, start+count) contiguous with pitch = count; arr = which caller array,
// dev = the slice inside the staged copy of that array
struct ScSeg { double* base; int start, count; int arr; double* dev; };

static int host_chunk_target() {          // Legendre fields per chunk; ECT_HOST_CHUNK_FIELDS=0 disables chunking
    const char* e = getenv("ECT_HOST_CHUNK_FIELDS");
    return e ? atoi(e) : 32;
}
static int host_chunk_count(EctHandle* h, int nleg) {
    const int target = host_chunk_target();
    if (target <= 0) return 1;
    const char* mb = getenv("ECT_HOST_CHUNK_MIN_BYTES");
    const double minb = mb ? atof(mb) : 256e6;
    if ((double)h->hp.ngptotg * nleg * 8.0 < minb || nleg < 2 * target) return 1;
    return (nleg + target - 1) / target;
}

static std::vector<HostChunk> plan_chunks(int kf_uv, int w_uv, const std::vector<ScSeg>& segs, int w_sc) {
    const int target = std::max(1, host_chunk_target());
    std::vector<HostChunk> out;
    auto split = [&](int n, int per, auto emit) {
        if (n <= 0) return;
        const int nc = (n + per - 1) / per;
        int i0 = 0;
        for (int c = 0; c < nc; ++c) { const int cnt = n / nc + (c < n % nc ? 1 : 0); emit(i0, cnt); i0 += cnt; }
    };
    split(kf_uv, std::max(1, target / w_uv), [&](int i0, int cnt) { out.push_back({i0, cnt, -1, 0, 0}); });
    for (int g = 0; g < (int)segs.size(); ++g)
        split(segs[g].count, std::max(1, target / w_sc), [&](int i0, int cnt) { out.push_back({0, 0, g, segs[g].start + i0, cnt}); });
    return out;
}

static int copy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                  cudaMemcpyKind kind, cudaStream_t st) {
    if (width == 0 || height == 0) return ECT_SUCCESS;
    if (height == 1 || (dpitch == width && spitch == width)) {
        ECT_CUDA(cudaMemcpyAsync(dst, src, width * height, kind, st));
        return ECT_SUCCESS;
    }
    const size_t maxp = (size_t)1 << 31;
    if (dpitch >= maxp || spitch >= maxp) {
        for (size_t r = 0; r < height; ++r)
            ECT_CUDA(cudaMemcpyAsync((char*)dst + r * dpitch, (const char*)src + r * spitch, width, kind, st));
        return ECT_SUCCESS;
    }
    ECT_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, kind, st));
    return ECT_SUCCESS;
}

// dense chunk array (nproma, nf_c, ngpblks) on the device <-> the caller's fields gidx[0..nf_c) (runs of fields that
// are neighbours in the caller's array move as one 2-D copy)
static int copy_gp_fields(bool to_host, char* dev, const std::vector<int>& gidx, const std::vector<double*>& hb,
                          const std::vector<i64>& hs, int nproma, int ngpblks, int es, cudaStream_t st) {
    const int nfc = (int)gidx.size();
    int rc;
    for (int i = 0; i < nfc;) {
        int j = i + 1;
        while (j < nfc && hs[gidx[j]] == hs[gidx[i]] &&
               (char*)hb[gidx[j]] == (char*)hb[gidx[i]] + (size_t)(j - i) * nproma * es) ++j;
        const size_t width = (size_t)(j - i) * nproma * es;
        char* dp = dev + (size_t)i * nproma * es;
        const size_t dpitch = (size_t)nfc * nproma * es, hpitch = (size_t)hs[gidx[i]] * es;
        if (to_host) rc = copy2d(hb[gidx[i]], hpitch, dp, dpitch, width, ngpblks, cudaMemcpyDeviceToHost, st);
        else rc = copy2d(dp, dpitch, hb[gidx[i]], hpitch, width, ngpblks, cudaMemcpyHostToDevice, st);
        if (rc) return rc;
        i = j;
    }
    return ECT_SUCCESS;
}

// ECT_HOST_CHUNK_DBG=1: per-chunk timeline (ms after the start of the call) on stderr
struct ChunkTrace {
    bool on; cudaEvent_t t0; std::vector<cudaEvent_t> ev; std::vector<const char*> tag; std::vector<int> chunk;
    explicit ChunkTrace(cudaEvent_t start) : t0(start) { const char* e = getenv("ECT_HOST_CHUNK_DBG"); on = e && atoi(e); }
    void mark(const char* what, int c, cudaStream_t st) {
        if (!on) return;
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st);
        ev.push_back(e); tag.push_back(what); chunk.push_back(c);
    }
    void dump(const char* name) {
        if (!on) return;
        for (size_t i = 0; i < ev.size(); ++i) {
            float ms = 0.f; cudaEventElapsedTime(&ms, t0, ev[i]);
            fprintf(stderr, "[%s] chunk %2d %-9s %8.2f ms\n", name, chunk[i], tag[i], ms);
            cudaEventDestroy(ev[i]);
        }
    }
};

static int ensure_pipe(EctDevice* d) {
    if (d->cin) return ECT_SUCCESS;
    ECT_CUDA(cudaStreamCreateWithFlags(&d->cin, cudaStreamNonBlocking));
    ECT_CUDA(cudaStreamCreateWithFlags(&d->cout, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        ECT_CUDA(cudaEventCreateWithFlags(&d->ev_in_ready[i], cudaEventDisableTiming));
        ECT_CUDA(cudaEventCreateWithFlags(&d->ev_cmp_done[i], cudaEventDisableTiming));
        ECT_CUDA(cudaEventCreateWithFlags(&d->ev_out_done[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 4; ++i) ECT_CUDA(cudaEventCreateWithFlags(&d->ev_sp[i], cudaEventDisableTiming));
    ECT_CUDA(cudaEventCreate(&d->ev_c0));
    ECT_CUDA(cudaEventCreate(&d->ev_c1));
    return ECT_SUCCESS;
}

// The spectral arrays are (nspec2, fields) with the field index fastest: a field chunk is a column block, whose
// narrow-row 2-D copies run far below PCIe speed.  They are therefore staged whole (they are the small side,
// 1/5 of the bytes) -- array by array in the order the chunks need them -- and the sub-calls address their
// columns inside the staged arrays (SubView); the grid-point side moves chunk by chunk.
struct SpStage {
    double* dev[5] = {};       // staged vor, div, scalar arrays (mode 1: spscalar; mode 2: sc2, sc3a, sc3b)
    double* host[5] = {};
    i64 elems[5] = {};
};

static std::vector<ScSeg> scalar_segments(bool mode2, const SpStage& st, int kf_sc, int nsc2, int lev_a, int fld_a,
                                          int lev_b, int fld_b, i64 nsp, int es) {
    std::vector<ScSeg> segs;
    if (!mode2) { if (kf_sc) segs.push_back({st.host[2], 0, kf_sc, 2, st.dev[2]}); return segs; }
    int s0 = 0;
    if (nsc2) { segs.push_back({st.host[2], s0, nsc2, 2, st.dev[2]}); s0 += nsc2; }
    for (int j3 = 0; j3 < fld_a; ++j3) {
        segs.push_back({adv(st.host[3], (i64)j3 * lev_a * nsp, es), s0, lev_a, 3, adv(st.dev[3], (i64)j3 * lev_a * nsp, es)});
        s0 += lev_a;
    }
    for (int j3 = 0; j3 < fld_b; ++j3) {
        segs.push_back({adv(st.host[4], (i64)j3 * lev_b * nsp, es), s0, lev_b, 4, adv(st.dev[4], (i64)j3 * lev_b * nsp, es)});
        s0 += lev_b;
    }
    return segs;
}

static int layout_sp_stage(EctDevice* d, SpStage& st, int es) {
    i64 tot = 0;
    for (int i = 0; i < 5; ++i) tot += (st.elems[i] + 31) / 32 * 32;
    int rc;
    if ((rc = ensure(d->stage_sp, d->stage_sp_elems, tot, d->stream, false))) return rc;
    double* p = d->stage_sp;
    for (int i = 0; i < 5; ++i) { st.dev[i] = st.elems[i] ? p : nullptr; p = adv(p, (st.elems[i] + 31) / 32 * 32, es); }
    return ECT_SUCCESS;
}

template <typename Fill>
static int inv_trans_chunked(int handle, EctHandle* h, const ect_inv_args* a, const EctFieldCfg& f, bool mode2_sp,
                             int nsc2, int n3a, int n3b, int nproma, int ngpblks, Fill& fill_gp_table, int adj) {
    EctDevice* d = h->d;
    const EctHostPlan& P = h->hp;
    const int es = f.fp32 ? 4 : 8;
    const i64 nsp = P.nspec2, blk = (i64)nproma * ngpblks;
    const int kf_uv = f.kf_uv, kf_sc = f.kf_sc;
    int rc;
    if ((rc = ensure_pipe(d))) return rc;
    SpStage st;
    st.host[0] = (double*)a->spvor; st.host[1] = (double*)a->spdiv; st.elems[0] = st.elems[1] = (i64)kf_uv * nsp;
    if (!mode2_sp) { st.host[2] = (double*)a->spscalar; st.elems[2] = (i64)kf_sc * nsp; }
    else {
        st.host[2] = (double*)a->spsc2; st.elems[2] = (i64)nsc2 * nsp;
        st.host[3] = (double*)a->spsc3a; st.elems[3] = (i64)n3a * nsp;
        st.host[4] = (double*)a->spsc3b; st.elems[4] = (i64)n3b * nsp;
    }
    if ((rc = layout_sp_stage(d, st, es))) return rc;
    const std::vector<ScSeg> segs = scalar_segments(mode2_sp, st, kf_sc, nsc2, a->nsc3a_lev, n3a ? a->nsc3a_fld : 0,
                                                    a->nsc3b_lev, n3b ? a->nsc3b_fld : 0, nsp, es);
    std::vector<double*> hgpb(f.nfs);
    std::vector<i64> hgps(f.nfs);
    fill_gp_table(a->gp, a->gpuv, a->gp2, a->gp3a, a->gp3b, hgpb.data(), hgps.data());
    // global Fourier-field index of each kind (order: [vor][div] u v scalars [nsd] [du dv] [ewd])
    const int n_vor = f.vorgp ? kf_uv : 0, n_div = f.divgp ? kf_uv : 0, n_nsd = f.scders ? kf_sc : 0;
    const int g_vor = 0, g_div = n_vor, g_u = n_vor + n_div, g_v = g_u + kf_uv, g_sc = g_v + kf_uv, g_nsd = g_sc + kf_sc,
              g_du = g_nsd + n_nsd, g_dv = g_du + kf_uv, g_ewd = g_du + (f.uvder ? 2 * kf_uv : 0);
    const int w_uv = 2 + (f.vorgp ? 1 : 0) + (f.divgp ? 1 : 0), w_sc = 1 + (f.scders ? 1 : 0);
    // scalar chunks first: their arrays are the smaller ones, so the first results leave sooner
    std::vector<HostChunk> chunks = plan_chunks(kf_uv, w_uv, segs, w_sc);
    std::stable_partition(chunks.begin(), chunks.end(), [](const HostChunk& c) { return c.ns > 0; });
    i64 max_out = 0;
    for (const HostChunk& c : chunks)
        max_out = std::max(max_out, (i64)(c.nj * (w_uv + (f.uvder ? 2 : 0)) + c.ns * (f.scders ? 3 : 1)) * blk);
    max_out = (max_out + 31) / 32 * 32;
    if ((rc = ensure(d->stage_gp, d->stage_gp_elems, 2 * max_out, d->stream, false))) return rc;
    ECT_CUDA(cudaEventRecord(d->ev_c0, d->stream));
    ECT_CUDA(cudaStreamWaitEvent(d->cin, d->ev_c0, 0));
    ChunkTrace trace(d->ev_c0);
    // ---- spectral inputs, whole arrays in the order of use (copy-in stream) ----
    for (int i : {2, 3, 4, 0, 1}) {
        if (st.elems[i]) ECT_CUDA(cudaMemcpyAsync(st.dev[i], st.host[i], (size_t)st.elems[i] * es, cudaMemcpyHostToDevice, d->cin));
        if (i >= 1) ECT_CUDA(cudaEventRecord(d->ev_sp[i - 1], d->cin));       // ev_sp[0]: vor and div, ev_sp[1..3]: scalar arrays
        trace.mark("sp_in", i, d->cin);
    }
    i64 launches = 0;
    for (size_t ic = 0; ic < chunks.size(); ++ic) {
        const HostChunk& c = chunks[ic];
        const int slot = (int)(ic & 1);
        double* out = adv(d->stage_gp, slot * max_out, es);
        // ---- transform (handle stream): a device-pointer call on columns of the staged arrays ----
        ECT_CUDA(cudaStreamWaitEvent(d->stream, d->ev_sp[c.nj ? 0 : segs[c.seg].arr - 1], 0));
        if (ic >= 2) ECT_CUDA(cudaStreamWaitEvent(d->stream, d->ev_out_done[slot], 0));
        ect_inv_args sub;
        memset(&sub, 0, sizeof(sub));
        sub.memspace = ECT_MEM_DEVICE; sub.nproma = a->nproma;
        sub.scders = a->scders; sub.vorgp = a->vorgp; sub.divgp = a->divgp; sub.uvder = a->uvder;
        SubView view{kf_uv, 0};
        if (c.nj) { sub.spvor = adv(st.dev[0], c.j0, es); sub.spdiv = adv(st.dev[1], c.j0, es); sub.nuv = c.nj; }
        if (c.ns) {
            const ScSeg& sg = segs[c.seg];
            sub.spscalar = adv(sg.dev, c.s0 - sg.start, es); sub.nscalar = c.ns; view.sc_stride = sg.count;
        }
        sub.gp = out;
        trace.mark("cmp_start", (int)ic, d->stream);
        if ((rc = inv_trans_impl(handle, &sub, &view, adj))) return rc;
        launches += d->launches;
        ECT_CUDA(cudaEventRecord(d->ev_cmp_done[slot], d->stream));
        trace.mark("cmp_done", (int)ic, d->stream);
        // ---- results (copy-out stream) ----
        std::vector<int> gidx;
        auto add = [&](int g0, int i0, int n) { for (int i = 0; i < n; ++i) gidx.push_back(g0 + i0 + i); };
        if (f.vorgp) add(g_vor, c.j0, c.nj);
        if (f.divgp) add(g_div, c.j0, c.nj);
        add(g_u, c.j0, c.nj); add(g_v, c.j0, c.nj);
        add(g_sc, c.s0, c.ns);
        if (f.scders) add(g_nsd, c.s0, c.ns);
        if (f.uvder) { add(g_du, c.j0, c.nj); add(g_dv, c.j0, c.nj); }
        if (f.scders) add(g_ewd, c.s0, c.ns);
        ECT_CUDA(cudaStreamWaitEvent(d->cout, d->ev_cmp_done[slot], 0));
        if ((rc = copy_gp_fields(true, (char*)out, gidx, hgpb, hgps, nproma, ngpblks, es, d->cout))) return rc;
        ECT_CUDA(cudaEventRecord(d->ev_out_done[slot], d->cout));
        trace.mark("out_done", (int)ic, d->cout);
    }
    for (int i = 0; i < 2 && i < (int)chunks.size(); ++i) ECT_CUDA(cudaStreamWaitEvent(d->stream, d->ev_out_done[i], 0));
    ECT_CUDA(cudaEventRecord(d->ev_c1, d->stream));
    ECT_CUDA(cudaStreamSynchronize(d->stream));
    ECT_CUDA(cudaEventElapsedTime(&d->chunked_ms, d->ev_c0, d->ev_c1));
    trace.dump("inv");
    d->launches = launches;
    d->last_dir = 0;
    return ECT_SUCCESS;
}

template <typename Fill>
static int dir_trans_chunked(int handle, EctHandle* h, const ect_dir_args* a, const EctFieldCfg& f, bool mode2,
                             int nsc2, int n3a, int n3b, int nproma, int ngpblks, Fill& fill_gp_table, int adj) {
    EctDevice* d = h->d;
    const EctHostPlan& P = h->hp;
    const int es = f.fp32 ? 4 : 8;
    const i64 nsp = P.nspec2, blk = (i64)nproma * ngpblks;
    const int kf_uv = f.kf_uv, kf_sc = f.kf_sc;
    int rc;
    if ((rc = ensure_pipe(d))) return rc;
    SpStage st;
    st.host[0] = a->spvor; st.host[1] = a->spdiv; st.elems[0] = st.elems[1] = (i64)kf_uv * nsp;
    if (!mode2) { st.host[2] = a->spscalar; st.elems[2] = (i64)kf_sc * nsp; }
    else {
        st.host[2] = a->spsc2; st.elems[2] = (i64)nsc2 * nsp;
        st.host[3] = a->spsc3a; st.elems[3] = (i64)n3a * nsp;
        st.host[4] = a->spsc3b; st.elems[4] = (i64)n3b * nsp;
    }
    if ((rc = layout_sp_stage(d, st, es))) return rc;
    const std::vector<ScSeg> segs = scalar_segments(mode2, st, kf_sc, nsc2, a->nsc3a_lev, n3a ? a->nsc3a_fld : 0,
                                                    a->nsc3b_lev, n3b ? a->nsc3b_fld : 0, nsp, es);
    std::vector<double*> hgpb(f.nfs);
    std::vector<i64> hgps(f.nfs);
    fill_gp_table((double*)a->gp, (double*)a->gpuv, (double*)a->gp2, (double*)a->gp3a, (double*)a->gp3b, hgpb.data(), hgps.data());
    const std::vector<HostChunk> chunks = plan_chunks(kf_uv, 2, segs, 1);
    i64 max_in = 0;
    for (const HostChunk& c : chunks) max_in = std::max(max_in, (2 * (i64)c.nj + c.ns) * blk);
    max_in = (max_in + 31) / 32 * 32;
    if ((rc = ensure(d->stage_gp, d->stage_gp_elems, 2 * max_in, d->stream, false))) return rc;
    ECT_CUDA(cudaEventRecord(d->ev_c0, d->stream));
    ECT_CUDA(cudaStreamWaitEvent(d->cin, d->ev_c0, 0));
    ECT_CUDA(cudaStreamWaitEvent(d->cout, d->ev_c0, 0));
    ChunkTrace trace(d->ev_c0);
    if (adj) {     // INV_TRANSAD adds to the spectral arrays: bring their present contents first
        for (int i = 0; i < 5; ++i)
            if (st.elems[i]) ECT_CUDA(cudaMemcpyAsync(st.dev[i], st.host[i], (size_t)st.elems[i] * es, cudaMemcpyHostToDevice, d->cin));
    }
    i64 launches = 0;
    for (size_t ic = 0; ic < chunks.size(); ++ic) {
        const HostChunk& c = chunks[ic];
        const int slot = (int)(ic & 1);
        double* in = adv(d->stage_gp, slot * max_in, es);
        // ---- grid-point inputs of the chunk (copy-in stream) ----
        if (ic >= 2) ECT_CUDA(cudaStreamWaitEvent(d->cin, d->ev_cmp_done[slot], 0));
        std::vector<int> gidx;     // global input field of each chunk field: u levels, v levels, scalars
        for (int i = 0; i < c.nj; ++i) gidx.push_back(c.j0 + i);
        for (int i = 0; i < c.nj; ++i) gidx.push_back(kf_uv + c.j0 + i);
        for (int i = 0; i < c.ns; ++i) gidx.push_back(2 * kf_uv + c.s0 + i);
        if ((rc = copy_gp_fields(false, (char*)in, gidx, hgpb, hgps, nproma, ngpblks, es, d->cin))) return rc;
        ECT_CUDA(cudaEventRecord(d->ev_in_ready[slot], d->cin));
        trace.mark("in_done", (int)ic, d->cin);
        // ---- transform: results go to columns of the staged spectral arrays ----
        ECT_CUDA(cudaStreamWaitEvent(d->stream, d->ev_in_ready[slot], 0));
        ect_dir_args sub;
        memset(&sub, 0, sizeof(sub));
        sub.memspace = ECT_MEM_DEVICE; sub.nproma = a->nproma;
        sub.nuv = c.nj; sub.nscalar = c.ns; sub.gp = in;
        SubView view{kf_uv, 0};
        if (c.nj) { sub.spvor = adv(st.dev[0], c.j0, es); sub.spdiv = adv(st.dev[1], c.j0, es); }
        if (c.ns) { const ScSeg& sg = segs[c.seg]; sub.spscalar = adv(sg.dev, c.s0 - sg.start, es); view.sc_stride = sg.count; }
        trace.mark("cmp_start", (int)ic, d->stream);
        if ((rc = dir_trans_impl(handle, &sub, &view, adj))) return rc;
        launches += d->launches;
        ECT_CUDA(cudaEventRecord(d->ev_cmp_done[slot], d->stream));
        trace.mark("cmp_done", (int)ic, d->stream);
        // ---- a spectral array goes back as soon as its last chunk is done (copy-out stream) ----
        const bool last = ic + 1 == chunks.size();
        const int arr_now = c.nj ? 0 : segs[c.seg].arr, arr_next = last ? -1 : (chunks[ic + 1].nj ? 0 : segs[chunks[ic + 1].seg].arr);
        if (arr_now != arr_next) {
            ECT_CUDA(cudaStreamWaitEvent(d->cout, d->ev_cmp_done[slot], 0));
            for (int i = (arr_now == 0 ? 0 : arr_now); i <= (arr_now == 0 ? 1 : arr_now); ++i)
                if (st.elems[i]) ECT_CUDA(cudaMemcpyAsync(st.host[i], st.dev[i], (size_t)st.elems[i] * es, cudaMemcpyDeviceToHost, d->cout));
            trace.mark("sp_out", arr_now, d->cout);
        }
    }
    ECT_CUDA(cudaEventRecord(d->ev_out_done[0], d->cout));
    ECT_CUDA(cudaStreamWaitEvent(d->stream, d->ev_out_done[0], 0));
    ECT_CUDA(cudaEventRecord(d->ev_c1, d->stream));
    ECT_CUDA(cudaStreamSynchronize(d->stream));
    ECT_CUDA(cudaEventElapsedTime(&d->chunked_ms, d->ev_c0, d->ev_c1));
    trace.dump("dir");
    d->launches = launches;
    d->last_dir = 1;
    return ECT_SUCCESS;
}

static int vset_guard(int handle, const char* who) {
    EctHandle* h = get_handle(handle);
    if (h && h->vs.V > 1) { ect_set_error("%s: this resolution has NPRTRV = %d; use the *_vset entry points (V-set arrays needed) -- not available for this call", who, h->vs.V); return ECT_ERR_NOTIMPL; }
    return ECT_SUCCESS;
}
extern "C" int ect_inv_trans(int handle, const ect_inv_args* a) { EctRange r("INV_TRANS      - Inverse transform"); int g = vset_guard(handle, "ect_inv_trans"); return g ? g : inv_trans_impl(handle, a, nullptr, 0); }
extern "C" int ect_dir_trans(int handle, const ect_dir_args* a) { EctRange r("DIR_TRANS      - Direct transform"); int g = vset_guard(handle, "ect_dir_trans"); return g ? g : dir_trans_impl(handle, a, nullptr, 0); }

// INV_TRANSAD / DIR_TRANSAD (reference cpu/external/inv_transad.F90, dir_transad.F90 and the *ad_mod
// files of cpu/internal).  With the inner products of the reference's adjoint tests (grid: plain sum; spectral:
// weight 2 for m > 0, 1 for the real parts of m = 0; tests/trans/test_invtrans_adjoint.F90:242-311)
//   INV_TRANSAD = M^-1 INV_TRANS^T = the direct pipeline without the 1/N of the FFT and without Gaussian weights,
//                 whose (U, V) -> (vor, div) step VDTUV^T equals diag(-RLAPIN(n)) . UVTVD; results are ADDED to the
//                 spectral arrays (prfi1bad_mod.F90:91-108);
//   DIR_TRANSAD = DIR_TRANS^T M = the inverse pipeline with every latitude row scaled by w / N, whose
//                 (vor, div) -> (U, V) step UVTVD^T equals VDTUV . diag(-1 / RLAPIN(n)).
// The derivative / vorticity-divergence grid-point options of INV_TRANSAD are not provided.
extern "C" int ect_inv_transad(int handle, const ect_inv_args* a) {
    if (int g = vset_guard(handle, "ect_inv_transad")) return g;
    if (!a) return ECT_ERR_MISSING;
    if (a->scders || a->vorgp || a->divgp || a->uvder) {
        ect_set_error("ect_inv_transad: scders / vorgp / divgp / uvder are not implemented for the adjoint");
        return ECT_ERR_NOTIMPL;
    }
    ect_dir_args d;
    memset(&d, 0, sizeof(d));
    d.memspace = a->memspace; d.nproma = a->nproma; d.nuv = a->nuv;
    d.gp = a->gp; d.gpuv = a->gpuv; d.gp2 = a->gp2; d.gp3a = a->gp3a; d.gp3b = a->gp3b;
    d.nsc2 = a->nsc2; d.nsc3a_lev = a->nsc3a_lev; d.nsc3a_fld = a->nsc3a_fld; d.nsc3b_lev = a->nsc3b_lev; d.nsc3b_fld = a->nsc3b_fld;
    d.nscalar = a->spscalar ? a->nscalar : (a->spsc2 ? a->nsc2 : 0) + (a->spsc3a ? a->nsc3a_lev * a->nsc3a_fld : 0) +
                                           (a->spsc3b ? a->nsc3b_lev * a->nsc3b_fld : 0);
    d.spvor = (double*)a->spvor; d.spdiv = (double*)a->spdiv; d.spscalar = (double*)a->spscalar;
    d.spsc2 = (double*)a->spsc2; d.spsc3a = (double*)a->spsc3a; d.spsc3b = (double*)a->spsc3b;
    return dir_trans_impl(handle, &d, nullptr, 1);
}
extern "C" int ect_dir_transad(int handle, const ect_dir_args* a) {
    if (int g = vset_guard(handle, "ect_dir_transad")) return g;
    if (!a) return ECT_ERR_MISSING;
    ect_inv_args v;
    memset(&v, 0, sizeof(v));
    v.memspace = a->memspace; v.nproma = a->nproma; v.nuv = a->nuv;
    v.spvor = a->spvor; v.spdiv = a->spdiv; v.spscalar = a->spscalar; v.nscalar = a->nscalar;
    v.spsc2 = a->spsc2; v.nsc2 = a->nsc2; v.spsc3a = a->spsc3a; v.nsc3a_lev = a->nsc3a_lev; v.nsc3a_fld = a->nsc3a_fld;
    v.spsc3b = a->spsc3b; v.nsc3b_lev = a->nsc3b_lev; v.nsc3b_fld = a->nsc3b_fld;
    v.gp = (double*)a->gp; v.gpuv = (double*)a->gpuv; v.gp2 = (double*)a->gp2; v.gp3a = (double*)a->gp3a; v.gp3b = (double*)a->gp3b;
    return inv_trans_impl(handle, &v, nullptr, 1);
}

static int inv_trans_impl(int handle, const ect_inv_args* a, const SubView* view, int adj) {
    EctHandle* h = get_handle(handle);
    if (!h) { ect_set_error("ect_inv_trans: invalid handle %d", handle); return ECT_ERR_HANDLE; }
    if (!a) return ECT_ERR_MISSING;
    if (!h->d) { ect_set_error("ect_inv_trans: handle was set up host-only (no device state)"); return ECT_ERR_CUDA; }
    const EctHostPlan& P = h->hp;
    EctDevice* d = h->d;
    ECT_CUDA(cudaSetDevice(d->dev));
    // ---- field counting (inv_trans.F90:212-387) ----
    const int kf_uv = a->nuv;
    if (kf_uv < 0 || a->nscalar < 0) { ect_set_error("ect_inv_trans: negative field count"); return ECT_ERR_BADARG; }
    if (kf_uv > 0 && (!a->spvor || !a->spdiv)) { ect_set_error("ect_inv_trans: nuv > 0 needs spvor and spdiv"); return ECT_ERR_MISSING; }
    const bool mode2_sp = (a->spscalar == nullptr) && (a->spsc2 || a->spsc3a || a->spsc3b);
    int nsc2 = 0, n3a = 0, n3b = 0, kf_sc = 0;
    if (mode2_sp) {
        nsc2 = a->spsc2 ? a->nsc2 : 0;
        n3a = a->spsc3a ? a->nsc3a_lev * a->nsc3a_fld : 0;
        n3b = a->spsc3b ? a->nsc3b_lev * a->nsc3b_fld : 0;
        kf_sc = nsc2 + n3a + n3b;
    } else {
        kf_sc = a->spscalar ? a->nscalar : 0;
    }
    const bool mode2_gp = (a->gp == nullptr);
    if (mode2_gp) {
        if (kf_uv > 0 && !a->gpuv) { ect_set_error("ect_inv_trans: gp == NULL and gpuv == NULL"); return ECT_ERR_MISSING; }
        if (!mode2_sp && kf_sc > 0) { ect_set_error("ect_inv_trans: PSPSCALAR requires PGP (inv_trans.F90:418-424)"); return ECT_ERR_BADARG; }
        if (nsc2 > 0 && !a->gp2) { ect_set_error("ect_inv_trans: spsc2 given without gp2"); return ECT_ERR_MISSING; }
        if (n3a > 0 && !a->gp3a) { ect_set_error("ect_inv_trans: spsc3a given without gp3a"); return ECT_ERR_MISSING; }
        if (n3b > 0 && !a->gp3b) { ect_set_error("ect_inv_trans: spsc3b given without gp3b"); return ECT_ERR_MISSING; }
    }
    EctFieldCfg f;
    f.kf_uv = kf_uv; f.kf_sc = kf_sc;
    f.scders = a->scders && kf_sc > 0; f.vorgp = a->vorgp && kf_uv > 0; f.divgp = a->divgp && kf_uv > 0;
    f.uvder = a->uvder && kf_uv > 0;
    const int n_vor = f.vorgp ? kf_uv : 0, n_div = f.divgp ? kf_uv : 0, n_nsd = f.scders ? kf_sc : 0;
    f.nleg = n_vor + n_div + 2 * kf_uv + kf_sc + n_nsd;
    f.nfs = f.nleg + (f.uvder ? 2 * kf_uv : 0) + n_nsd;
    if (f.nfs == 0) return ECT_SUCCESS;
    f.cp = round_up(2 * f.nleg, ECT_CPAD);
    f.fp32 = (h->precision == ECT_PREC_SP);
    f.adj = adj;
    const int es = f.fp32 ? 4 : 8;
    const int nproma = (a->nproma > 0 && a->nproma < P.ngptot) ? a->nproma : std::max(P.ngptot, 1);
    const int ngpblks = (P.ngptot + nproma - 1) / nproma;
    int rc;
    const bool host = (a->memspace == ECT_MEM_HOST);
    // grid-point field table: base pointer and block stride of every Fourier field (trltog_mod.F90:579-731);
    // pure address arithmetic, used with device pointers (kernel tables) and with the caller's host pointers
    // (chunked host path)
    const int nvar_uv = (f.vorgp ? 1 : 0) + (f.divgp ? 1 : 0) + 2 + (f.uvder ? 2 : 0);
    const int dfac = f.scders ? 3 : 1;
    auto fill_gp_table = [&](double* q_gp, double* q_uv, double* q_2, double* q_3a, double* q_3b, double** t_gpb, i64* t_gps) {
        if (!mode2_gp) {
            for (int i = 0; i < f.nfs; ++i) { t_gpb[i] = adv(q_gp, (i64)i * nproma, es); t_gps[i] = (i64)f.nfs * nproma; }
        } else {
            int fi = 0, var = 0;
            auto uvgroup = [&](int v) { for (int l = 0; l < kf_uv; ++l, ++fi) { t_gpb[fi] = adv(q_uv, ((i64)v * kf_uv + l) * nproma, es); t_gps[fi] = (i64)nproma * kf_uv * nvar_uv; } };
            if (f.vorgp) uvgroup(var++);
            if (f.divgp) uvgroup(var++);
            if (kf_uv) { uvgroup(var++); uvgroup(var++); }
            auto scgroup = [&](int part) {   // part 0: fields, 1: N-S derivatives, 2: E-W derivatives
                for (int j = 0; j < nsc2; ++j, ++fi) { t_gpb[fi] = adv(q_2, ((i64)part * nsc2 + j) * nproma, es); t_gps[fi] = (i64)nproma * nsc2 * dfac; }
                for (int j3 = 0; j3 < (n3a ? a->nsc3a_fld : 0); ++j3)
                    for (int l = 0; l < a->nsc3a_lev; ++l, ++fi) {
                        t_gpb[fi] = adv(q_3a, (((i64)part * a->nsc3a_fld + j3) * a->nsc3a_lev + l) * nproma, es);
                        t_gps[fi] = (i64)nproma * a->nsc3a_lev * a->nsc3a_fld * dfac;
                    }
                for (int j3 = 0; j3 < (n3b ? a->nsc3b_fld : 0); ++j3)
                    for (int l = 0; l < a->nsc3b_lev; ++l, ++fi) {
                        t_gpb[fi] = adv(q_3b, (((i64)part * a->nsc3b_fld + j3) * a->nsc3b_lev + l) * nproma, es);
                        t_gps[fi] = (i64)nproma * a->nsc3b_lev * a->nsc3b_fld * dfac;
                    }
            };
            scgroup(0);
            if (f.scders) scgroup(1);
            if (f.uvder) { uvgroup(var++); uvgroup(var++); }
            if (f.scders) scgroup(2);
        }
    };
    if (host) {
        const int nch = host_chunk_count(h, f.nleg);
        if (nch > 1) return inv_trans_chunked(handle, h, a, f, mode2_sp, nsc2, n3a, n3b, nproma, ngpblks, fill_gp_table, adj);
    }
    if ((rc = ensure_work(h, f))) return rc;
    d->launches = 0;
    d->chunked_ms = -1.f;
    ECT_CUDA(cudaEventRecord(d->ev[0], d->stream));
    // ---- spectral inputs on the device ----
    const i64 nsp = P.nspec2;
    const double *dvor = a->spvor, *ddiv = a->spdiv, *dsc = a->spscalar, *dsc2 = a->spsc2, *dsc3a = a->spsc3a, *dsc3b = a->spsc3b;
    if (host) {
        const i64 tot = (2 * (i64)kf_uv + kf_sc) * nsp;
        if ((rc = ensure(d->stage_sp, d->stage_sp_elems, tot, d->stream, false))) return rc;
        double* p = d->stage_sp;
        auto stage = [&](const double*& ptr, i64 n) -> int {
            if (!ptr || n == 0) return ECT_SUCCESS;
            ECT_CUDA(cudaMemcpyAsync(p, ptr, (size_t)n * es, cudaMemcpyHostToDevice, d->stream));
            ptr = p; p = adv(p, n, es);
            return ECT_SUCCESS;
        };
        if (kf_uv) { if ((rc = stage(dvor, kf_uv * nsp))) return rc; if ((rc = stage(ddiv, kf_uv * nsp))) return rc; }
        if (!mode2_sp) { if ((rc = stage(dsc, kf_sc * nsp))) return rc; }
        else {
            if ((rc = stage(dsc2, nsc2 * nsp))) return rc;
            if ((rc = stage(dsc3a, n3a * nsp))) return rc;
            if ((rc = stage(dsc3b, n3b * nsp))) return rc;
        }
    }
    ECT_CUDA(cudaEventRecord(d->ev[1], d->stream));
    // ---- grid-point outputs on the device ----
    double *dgp = a->gp, *dgpuv = a->gpuv, *dgp2 = a->gp2, *dgp3a = a->gp3a, *dgp3b = a->gp3b;
    const i64 blk = (i64)nproma * ngpblks;
    const i64 sz_gp = (i64)f.nfs * blk, sz_uv = (i64)kf_uv * nvar_uv * blk, sz_2 = (i64)nsc2 * dfac * blk,
              sz_3a = (i64)n3a * dfac * blk, sz_3b = (i64)n3b * dfac * blk;
    if (host) {
        const i64 tot = mode2_gp ? (sz_uv + sz_2 + sz_3a + sz_3b) : sz_gp;
        if ((rc = ensure(d->stage_gp, d->stage_gp_elems, tot, d->stream, false))) return rc;
        double* p = d->stage_gp;
        if (!mode2_gp) dgp = p;
        else { dgpuv = p; p = adv(p, sz_uv, es); dgp2 = p; p = adv(p, sz_2, es); dgp3a = p; p = adv(p, sz_3a, es); dgp3b = p; }
    }
    // ---- per-call tables ----
    const size_t n_spec = (size_t)(2 * kf_uv + kf_sc);
    // field groups: only fields of the same kind / array share a complex transform (keeps their
    // rounding errors decoupled: magnitudes within a group are comparable)
    std::vector<int> groups;   // sizes in Fourier order
    {
        auto scal = [&]() {
            if (!mode2_sp) { if (kf_sc) groups.push_back(kf_sc); return; }
            if (nsc2) groups.push_back(nsc2);
            for (int j3 = 0; j3 < (n3a ? a->nsc3a_fld : 0); ++j3) groups.push_back(a->nsc3a_lev);
            for (int j3 = 0; j3 < (n3b ? a->nsc3b_fld : 0); ++j3) groups.push_back(a->nsc3b_lev);
        };
        if (f.vorgp) groups.push_back(kf_uv);
        if (f.divgp) groups.push_back(kf_uv);
        if (kf_uv) { groups.push_back(kf_uv); groups.push_back(kf_uv); }
        scal();
        if (f.scders) scal();
        if (f.uvder) { groups.push_back(kf_uv); groups.push_back(kf_uv); }
        if (f.scders) scal();
    }
    std::vector<int2> pairs = make_pairs(groups);
    f.npairs = (int)pairs.size();
    const bool gpx = P.gp_eq;       // TRLTOG is an exchange: the Fourier stage writes the band buffer, not the caller's arrays
    if (gpx && (rc = gp_exchange_setup(h, f.nfs, f.nfs))) return rc;
    const size_t bytes = n_spec * sizeof(EctSpecFieldH) + (size_t)f.nfs * (2 * (sizeof(double*) + sizeof(i64)) + sizeof(EctFsField)) +
                         pairs.size() * sizeof(int2) + 64;
    if ((rc = ensure_callbuf(d, bytes))) return rc;
    char* hb = (char*)d->h_callbuf;
    EctSpecFieldH* t_vor = (EctSpecFieldH*)hb;
    EctSpecFieldH* t_div = t_vor + kf_uv;
    EctSpecFieldH* t_sc = t_div + kf_uv;
    double** t_gpb = (double**)(t_sc + kf_sc);
    i64* t_gps = (i64*)(t_gpb + f.nfs);
    double** t_bandb = (double**)(t_gps + f.nfs);       // gp_eq: field table of the band buffer
    i64* t_bands = (i64*)(t_bandb + f.nfs);
    int2* t_pairs = (int2*)(t_bands + f.nfs);
    EctFsField* t_fs = (EctFsField*)(t_pairs + pairs.size());
    memcpy(t_pairs, pairs.data(), pairs.size() * sizeof(int2));
    for (int i = 0; i < f.nfs; ++i) { t_bandb[i] = gpx ? adv(d->gpband, (i64)i * P.ngpband, es) : nullptr; t_bands[i] = (i64)f.nfs * P.ngpband; }
    const int st_uv = view ? view->uv_stride : kf_uv, st_sc = view ? view->sc_stride : kf_sc;
    for (int j = 0; j < kf_uv; ++j) { t_vor[j] = {adv(dvor, j, es), st_uv}; t_div[j] = {adv(ddiv, j, es), st_uv}; }
    if (!mode2_sp) for (int s = 0; s < kf_sc; ++s) t_sc[s] = {adv(dsc, s, es), st_sc};
    else {
        int s = 0;
        for (int j = 0; j < nsc2; ++j) t_sc[s++] = {adv(dsc2, j, es), nsc2};
        for (int j3 = 0; j3 < (a->spsc3a ? a->nsc3a_fld : 0); ++j3)
            for (int l = 0; l < a->nsc3a_lev; ++l) t_sc[s++] = {adv(dsc3a, (i64)j3 * a->nsc3a_lev * nsp + l, es), a->nsc3a_lev};
        for (int j3 = 0; j3 < (a->spsc3b ? a->nsc3b_fld : 0); ++j3)
            for (int l = 0; l < a->nsc3b_lev; ++l) t_sc[s++] = {adv(dsc3b, (i64)j3 * a->nsc3b_lev * nsp + l, es), a->nsc3b_lev};
    }
    // Fourier-space field list: [vor][div] u v scalars [nsd] [du dv] [ewd]   (ftinv_ctl_mod.F90:144-166)
    {
        int fi = 0, lf = 0;    // Fourier index, Legendre index
        const int l_u = n_vor + n_div, l_sc = l_u + 2 * kf_uv;
        for (int j = 0; j < n_vor + n_div; ++j, ++fi, ++lf) t_fs[fi] = {2 * lf, 0, 0};
        for (int j = 0; j < 2 * kf_uv; ++j, ++fi, ++lf) t_fs[fi] = {2 * lf, 1, 0};
        for (int j = 0; j < kf_sc; ++j, ++fi, ++lf) t_fs[fi] = {2 * lf, 0, 0};
        for (int j = 0; j < n_nsd; ++j, ++fi, ++lf) t_fs[fi] = {2 * lf, 1, 0};
        if (f.uvder) for (int j = 0; j < 2 * kf_uv; ++j, ++fi) t_fs[fi] = {2 * (l_u + j), 1, 1};
        for (int j = 0; j < n_nsd; ++j, ++fi) t_fs[fi] = {2 * (l_sc + j), 0, 1};
    }
    // grid-point destinations (trltog_mod.F90:579-731)
    fill_gp_table(dgp, dgpuv, dgp2, dgp3a, dgp3b, t_gpb, t_gps);
    if ((rc = upload_callbuf(d, bytes))) return rc;
    char* db = (char*)d->callbuf;
    const void* d_vor = db;
    const void* d_div = db + ((char*)t_div - hb);
    const void* d_sc = db + ((char*)t_sc - hb);
    double* const* d_gpb = (double* const*)(db + ((char*)t_gpb - hb));
    const i64* d_gps = (const i64*)(db + ((char*)t_gps - hb));
    const void* d_fs = db + ((char*)t_fs - hb);
    const void* d_pairs = db + ((char*)t_pairs - hb);
    // ---- stages ----
    {
        EctRange r("LTINV_CTL      - Inv. Legendre transform");
        ect_launch_ltinv_prologue(h, f, d_vor, d_div, d_sc);
        ECT_CUDA(cudaEventRecord(d->ev[2], d->stream));
        if ((rc = ect_transpose_enter(h, 1))) return rc;
        ect_launch_leinv(h, f);
        ECT_CUDA(cudaEventRecord(d->ev[3], d->stream));
    }
    {
        EctRange r("LTINV_CTL      - M to L transposition");
        if ((rc = ect_transpose(h, f, 1))) return rc;
        ECT_CUDA(cudaEventRecord(d->ev[4], d->stream));
    }
    EctRange r_ft("FTINV_CTL      - Inv. Fourier transform");
    if (gpx) {
        double* const* d_bandb = (double* const*)(db + ((char*)t_bandb - hb));
        const i64* d_bands = (const i64*)(db + ((char*)t_bands - hb));
        ect_launch_ftinv(h, f, d_bandb, d_bands, d_fs, d_pairs, std::max(P.ngpband, 1));
        ECT_CUDA(cudaEventRecord(d->ev[5], d->stream));
        if ((rc = gp_exchange(h, f.nfs, f.nfs, std::vector<int>(1, f.nfs), es, 1, d_gpb, d_gps, nproma))) return rc;      // TRLTOG (timed with the d2h interval)
    } else {
        ect_launch_ftinv(h, f, d_gpb, d_gps, d_fs, d_pairs, nproma);
        ECT_CUDA(cudaEventRecord(d->ev[5], d->stream));
    }
    ECT_CUDA(cudaGetLastError());
    if (d->launch_error) { d->launch_error = 0; return ECT_ERR_CUDA; }
    if ((rc = release_callbuf(d))) return rc;
    if (host) {
        if (!mode2_gp) ECT_CUDA(cudaMemcpyAsync(a->gp, dgp, (size_t)sz_gp * es, cudaMemcpyDeviceToHost, d->stream));
        else {
            if (sz_uv) ECT_CUDA(cudaMemcpyAsync(a->gpuv, dgpuv, (size_t)sz_uv * es, cudaMemcpyDeviceToHost, d->stream));
            if (sz_2) ECT_CUDA(cudaMemcpyAsync(a->gp2, dgp2, (size_t)sz_2 * es, cudaMemcpyDeviceToHost, d->stream));
            if (sz_3a) ECT_CUDA(cudaMemcpyAsync(a->gp3a, dgp3a, (size_t)sz_3a * es, cudaMemcpyDeviceToHost, d->stream));
            if (sz_3b) ECT_CUDA(cudaMemcpyAsync(a->gp3b, dgp3b, (size_t)sz_3b * es, cudaMemcpyDeviceToHost, d->stream));
        }
    }
    ECT_CUDA(cudaEventRecord(d->ev[6], d->stream));
    d->last_dir = 0;
    d->timed = true;
    if (host) ECT_CUDA(cudaStreamSynchronize(d->stream));
    return ECT_SUCCESS;
}

static int dir_trans_impl(int handle, const ect_dir_args* a, const SubView* view, int adj) {
    EctHandle* h = get_handle(handle);
    if (!h) { ect_set_error("ect_dir_trans: invalid handle %d", handle); return ECT_ERR_HANDLE; }
    if (!a) return ECT_ERR_MISSING;
    if (!h->d) { ect_set_error("ect_dir_trans: handle was set up host-only (no device state)"); return ECT_ERR_CUDA; }
    const EctHostPlan& P = h->hp;
    EctDevice* d = h->d;
    ECT_CUDA(cudaSetDevice(d->dev));
    const bool mode2 = (a->gp == nullptr);
    const int kf_uv = a->nuv;
    int nsc2 = 0, n3a = 0, n3b = 0, kf_sc = 0;
    if (kf_uv < 0 || a->nscalar < 0) { ect_set_error("ect_dir_trans: negative field count"); return ECT_ERR_BADARG; }
    if (mode2) {
        nsc2 = a->gp2 ? a->nsc2 : 0;
        n3a = a->gp3a ? a->nsc3a_lev * a->nsc3a_fld : 0;
        n3b = a->gp3b ? a->nsc3b_lev * a->nsc3b_fld : 0;
        kf_sc = nsc2 + n3a + n3b;
        if (kf_uv > 0 && !a->gpuv) { ect_set_error("ect_dir_trans: nuv > 0 but neither gp nor gpuv given"); return ECT_ERR_MISSING; }
        if (a->nscalar > 0 && kf_sc == 0) { ect_set_error("ect_dir_trans: nscalar > 0 but no grid-point input array given"); return ECT_ERR_MISSING; }
        if ((nsc2 && !a->spsc2) || (n3a && !a->spsc3a) || (n3b && !a->spsc3b)) { ect_set_error("ect_dir_trans: missing spectral output for call mode 2"); return ECT_ERR_MISSING; }
    } else {
        kf_sc = a->nscalar;
        if (kf_sc > 0 && !a->spscalar) { ect_set_error("ect_dir_trans: nscalar > 0 needs spscalar"); return ECT_ERR_MISSING; }
    }
    if (kf_uv > 0 && (!a->spvor || !a->spdiv)) { ect_set_error("ect_dir_trans: nuv > 0 needs spvor and spdiv"); return ECT_ERR_MISSING; }
    EctFieldCfg f;
    f.kf_uv = kf_uv; f.kf_sc = kf_sc;
    f.nleg = f.nfs = 2 * kf_uv + kf_sc;
    if (f.nfs == 0) return ECT_SUCCESS;
    f.cp = round_up(2 * f.nleg, ECT_CPAD);
    f.fp32 = (h->precision == ECT_PREC_SP);
    f.adj = adj;
    const int es = f.fp32 ? 4 : 8;
    const int nproma = (a->nproma > 0 && a->nproma < P.ngptot) ? a->nproma : std::max(P.ngptot, 1);
    const int ngpblks = (P.ngptot + nproma - 1) / nproma;
    int rc;
    const bool host = (a->memspace == ECT_MEM_HOST);
    // grid-point field table (trgtol_mod.F90): base pointer and block stride of every input field, order u v scalars
    auto fill_gp_table = [&](double* q_gp, double* q_uv, double* q_2, double* q_3a, double* q_3b, double** t_gpb, i64* t_gps) {
        if (!mode2) {
            for (int i = 0; i < f.nfs; ++i) { t_gpb[i] = adv(q_gp, (i64)i * nproma, es); t_gps[i] = (i64)f.nfs * nproma; }
            return;
        }
        int fi = 0;
        for (int v = 0; v < (kf_uv ? 2 : 0); ++v)
            for (int l = 0; l < kf_uv; ++l, ++fi) { t_gpb[fi] = adv(q_uv, ((i64)v * kf_uv + l) * nproma, es); t_gps[fi] = (i64)nproma * kf_uv * 2; }
        for (int j = 0; j < nsc2; ++j, ++fi) { t_gpb[fi] = adv(q_2, (i64)j * nproma, es); t_gps[fi] = (i64)nproma * nsc2; }
        for (int j3 = 0; j3 < (n3a ? a->nsc3a_fld : 0); ++j3)
            for (int l = 0; l < a->nsc3a_lev; ++l, ++fi) {
                t_gpb[fi] = adv(q_3a, ((i64)j3 * a->nsc3a_lev + l) * nproma, es); t_gps[fi] = (i64)nproma * a->nsc3a_lev * a->nsc3a_fld;
            }
        for (int j3 = 0; j3 < (n3b ? a->nsc3b_fld : 0); ++j3)
            for (int l = 0; l < a->nsc3b_lev; ++l, ++fi) {
                t_gpb[fi] = adv(q_3b, ((i64)j3 * a->nsc3b_lev + l) * nproma, es); t_gps[fi] = (i64)nproma * a->nsc3b_lev * a->nsc3b_fld;
            }
    };
    if (host) {
        const int nch = host_chunk_count(h, f.nleg);
        if (nch > 1) return dir_trans_chunked(handle, h, a, f, mode2, nsc2, n3a, n3b, nproma, ngpblks, fill_gp_table, adj);
    }
    if ((rc = ensure_work(h, f))) return rc;
    d->launches = 0;
    d->chunked_ms = -1.f;
    const i64 nsp = P.nspec2, blk = (i64)nproma * ngpblks;
    const i64 sz_gp = (i64)f.nfs * blk, sz_uv = (i64)kf_uv * 2 * blk, sz_2 = (i64)nsc2 * blk, sz_3a = (i64)n3a * blk, sz_3b = (i64)n3b * blk;
    ECT_CUDA(cudaEventRecord(d->ev[0], d->stream));
    const double *dgp = a->gp, *dgpuv = a->gpuv, *dgp2 = a->gp2, *dgp3a = a->gp3a, *dgp3b = a->gp3b;
    if (host) {
        const i64 tot = mode2 ? (sz_uv + sz_2 + sz_3a + sz_3b) : sz_gp;
        if ((rc = ensure(d->stage_gp, d->stage_gp_elems, tot, d->stream, false))) return rc;
        double* p = d->stage_gp;
        auto stage = [&](const double*& ptr, i64 n) -> int {
            if (!ptr || n == 0) return ECT_SUCCESS;
            ECT_CUDA(cudaMemcpyAsync(p, ptr, (size_t)n * es, cudaMemcpyHostToDevice, d->stream));
            ptr = p; p = adv(p, n, es);
            return ECT_SUCCESS;
        };
        if (!mode2) { if ((rc = stage(dgp, sz_gp))) return rc; }
        else {
            if ((rc = stage(dgpuv, sz_uv))) return rc;
            if ((rc = stage(dgp2, sz_2))) return rc;
            if ((rc = stage(dgp3a, sz_3a))) return rc;
            if ((rc = stage(dgp3b, sz_3b))) return rc;
        }
    }
    ECT_CUDA(cudaEventRecord(d->ev[1], d->stream));
    double *dvor = a->spvor, *ddiv = a->spdiv, *dsc = a->spscalar, *dsc2 = a->spsc2, *dsc3a = a->spsc3a, *dsc3b = a->spsc3b;
    if (host) {
        const i64 tot = (2 * (i64)kf_uv + kf_sc) * nsp;
        if ((rc = ensure(d->stage_sp, d->stage_sp_elems, tot, d->stream, false))) return rc;
        double* p = d->stage_sp;
        dvor = p; p = adv(p, kf_uv * nsp, es); ddiv = p; p = adv(p, kf_uv * nsp, es);
        if (!mode2) { dsc = p; }
        else { dsc2 = p; p = adv(p, nsc2 * nsp, es); dsc3a = p; p = adv(p, n3a * nsp, es); dsc3b = p; }
        if (adj) {     // INV_TRANSAD adds to the spectral arrays: bring their present contents first
            auto pre = [&](double* dst, const double* src, i64 n) -> int {
                if (src && n) ECT_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * es, cudaMemcpyHostToDevice, d->stream));
                return ECT_SUCCESS;
            };
            if ((rc = pre(dvor, a->spvor, kf_uv * nsp)) || (rc = pre(ddiv, a->spdiv, kf_uv * nsp))) return rc;
            if (!mode2) { if ((rc = pre(dsc, a->spscalar, kf_sc * nsp))) return rc; }
            else if ((rc = pre(dsc2, a->spsc2, nsc2 * nsp)) || (rc = pre(dsc3a, a->spsc3a, n3a * nsp)) || (rc = pre(dsc3b, a->spsc3b, n3b * nsp))) return rc;
        }
    }
    const size_t n_spec = (size_t)(2 * kf_uv + kf_sc);
    std::vector<int> groups;
    if (kf_uv) { groups.push_back(kf_uv); groups.push_back(kf_uv); }
    if (!mode2) { if (kf_sc) groups.push_back(kf_sc); }
    else {
        if (nsc2) groups.push_back(nsc2);
        for (int j3 = 0; j3 < (n3a ? a->nsc3a_fld : 0); ++j3) groups.push_back(a->nsc3a_lev);
        for (int j3 = 0; j3 < (n3b ? a->nsc3b_fld : 0); ++j3) groups.push_back(a->nsc3b_lev);
    }
    std::vector<int2> pairs = make_pairs(groups);
    f.npairs = (int)pairs.size();
    const bool gpx = P.gp_eq;       // TRGTOL is an exchange: the caller's points travel to the band buffer first
    if (gpx && (rc = gp_exchange_setup(h, f.nfs, f.nfs))) return rc;
    const size_t bytes = n_spec * sizeof(EctSpecFieldH) + (size_t)f.nfs * 2 * (sizeof(double*) + sizeof(i64)) +
                         pairs.size() * sizeof(int2) + 64;
    if ((rc = ensure_callbuf(d, bytes))) return rc;
    char* hb = (char*)d->h_callbuf;
    EctSpecFieldH* t_vor = (EctSpecFieldH*)hb;
    EctSpecFieldH* t_div = t_vor + kf_uv;
    EctSpecFieldH* t_sc = t_div + kf_uv;
    double** t_gpb = (double**)(t_sc + kf_sc);
    i64* t_gps = (i64*)(t_gpb + f.nfs);
    double** t_bandb = (double**)(t_gps + f.nfs);
    i64* t_bands = (i64*)(t_bandb + f.nfs);
    int2* t_pairs = (int2*)(t_bands + f.nfs);
    memcpy(t_pairs, pairs.data(), pairs.size() * sizeof(int2));
    for (int i = 0; i < f.nfs; ++i) { t_bandb[i] = gpx ? adv(d->gpband, (i64)i * P.ngpband, es) : nullptr; t_bands[i] = (i64)f.nfs * P.ngpband; }
    const int st_uv = view ? view->uv_stride : kf_uv, st_sc = view ? view->sc_stride : kf_sc;
    for (int j = 0; j < kf_uv; ++j) { t_vor[j] = {adv(dvor, j, es), st_uv}; t_div[j] = {adv(ddiv, j, es), st_uv}; }
    if (!mode2) {
        for (int s = 0; s < kf_sc; ++s) t_sc[s] = {adv(dsc, s, es), st_sc};
    } else {
        int s = 0;
        for (int j = 0; j < nsc2; ++j) t_sc[s++] = {adv(dsc2, j, es), nsc2};
        for (int j3 = 0; j3 < (n3a ? a->nsc3a_fld : 0); ++j3)
            for (int l = 0; l < a->nsc3a_lev; ++l) t_sc[s++] = {adv(dsc3a, (i64)j3 * a->nsc3a_lev * nsp + l, es), a->nsc3a_lev};
        for (int j3 = 0; j3 < (n3b ? a->nsc3b_fld : 0); ++j3)
            for (int l = 0; l < a->nsc3b_lev; ++l) t_sc[s++] = {adv(dsc3b, (i64)j3 * a->nsc3b_lev * nsp + l, es), a->nsc3b_lev};
    }
    fill_gp_table((double*)dgp, (double*)dgpuv, (double*)dgp2, (double*)dgp3a, (double*)dgp3b, t_gpb, t_gps);
    if ((rc = upload_callbuf(d, bytes))) return rc;
    char* db = (char*)d->callbuf;
    void* d_vor = db;
    void* d_div = db + ((char*)t_div - hb);
    void* d_sc = db + ((char*)t_sc - hb);
    double* const* d_gpb = (double* const*)(db + ((char*)t_gpb - hb));
    const i64* d_gps = (const i64*)(db + ((char*)t_gps - hb));
    const void* d_pairs = db + ((char*)t_pairs - hb);
    nvtxRangePushA("FTDIR_CTL      - Dir. Fourier transform");
    if (gpx) {
        double* const* d_bandb = (double* const*)(db + ((char*)t_bandb - hb));
        const i64* d_bands = (const i64*)(db + ((char*)t_bands - hb));
        if ((rc = gp_exchange(h, f.nfs, f.nfs, std::vector<int>(1, f.nfs), es, 0, d_gpb, d_gps, nproma))) { nvtxRangePop(); return rc; }      // TRGTOL (timed with the Fourier interval)
        if ((rc = ect_transpose_enter(h, 0))) return rc;
        ect_launch_ftdir(h, f, d_bandb, d_bands, d_pairs, std::max(P.ngpband, 1));
    } else {
        if ((rc = ect_transpose_enter(h, 0))) return rc;
        ect_launch_ftdir(h, f, d_gpb, d_gps, d_pairs, nproma);
    }
    nvtxRangePop();
    ECT_CUDA(cudaEventRecord(d->ev[2], d->stream));
    {
        EctRange r("LTDIR_CTL      - L to M transposition");
        if ((rc = ect_transpose(h, f, 0))) return rc;
        ECT_CUDA(cudaEventRecord(d->ev[3], d->stream));
    }
    {
        EctRange r("LTDIR_CTL      - Dir. Legendre transform");
        ect_launch_ledir(h, f);
        ECT_CUDA(cudaEventRecord(d->ev[4], d->stream));
        ect_launch_ltdir_epilogue(h, f, d_vor, d_div, d_sc);
        ECT_CUDA(cudaEventRecord(d->ev[5], d->stream));
    }
    ECT_CUDA(cudaGetLastError());
    if (d->launch_error) { d->launch_error = 0; return ECT_ERR_CUDA; }
    if ((rc = release_callbuf(d))) return rc;
    if (host) {
        auto back = [&](double* dst, const double* src, i64 n) -> int {
            if (!dst || n == 0) return ECT_SUCCESS;
            ECT_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * es, cudaMemcpyDeviceToHost, d->stream));
            return ECT_SUCCESS;
        };
        if ((rc = back(a->spvor, dvor, kf_uv * nsp))) return rc;
        if ((rc = back(a->spdiv, ddiv, kf_uv * nsp))) return rc;
        if (!mode2) { if ((rc = back(a->spscalar, dsc, kf_sc * nsp))) return rc; }
        else {
            if ((rc = back(a->spsc2, dsc2, nsc2 * nsp))) return rc;
            if ((rc = back(a->spsc3a, dsc3a, n3a * nsp))) return rc;
            if ((rc = back(a->spsc3b, dsc3b, n3b * nsp))) return rc;
        }
    }
    ECT_CUDA(cudaEventRecord(d->ev[6], d->stream));
    d->last_dir = 1;
    d->timed = true;
    if (host) ECT_CUDA(cudaStreamSynchronize(d->stream));
    return ECT_SUCCESS;
}

extern "C" int ect_comm_info(int handle, int* peer_memory, long long* entry_barriers) {
    EctHandle* h = get_handle(handle);
    if (!h || !h->d) return ECT_ERR_HANDLE;
    if (peer_memory) *peer_memory = h->d->p2p ? 1 : 0;
    if (entry_barriers) *entry_barriers = h->d->entry_barriers;
    return ECT_SUCCESS;
}

extern "C" int ect_synchronize(int handle) {
    EctHandle* h = get_handle(handle);
    if (!h || !h->d) return ECT_ERR_HANDLE;
    ECT_CUDA(cudaStreamSynchronize(h->d->stream));
    return ECT_SUCCESS;
}

extern "C" int ect_get_timings(int handle, ect_timings* t) {
    EctHandle* h = get_handle(handle);
    if (!h || !h->d) return ECT_ERR_HANDLE;
    if (!t) return ECT_ERR_MISSING;
    EctDevice* d = h->d;
    memset(t, 0, sizeof(*t));
    if (!d->timed) return ECT_SUCCESS;
    ECT_CUDA(cudaStreamSynchronize(d->stream));
    if (d->chunked_ms >= 0.f) {       // chunked host call: copies and stages of different chunks overlap, only the total is defined
        t->total = d->chunked_ms;
        t->launches = d->launches;
        return ECT_SUCCESS;
    }
    float ms[6];
    for (int i = 0; i < 6; ++i) ECT_CUDA(cudaEventElapsedTime(&ms[i], d->ev[i], d->ev[i + 1]));
    ECT_CUDA(cudaEventElapsedTime(&t->total, d->ev[0], d->ev[6]));
    if (d->last_dir == 0) {
        t->h2d = ms[0]; t->prologue = ms[1]; t->legendre = ms[2]; t->transpose = ms[3]; t->fourier = ms[4]; t->d2h = ms[5];
    } else {
        t->h2d = ms[0]; t->fourier = ms[1]; t->transpose = ms[2]; t->legendre = ms[3]; t->epilogue = ms[4]; t->d2h = ms[5];
    }
    t->launches = d->launches;
    return ECT_SUCCESS;
}

// ---------------------------------------------------------------------------------------
// SPECNORM: cpu/internal/spnormd_mod.F90:36-51, spnorm_ctl_mod.F90:56-57
// ---------------------------------------------------------------------------------------
// one thread per (local wavenumber, field): the n-sum in the reference's order (spnormd_mod.F90:36-51); the sums over m
// are added on the host in ascending m (spnorm_ctl_mod.F90:56-57), so the norm does not depend on the decomposition
template <bool FP32>
__global__ void k_specnorm(const double* __restrict__ sp, int nfld, int T, const EctLegM* __restrict__ legm,
                           const int* __restrict__ nasm0, const double* __restrict__ pmet, double* __restrict__ zgm) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfld) return;
    const int m = legm[blockIdx.y].m, base = nasm0[blockIdx.y];
    double s = 0.0;
    for (int n = m; n <= T; ++n) {
        const long long i = base + 2 * (n - m);
        const double re = FP32 ? (double)reinterpret_cast<const float*>(sp)[i * nfld + f] : sp[i * nfld + f];
        const double w = pmet ? pmet[n] : 1.0;
        if (m == 0) s += w * re * re;
        else {
            const double im = FP32 ? (double)reinterpret_cast<const float*>(sp)[(i + 1) * nfld + f] : sp[(i + 1) * nfld + f];
            s += 2.0 * w * (re * re + im * im);
        }
    }
    zgm[(long long)f * (T + 1) + m] = s;
}

extern "C" int ect_specnorm_met(int handle, const double* spec, int nfld, int memspace, const double* pmet, double* norms) {
    if (int g = vset_guard(handle, "ect_specnorm")) return g;
    EctHandle* h = get_handle(handle);
    if (!h || !h->d) { ect_set_error("ect_specnorm: invalid handle"); return ECT_ERR_HANDLE; }
    if (!spec || !norms || nfld <= 0) return ECT_ERR_MISSING;
    const EctHostPlan& P = h->hp;
    EctDevice* d = h->d;
    ECT_CUDA(cudaSetDevice(d->dev));
    const double* dsp = spec;
    int rc;
    if (memspace == ECT_MEM_HOST) {
        if ((rc = ensure(d->stage_sp, d->stage_sp_elems, (i64)nfld * P.nspec2, d->stream, false))) return rc;
        ECT_CUDA(cudaMemcpyAsync(d->stage_sp, spec, (size_t)nfld * P.nspec2 * (h->precision == ECT_PREC_SP ? 4 : 8), cudaMemcpyHostToDevice, d->stream));
        dsp = d->stage_sp;
    }
    const int nm = P.nsmax + 1;
    const int need = nfld * nm + nm;
    if (d->normbuf_n < need) {
        if (d->normbuf) cudaFree(d->normbuf);
        ECT_CUDA(cudaMalloc(&d->normbuf, (size_t)need * sizeof(double)));
        d->normbuf_n = need;
    }
    double* zgm = d->normbuf; double* dmet = nullptr;
    ECT_CUDA(cudaMemsetAsync(zgm, 0, (size_t)nfld * nm * sizeof(double), d->stream));
    if (pmet) {
        dmet = d->normbuf + (size_t)nfld * nm;
        ECT_CUDA(cudaMemcpyAsync(dmet, pmet, nm * sizeof(double), cudaMemcpyHostToDevice, d->stream));
    }
    if (P.nump > 0) {
        dim3 grid((nfld + 127) / 128, P.nump);
        if (h->precision == ECT_PREC_SP) k_specnorm<true><<<grid, 128, 0, d->stream>>>(dsp, nfld, P.nsmax, d->legm, d->nasm0, dmet, zgm);
        else k_specnorm<false><<<grid, 128, 0, d->stream>>>(dsp, nfld, P.nsmax, d->legm, d->nasm0, dmet, zgm);
    }
    // every (m, field) sum lives on exactly one rank: adding zeros is exact
    if (P.nranks > 1) ECT_NCCL(ncclAllReduce(zgm, zgm, (size_t)nfld * nm, ncclDouble, ncclSum, (ncclComm_t)d->comm, d->stream));
    std::vector<double> hz((size_t)nfld * nm);
    ECT_CUDA(cudaMemcpyAsync(hz.data(), zgm, hz.size() * sizeof(double), cudaMemcpyDeviceToHost, d->stream));
    ECT_CUDA(cudaStreamSynchronize(d->stream));
    for (int f = 0; f < nfld; ++f) {
        double t = 0.0;
        for (int m = 0; m < nm; ++m) t += hz[(size_t)f * nm + m];       // PNORM = SUM(ZGM, DIM=2)
        norms[f] = sqrt(t);
    }
    return ECT_SUCCESS;
}
extern "C" int ect_specnorm(int handle, const double* spec, int nfld, int memspace, double* norms) {
    return ect_specnorm_met(handle, spec, nfld, memspace, nullptr, norms);
}

// SPECNORM with KVSET (specnorm.h; spnorm_ctl_mod.F90:44-57 + spnormc_mod.F90): spec holds the nfld fields of this task's
// V-set, kvset (1-based V-set per GLOBAL field, nfld_g entries) says which global fields they are; norms: nfld_g values,
// on every task.  Works for NPRTRV = 1 too (kvset all ones).
extern "C" int ect_specnorm_vset(int handle, const double* spec, int nfld, int memspace, const double* pmet, const int* kvset,
                                 int nfld_g, double* norms) {
    EctHandle* h = get_handle(handle);
    if (!h || !h->d) { ect_set_error("ect_specnorm: invalid handle"); return ECT_ERR_HANDLE; }
    if (!norms || !kvset || nfld_g <= 0 || nfld < 0 || (nfld > 0 && !spec)) return ECT_ERR_MISSING;
    const EctHostPlan& P = h->hp;
    EctDevice* d = h->d;
    const int V = h->vs.V, me = h->vs.v + 1, nm = P.nsmax + 1;
    std::vector<int> mine;                     // global index of my local fields, in order
    for (int f = 0; f < nfld_g; ++f) {
        if (kvset[f] < 1 || kvset[f] > V) { ect_set_error("SPECNORM: KVSET TOO LONG OR CONTAINS VALUES OUTSIDE RANGE"); return ECT_ERR_BADARG; }
        if (kvset[f] == me) mine.push_back(f);
    }
    if ((int)mine.size() != nfld) { ect_set_error("SPECNORM: PSPEC holds %d fields, KVSET gives this V-set %d", nfld, (int)mine.size()); return ECT_ERR_BADARG; }
    ECT_CUDA(cudaSetDevice(d->dev));
    const int es = h->precision == ECT_PREC_SP ? 4 : 8;
    const double* dsp = spec;
    int rc;
    if (memspace == ECT_MEM_HOST && nfld > 0) {
        if ((rc = ensure(d->stage_sp, d->stage_sp_elems, (i64)nfld * P.nspec2, d->stream, false))) return rc;
        ECT_CUDA(cudaMemcpyAsync(d->stage_sp, spec, (size_t)nfld * P.nspec2 * es, cudaMemcpyHostToDevice, d->stream));
        dsp = d->stage_sp;
    }
    double* w = nullptr;        // [local sums nfld x nm][global sums nfld_g x nm][metric nm]
    ECT_CUDA(cudaMalloc(&w, ((size_t)(nfld + nfld_g) * nm + nm) * sizeof(double)));
    double* zl = w; double* zg = w + (size_t)nfld * nm; double* dmet = nullptr;
    ECT_CUDA(cudaMemsetAsync(w, 0, (size_t)(nfld + nfld_g) * nm * sizeof(double), d->stream));
    if (pmet) { dmet = zg + (size_t)nfld_g * nm; ECT_CUDA(cudaMemcpyAsync(dmet, pmet, nm * sizeof(double), cudaMemcpyHostToDevice, d->stream)); }
    if (P.nump > 0 && nfld > 0) {
        dim3 grid((nfld + 127) / 128, P.nump);
        if (es == 4) k_specnorm<true><<<grid, 128, 0, d->stream>>>(dsp, nfld, P.nsmax, d->legm, d->nasm0, dmet, zl);
        else k_specnorm<false><<<grid, 128, 0, d->stream>>>(dsp, nfld, P.nsmax, d->legm, d->nasm0, dmet, zl);
    }
    for (int i = 0; i < nfld; ++i)      // local field i is global field mine[i]
        ECT_CUDA(cudaMemcpyAsync(zg + (size_t)mine[i] * nm, zl + (size_t)i * nm, nm * sizeof(double), cudaMemcpyDeviceToDevice, d->stream));
    // every (m, global field) sum lives on exactly one task: adding zeros is exact
    ncclComm_t comm = (ncclComm_t)(V > 1 ? d->comm_world : d->comm);
    if (comm) ECT_NCCL(ncclAllReduce(zg, zg, (size_t)nfld_g * nm, ncclDouble, ncclSum, comm, d->stream));
    std::vector<double> hz((size_t)nfld_g * nm);
    ECT_CUDA(cudaMemcpyAsync(hz.data(), zg, hz.size() * sizeof(double), cudaMemcpyDeviceToHost, d->stream));
    ECT_CUDA(cudaStreamSynchronize(d->stream));
    ECT_CUDA(cudaFree(w));
    for (int f = 0; f < nfld_g; ++f) {
        double t = 0.0;
        for (int m = 0; m < nm; ++m) t += hz[(size_t)f * nm + m];
        norms[f] = sqrt(t);
    }
    return ECT_SUCCESS;
}

// ---------------------------------------------------------------------------------------
// test / debug access (not part of the reference API)
// ---------------------------------------------------------------------------------------
extern "C" int ect_debug_get_table(int handle, int ml, int par, double* out, long long cap) {
    EctHandle* h = get_handle(handle);
    if (!h || !h->d) return ECT_ERR_HANDLE;
    return ect_legendre_get_table(h, ml, par, out, cap);
}

// ---------------------------------------------------------------------------------------
// FP64 peak microbenchmarks (roofline denominators for the Legendre stage)
// ---------------------------------------------------------------------------------------
__global__ void k_peak_dmma(double* out, int iters) {
    double c[8][2];
    for (int i = 0; i < 8; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_peak_dfma(double* out, int iters) {
    double c[16];
    for (int i = 0; i < 16; ++i) c[i] = i * 1e-3;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: even warps issue DMMA, odd warps DFMA -- do the two share one pipe?  (aggregate flops reported)
__global__ void k_peak_mixed(double* out, int iters) {
    double c[16];
    for (int i = 0; i < 16; ++i) c[i] = i * 1e-3;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    if ((threadIdx.x >> 5) & 1) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[2 * i]), "+d"(c[2 * i + 1]) : "d"(a), "d"(b));
        }
    }
    double s = 0.0;
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// legacy tensor path, TF32 m16n8k8 with fp32 accumulators (sp contraction: 3 x TF32 split products)
__global__ void k_peak_tf32(double* out, int iters) {
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    const unsigned a0 = __float_as_uint(1.0f + threadIdx.x * 1e-3f), b0 = __float_as_uint(1.0f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a0), "r"(a0), "r"(a0), "r"(a0), "r"(b0), "r"(b0));
    }
    float s = 0.f;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_peak_ffma(double* out, int iters) {
    float c[16];
    for (int i = 0; i < 16; ++i) c[i] = i * 1e-3f;
    const float a = 1.0f + threadIdx.x * 1e-9f, b = 1e-9f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fmaf(c[i], a, b);
    }
    float s = 0.f;
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int ect_measure_fp64_peak(int which, double* tflops) {
    if (!tflops) return ECT_ERR_MISSING;
    int dev = 0, sms = 0;
    ECT_CUDA(cudaGetDevice(&dev));
    ECT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* out = nullptr;
    ECT_CUDA(cudaMalloc(&out, (size_t)sms * 8 * 256 * sizeof(double)));
    cudaEvent_t e0, e1;
    ECT_CUDA(cudaEventCreate(&e0));
    ECT_CUDA(cudaEventCreate(&e1));
    const int blocks = sms * 4, threads = 256, iters = 20000;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        ECT_CUDA(cudaEventRecord(e0));
        if (which == 0) k_peak_dmma<<<blocks, threads>>>(out, iters);
        else if (which == 2) k_peak_mixed<<<blocks, threads>>>(out, iters);
        else if (which == 3) k_peak_tf32<<<blocks, threads>>>(out, iters);
        else if (which == 4) k_peak_ffma<<<blocks, threads>>>(out, iters);
        else k_peak_dfma<<<blocks, threads>>>(out, iters);
        ECT_CUDA(cudaEventRecord(e1));
        ECT_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        ECT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double flops;
        if (which == 0) flops = (double)blocks * (threads / 32) * (double)iters * 8.0 * (2.0 * 8 * 8 * 4);
        else if (which == 2) flops = (double)blocks * (threads / 64) * (double)iters * (8.0 * (2.0 * 8 * 8 * 4) + 32 * 16.0 * 2.0);
        else if (which == 3) flops = (double)blocks * (threads / 32) * (double)iters * 8.0 * (2.0 * 16 * 8 * 8);
        else flops = (double)blocks * threads * (double)iters * 16.0 * 2.0;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    ECT_CUDA(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return ECT_SUCCESS;
}

// ---------------------------------------------------------------------------------------
// GATH_GRID / DIST_GRID / GATH_SPEC / DIST_SPEC: global <-> distributed arrays on host memory.
// Reference: cpu/internal/gath_grid_ctl_mod.F90, dist_grid_ctl_mod.F90, gath_spec_control_mod.F90,
// dist_spec_control_mod.F90 (interfaces include/ectrans/gath_grid.h:12-60 ...).  Field f of the global array
// lives on rank kto[f] / kfrom[f] (0-based); a rank's global array holds the fields it owns, in order.
// Grid points: rank r owns the global range of its latitude band; spectral: global order is m ascending,
// n ascending (IASM0G, gath_spec_control_mod.F90:111-114) whatever the wave distribution.
// The pieces travel as NCCL send/recv on device staging buffers; with one rank it is a host-side reorder.
// ---------------------------------------------------------------------------------------
struct PieceMap { std::vector<i64> off, cnt; };

static PieceMap grid_pieces(const EctHostPlan& P) {
    PieceMap g; g.off.assign(P.nranks, 0); g.cnt.assign(P.nranks, 0);
    i64 pos = 0;
    for (int r = 0; r < P.nranks; ++r) {
        i64 n = 0;
        if (P.gp_eq) for (int i = P.gp_all_seg0[r]; i < P.gp_all_seg0[r + 1]; ++i) n += P.gp_all_segs[i].count;   // scattered: see gpos
        else for (int j = 0; j < P.lat_count[r]; ++j) n += P.nloen[P.lat_first[r] + j];
        g.off[r] = pos; g.cnt[r] = n; pos += n;
    }
    return g;
}
static PieceMap spec_pieces(const EctHostPlan& P) {
    PieceMap g; g.off.assign(P.nranks, 0); g.cnt.assign(P.nranks, 0);
    i64 pos = 0;
    for (int r = 0; r < P.nranks; ++r) {
        i64 n = 0;
        for (int m : P.ms_of[r]) n += 2 * (P.nsmax - m + 1);
        g.off[r] = pos; g.cnt[r] = n; pos += n;
    }
    return g;
}

// exchange of per-rank pieces: gather (everybody -> owners) or scatter (owners -> everybody).
// sendbytes[t] bytes at sendoff[t] of h_send go to rank t; recvbytes[s] bytes from rank s land at recvoff[s] of h_recv.
static int exchange_host(EctHandle* h, const char* h_send, const std::vector<i64>& sendoff, const std::vector<i64>& sendbytes,
                         char* h_recv, const std::vector<i64>& recvoff, const std::vector<i64>& recvbytes) {
    const EctHostPlan& P = h->hp;
    EctDevice* d = h->d;
    i64 stot = 0, rtot = 0;
    for (int r = 0; r < P.nranks; ++r) { stot = std::max(stot, sendoff[r] + sendbytes[r]); rtot = std::max(rtot, recvoff[r] + recvbytes[r]); }
    if (P.nranks == 1) {
        if (sendbytes[0] != recvbytes[0]) { ect_set_error("gath/dist: inconsistent piece sizes"); return ECT_ERR_GENERIC; }
        memcpy(h_recv + recvoff[0], h_send + sendoff[0], (size_t)sendbytes[0]);
        return ECT_SUCCESS;
    }
    int rc;
    if ((rc = ensure(d->stage_sp, d->stage_sp_elems, (stot + 7) / 8 + 1, d->stream, false))) return rc;
    if ((rc = ensure(d->stage_gp, d->stage_gp_elems, (rtot + 7) / 8 + 1, d->stream, false))) return rc;
    char* ds = (char*)d->stage_sp; char* dr = (char*)d->stage_gp;
    if (stot) ECT_CUDA(cudaMemcpyAsync(ds, h_send, (size_t)stot, cudaMemcpyHostToDevice, d->stream));
    ncclComm_t comm = (ncclComm_t)d->comm;
    ECT_NCCL(ncclGroupStart());
    for (int r = 0; r < P.nranks; ++r) {
        if (sendbytes[r]) ECT_NCCL(ncclSend(ds + sendoff[r], (size_t)sendbytes[r], ncclChar, r, comm, d->stream));
        if (recvbytes[r]) ECT_NCCL(ncclRecv(dr + recvoff[r], (size_t)recvbytes[r], ncclChar, r, comm, d->stream));
    }
    ECT_NCCL(ncclGroupEnd());
    if (rtot) ECT_CUDA(cudaMemcpyAsync(h_recv, dr, (size_t)rtot, cudaMemcpyDeviceToHost, d->stream));
    ECT_CUDA(cudaStreamSynchronize(d->stream));
    return ECT_SUCCESS;
}

static int check_owner(const EctHostPlan& P, const int* own, int nfld, const char* who) {
    if (nfld < 0 || (nfld > 0 && !own)) { ect_set_error("%s: owner list missing", who); return ECT_ERR_MISSING; }
    for (int f = 0; f < nfld; ++f)
        if (own[f] < 0 || own[f] >= P.nranks) { ect_set_error("%s: field %d owner %d outside 0..%d", who, f, own[f], P.nranks - 1); return ECT_ERR_BADARG; }
    return ECT_SUCCESS;
}

// gather = true : local (blocked grid / local spectral) -> global arrays on the owners
// gather = false: global arrays on the owners -> local
static int gath_dist(int handle, bool grid, bool gather, void* v_local, void* v_global, int nfld, int nproma_in, const int* own) {
    if (int g = vset_guard(handle, "gath/dist")) return g;
    EctHandle* h = get_handle(handle);
    if (!h) { ect_set_error("gath/dist: invalid handle %d", handle); return ECT_ERR_HANDLE; }
    const EctHostPlan& P = h->hp;
    if (!h->d && P.nranks > 1) { ect_set_error("gath/dist: handle was set up host-only"); return ECT_ERR_CUDA; }
    const char* who = grid ? (gather ? "ect_gath_grid" : "ect_dist_grid") : (gather ? "ect_gath_spec" : "ect_dist_spec");
    int rc;
    if ((rc = check_owner(P, own, nfld, who))) return rc;
    if (nfld == 0) return ECT_SUCCESS;
    if (h->d) ECT_CUDA(cudaSetDevice(h->d->dev));
    const int es = h->precision == ECT_PREC_SP ? 4 : 8;
    const PieceMap pm = grid ? grid_pieces(P) : spec_pieces(P);
    const i64 nloc = pm.cnt[P.rank];
    const i64 nglob = grid ? (i64)P.ngptotg : (i64)P.nspec2_g;
    std::vector<int> nown(P.nranks, 0), slot(nfld, 0);      // fields per owner, position of f among its owner's fields
    for (int f = 0; f < nfld; ++f) slot[f] = nown[own[f]]++;
    const int mine = nown[P.rank];
    if (nloc > 0 && !v_local) { ect_set_error("%s: local array missing", who); return ECT_ERR_MISSING; }
    if (mine > 0 && !v_global) { ect_set_error("%s: this rank owns %d field(s) but the global array is missing", who, mine); return ECT_ERR_MISSING; }
    char* loc = (char*)v_local; char* glob = (char*)v_global;
    const int nproma = (nproma_in > 0 && nproma_in < P.ngptot) ? nproma_in : std::max(P.ngptot, 1);
    // "wire" layout between rank s (data of its piece) and owner t: grid (fields of t, n_s) ; spectral (n_s, fields of t)
    // local-side buffer: for every owner t a block of nown[t] * nloc elements; owner-side buffer: for every source s a block
    // of mine * cnt[s] elements
    std::vector<i64> loff(P.nranks), lbytes(P.nranks), goff(P.nranks), gbytes(P.nranks);
    i64 lp = 0, gp = 0;
    for (int r = 0; r < P.nranks; ++r) {
        loff[r] = lp; lbytes[r] = (i64)nown[r] * nloc * es; lp += lbytes[r];
        goff[r] = gp; gbytes[r] = (i64)mine * pm.cnt[r] * es; gp += gbytes[r];
    }
    std::vector<char> lbuf((size_t)lp + 8), gbuf((size_t)gp + 8);
    // element (field f, local index i) in the local-side buffer / in the caller's local array
    auto lwire = [&](int f, i64 i) -> char* {
        const int t = own[f];
        const i64 e = grid ? (i64)slot[f] * nloc + i : i * nown[t] + slot[f];
        return lbuf.data() + loff[t] + e * es;
    };
    auto lcaller = [&](int f, i64 i) -> char* {
        if (!grid) return loc + (i * nfld + f) * es;                    // PSPEC(nfld, nspec2)
        const i64 b = i / nproma, k = i - b * nproma;                  // PGP(nproma, nfld, ngpblks)
        return loc + ((b * nfld + f) * nproma + k) * es;
    };
    // element (owned field slot q, source rank s, index i within the piece of s) on the owner side
    auto gwire = [&](int q, int s, i64 i) -> char* {
        const i64 e = grid ? (i64)q * pm.cnt[s] + i : i * mine + q;
        return gbuf.data() + goff[s] + e * es;
    };
    // global position of element i of the piece of rank s
    std::vector<i64> gpos;      // spectral only: global index of every local index, per source rank, concatenated
    std::vector<i64> gpos0(P.nranks + 1, 0);
    if (!grid) {
        gpos.reserve((size_t)nglob);
        std::vector<i64> iasm0g(P.nsmax + 2, 0);
        for (int m = 0; m <= P.nsmax; ++m) iasm0g[m + 1] = iasm0g[m] + 2 * (P.nsmax - m + 1);
        for (int s = 0; s < P.nranks; ++s) {
            gpos0[s] = (i64)gpos.size();
            for (int m : P.ms_of[s]) for (int i = 0; i < 2 * (P.nsmax - m + 1); ++i) gpos.push_back(iasm0g[m] + i);
        }
        gpos0[P.nranks] = (i64)gpos.size();
    } else if (P.gp_eq) {       // eq_regions partition: a task's points are pieces of latitudes, scattered in the global order
        gpos.reserve((size_t)nglob);
        std::vector<i64> latoff(P.ndgl + 1, 0);
        for (int j = 0; j < P.ndgl; ++j) latoff[j + 1] = latoff[j] + P.nloen[j];
        for (int s = 0; s < P.nranks; ++s) {
            gpos0[s] = (i64)gpos.size();
            for (int i = P.gp_all_seg0[s]; i < P.gp_all_seg0[s + 1]; ++i)
                for (int j = 0; j < P.gp_all_segs[i].count; ++j) gpos.push_back(latoff[P.gp_all_segs[i].lat] + P.gp_all_segs[i].first + j);
        }
        gpos0[P.nranks] = (i64)gpos.size();
    }
    auto gcaller = [&](int q, int s, i64 i) -> char* {
        if (grid && P.gp_eq) return glob + ((i64)q * nglob + gpos[gpos0[s] + i]) * es;
        if (grid) return glob + ((i64)q * nglob + pm.off[s] + i) * es;   // PGPG(ngptotg, nfld)
        return glob + (gpos[gpos0[s] + i] * mine + q) * es;              // PSPECG(nfld, nspec2g)
    };
    if (gather) {
        for (int f = 0; f < nfld; ++f) for (i64 i = 0; i < nloc; ++i) memcpy(lwire(f, i), lcaller(f, i), es);
        if ((rc = exchange_host(h, lbuf.data(), loff, lbytes, gbuf.data(), goff, gbytes))) return rc;
        for (int q = 0; q < mine; ++q) for (int s = 0; s < P.nranks; ++s) for (i64 i = 0; i < pm.cnt[s]; ++i) memcpy(gcaller(q, s, i), gwire(q, s, i), es);
        if (!grid && mine > 0) {
            // LDZA0IP: imaginary parts of the zonal coefficients are zero in the global array (gath_spec_control_mod.F90:178-184)
            for (int n = 0; n <= P.nsmax; ++n) for (int q = 0; q < mine; ++q) memset(glob + ((i64)(2 * n + 1) * mine + q) * es, 0, es);
        }
    } else {
        for (int q = 0; q < mine; ++q) for (int s = 0; s < P.nranks; ++s) for (i64 i = 0; i < pm.cnt[s]; ++i) memcpy(gwire(q, s, i), gcaller(q, s, i), es);
        if ((rc = exchange_host(h, gbuf.data(), goff, gbytes, lbuf.data(), loff, lbytes))) return rc;
        for (int f = 0; f < nfld; ++f) for (i64 i = 0; i < nloc; ++i) memcpy(lcaller(f, i), lwire(f, i), es);
    }
    return ECT_SUCCESS;
}

extern "C" int ect_gath_grid(int handle, const void* gp_local, int nfld, int nproma, const int* kto, void* gp_global) {
    return gath_dist(handle, true, true, (void*)gp_local, gp_global, nfld, nproma, kto);
}
extern "C" int ect_dist_grid(int handle, const void* gp_global, int nfld, int nproma, const int* kfrom, void* gp_local) {
    return gath_dist(handle, true, false, gp_local, (void*)gp_global, nfld, nproma, kfrom);
}
extern "C" int ect_gath_spec(int handle, const void* sp_local, int nfld, const int* kto, void* sp_global) {
    return gath_dist(handle, false, true, (void*)sp_local, sp_global, nfld, 0, kto);
}
extern "C" int ect_dist_spec(int handle, const void* sp_global, int nfld, const int* kfrom, void* sp_local) {
    return gath_dist(handle, false, false, sp_local, (void*)sp_global, nfld, 0, kfrom);
}

// ---------------------------------------------------------------------------------------
// GPNORM_TRANS: cpu/external/gpnorm_trans.F90, cpu/internal/gpnorm_trans_ctl_mod.F90:170-215 (per-latitude sums,
// weight RW(lat) / NLOEN(lat)), :422-428 (latitudes added in global order, which makes the average independent
// of the decomposition).  One CTA per (local latitude, field); the order of the additions inside a latitude
// depends only on its length.
// ---------------------------------------------------------------------------------------
template <bool FP32>
__global__ void k_gpnorm_lat(const void* __restrict__ gp, int nfld, int nproma, const int* __restrict__ gpoff,
                             const int* __restrict__ nloen_loc, const double* __restrict__ rw_loc, int lat0, int ndgl,
                             double* __restrict__ aveg, double* __restrict__ mn, double* __restrict__ mx) {
    const int l = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const int n = nloen_loc[l], g0 = gpoff[l];
    double s = 0.0, lo = INFINITY, hi = -INFINITY;
    for (int j = tid; j < n; j += blockDim.x) {
        const int g = g0 + j, blk = g / nproma;
        const long long idx = ((long long)blk * nfld + f) * nproma + (g - blk * nproma);
        const double v = FP32 ? (double)reinterpret_cast<const float*>(gp)[idx] : reinterpret_cast<const double*>(gp)[idx];
        s += v; lo = fmin(lo, v); hi = fmax(hi, v);
    }
    __shared__ double ss[256], sl[256], sh[256];
    ss[tid] = s; sl[tid] = lo; sh[tid] = hi;
    __syncthreads();
    for (int w = blockDim.x >> 1; w > 0; w >>= 1) {
        if (tid < w) { ss[tid] += ss[tid + w]; sl[tid] = fmin(sl[tid], sl[tid + w]); sh[tid] = fmax(sh[tid], sh[tid + w]); }
        __syncthreads();
    }
    if (tid == 0) {
        aveg[(long long)f * ndgl + lat0 + l] = ss[0] * rw_loc[l] / (double)n;
        mn[(long long)f * gridDim.x + l] = sl[0];
        mx[(long long)f * gridDim.x + l] = sh[0];
    }
}
// per field: min / max over the local latitudes, then (all ranks hold aveg after the all-reduce) the sum over latitudes
__global__ void k_gpnorm_fold(int nfld, int nlat, const double* __restrict__ mn, const double* __restrict__ mx, double* __restrict__ out /* [2][nfld] */) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfld) return;
    double lo = INFINITY, hi = -INFINITY;
    for (int l = 0; l < nlat; ++l) { lo = fmin(lo, mn[(long long)f * nlat + l]); hi = fmax(hi, mx[(long long)f * nlat + l]); }
    out[f] = lo; out[nfld + f] = hi;
}

extern "C" int ect_gpnorm_trans(int handle, const void* gp, int nfld, int nproma, int memspace, double* ave, double* pmin,
                                double* pmax, int ave_only) {
    if (int g = vset_guard(handle, "ect_gpnorm_trans")) return g;
    EctHandle* h = get_handle(handle);
    if (!h || !h->d) { ect_set_error("ect_gpnorm_trans: invalid handle"); return ECT_ERR_HANDLE; }
    if (!gp || !ave || !pmin || !pmax || nfld <= 0) return ECT_ERR_MISSING;
    const EctHostPlan& P = h->hp;
    EctDevice* d = h->d;
    ECT_CUDA(cudaSetDevice(d->dev));
    if (nproma <= 0) nproma = std::max(P.ngptot, 1);
    const int ngpblks = (P.ngptot + nproma - 1) / nproma;
    const int es = h->precision == ECT_PREC_SP ? 4 : 8;
    const void* dgp = gp;
    int rc;
    if (memspace == ECT_MEM_HOST) {
        const i64 elems = (i64)ngpblks * nfld * nproma;
        if ((rc = ensure(d->stage_gp, d->stage_gp_elems, elems, d->stream, false))) return rc;
        ECT_CUDA(cudaMemcpyAsync(d->stage_gp, gp, (size_t)elems * es, cudaMemcpyHostToDevice, d->stream));
        dgp = d->stage_gp;
    }
    int nproma_k = nproma;
    if (P.gp_eq) {
        // like the reference (gpnorm_trans_ctl_mod.F90:166-168): TRGTOL first, then whole latitudes are summed by their band owner
        if ((rc = gp_exchange_setup(h, nfld, nfld))) return rc;
        std::vector<double*> hb(nfld); std::vector<i64> hs(nfld, (i64)nfld * nproma);
        for (int f = 0; f < nfld; ++f) hb[f] = (double*)((char*)dgp + (size_t)f * nproma * es);
        char* tab = nullptr;
        ECT_CUDA(cudaMalloc(&tab, nfld * (sizeof(double*) + sizeof(i64))));
        ECT_CUDA(cudaMemcpyAsync(tab, hb.data(), nfld * sizeof(double*), cudaMemcpyHostToDevice, d->stream));
        ECT_CUDA(cudaMemcpyAsync(tab + nfld * sizeof(double*), hs.data(), nfld * sizeof(i64), cudaMemcpyHostToDevice, d->stream));
        rc = gp_exchange(h, nfld, nfld, std::vector<int>(1, nfld), es, 0, (double* const*)tab, (const i64*)(tab + nfld * sizeof(double*)), nproma);
        ECT_CUDA(cudaStreamSynchronize(d->stream));
        ECT_CUDA(cudaFree(tab));
        if (rc) return rc;
        dgp = d->gpband; nproma_k = std::max(P.ngpband, 1);
    }
    const i64 nwork = (i64)nfld * P.ndgl + 2 * (i64)nfld * std::max(P.nlat, 1) + 2 * (i64)nfld;
    double* w = nullptr;
    ECT_CUDA(cudaMalloc(&w, (size_t)nwork * sizeof(double)));
    double* aveg = w; double* mn = aveg + (i64)nfld * P.ndgl; double* mx = mn + (i64)nfld * std::max(P.nlat, 1);
    double* fold = mx + (i64)nfld * std::max(P.nlat, 1);
    ECT_CUDA(cudaMemsetAsync(aveg, 0, (size_t)nfld * P.ndgl * sizeof(double), d->stream));
    if (P.nlat > 0) {
        dim3 grid(P.nlat, nfld);
        if (es == 4) k_gpnorm_lat<true><<<grid, 256, 0, d->stream>>>(dgp, nfld, nproma_k, d->gpoff, d->nloen, d->rw_loc, P.lat0, P.ndgl, aveg, mn, mx);
        else k_gpnorm_lat<false><<<grid, 256, 0, d->stream>>>(dgp, nfld, nproma_k, d->gpoff, d->nloen, d->rw_loc, P.lat0, P.ndgl, aveg, mn, mx);
    }
    k_gpnorm_fold<<<(nfld + 127) / 128, 128, 0, d->stream>>>(nfld, P.nlat, mn, mx, fold);
    if (ave_only) {      // LDAVE_ONLY: PMIN / PMAX already hold the local extrema (gpnorm_trans_ctl_mod.F90:243-245)
        std::vector<double> loc(2 * (size_t)nfld);
        for (int f = 0; f < nfld; ++f) { loc[f] = pmin[f]; loc[nfld + f] = pmax[f]; }
        ECT_CUDA(cudaMemcpyAsync(fold, loc.data(), loc.size() * sizeof(double), cudaMemcpyHostToDevice, d->stream));
        ECT_CUDA(cudaStreamSynchronize(d->stream));
    }
    if (P.nranks > 1) {
        // every (latitude, field) sum lives on exactly one rank: adding zeros is exact
        ECT_NCCL(ncclAllReduce(aveg, aveg, (size_t)nfld * P.ndgl, ncclDouble, ncclSum, (ncclComm_t)d->comm, d->stream));
        ECT_NCCL(ncclAllReduce(fold, fold, nfld, ncclDouble, ncclMin, (ncclComm_t)d->comm, d->stream));
        ECT_NCCL(ncclAllReduce(fold + nfld, fold + nfld, nfld, ncclDouble, ncclMax, (ncclComm_t)d->comm, d->stream));
    }
    std::vector<double> hav((size_t)nfld * P.ndgl), hf(2 * (size_t)nfld);
    ECT_CUDA(cudaMemcpyAsync(hav.data(), aveg, hav.size() * sizeof(double), cudaMemcpyDeviceToHost, d->stream));
    ECT_CUDA(cudaMemcpyAsync(hf.data(), fold, hf.size() * sizeof(double), cudaMemcpyDeviceToHost, d->stream));
    ECT_CUDA(cudaStreamSynchronize(d->stream));
    ECT_CUDA(cudaFree(w));
    for (int f = 0; f < nfld; ++f) {
        double s = 0.0;
        for (int g = 0; g < P.ndgl; ++g) s += hav[(size_t)f * P.ndgl + g];     // PAVE(:) = PAVE(:) + ZAVEG(JGL,:), JGL = 1..NDGL
        ave[f] = s; pmin[f] = hf[f]; pmax[f] = hf[nfld + f];
    }
    d->launches += 2;
    return ECT_SUCCESS;
}

// ---------------------------------------------------------------------------------------
// VORDIV_TO_UV: cpu/external/vordiv_to_uv.F90, cpu/internal/vd2uv_mod.F90:86-112.  handle > 0: the wavenumbers of
// that handle's task, its precision and stream; handle == 0: one task holding every m of truncation nsmax in
// double precision (the reference builds a temporary LDSPSETUPONLY resolution for exactly this).
// ---------------------------------------------------------------------------------------
int ect_launch_vd2uv(const void* vor, const void* div, void* u, void* v, int nfld, int nsmax, int nump,
                     const EctLegM* legm, const int* nasm0, bool fp32, cudaStream_t st);

extern "C" int ect_vordiv_to_uv(int handle, int nsmax, const void* spvor, const void* spdiv, void* spu, void* spv,
                                int nfld, int memspace) {
    if (!spvor || !spdiv || !spu || !spv) return ECT_ERR_MISSING;
    if (nfld <= 0) return ECT_SUCCESS;
    EctHandle* h = nullptr;
    if (handle != 0) {
        h = get_handle(handle);
        if (!h || !h->d) { ect_set_error("ect_vordiv_to_uv: invalid handle"); return ECT_ERR_HANDLE; }
        if (nsmax != h->hp.nsmax) { ect_set_error("ect_vordiv_to_uv: nsmax %d differs from the handle's %d", nsmax, h->hp.nsmax); return ECT_ERR_BADARG; }
        ECT_CUDA(cudaSetDevice(h->d->dev));
    } else {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { ect_set_error("ect_vordiv_to_uv: no CUDA device (there is no CPU fallback)"); return ECT_ERR_CUDA; }
        if (nsmax < 0) return ECT_ERR_BADARG;
    }
    const bool fp32 = h && h->precision == ECT_PREC_SP;
    const int es = fp32 ? 4 : 8;
    const i64 nspec2 = h ? h->hp.nspec2 : (i64)(nsmax + 1) * (nsmax + 2);
    const int nump = h ? h->hp.nump : nsmax + 1;
    cudaStream_t st = h ? h->d->stream : (cudaStream_t)0;
    const size_t bytes = (size_t)nspec2 * nfld * es;
    const void *dv = spvor, *dd = spdiv; void *du = spu, *dvv = spv;
    char* tmp = nullptr;
    if (memspace == ECT_MEM_HOST) {
        ECT_CUDA(cudaMalloc(&tmp, 4 * bytes + 64));
        ECT_CUDA(cudaMemcpyAsync(tmp, spvor, bytes, cudaMemcpyHostToDevice, st));
        ECT_CUDA(cudaMemcpyAsync(tmp + bytes, spdiv, bytes, cudaMemcpyHostToDevice, st));
        dv = tmp; dd = tmp + bytes; du = tmp + 2 * bytes; dvv = tmp + 3 * bytes;
    }
    int rc = ECT_SUCCESS;
    if (nump > 0) rc = ect_launch_vd2uv(dv, dd, du, dvv, nfld, nsmax, nump, h ? h->d->legm : nullptr, h ? h->d->nasm0 : nullptr, fp32, st);
    if (h) h->d->launches++;
    if (memspace == ECT_MEM_HOST) {
        if (!rc) {
            ECT_CUDA(cudaMemcpyAsync(spu, du, bytes, cudaMemcpyDeviceToHost, st));
            ECT_CUDA(cudaMemcpyAsync(spv, dvv, bytes, cudaMemcpyDeviceToHost, st));
        }
        ECT_CUDA(cudaStreamSynchronize(st));
        ECT_CUDA(cudaFree(tmp));
    }
    return rc;
}

// ---------------------------------------------------------------------------------------
// Legendre polynomials in the reference's own layout.
//   ect_inquire_rpnm <- TRANS_INQ(PRPNM=...) cpu/external/trans_inq.F90:426-466: PRPNM(NDGNH, NSPOLEGL), column
//       NPMS(m) + p (p = 1 .. T+2-m) holds n = T+2-p, rows ISL .. NDGNH (ISL = NDGNH-NDGLU(m)+1), the others zero;
//       NPMS runs over this task's wavenumbers in MYMS order (sump_trans_preleg_mod.F90:124-131).
//   ect_trans_pnm    <- TRANS_PNM cpu/external/trans_pnm.F90:127-177, one wavenumber: PRPNM(ld, T-m+3)
// Both are read back from the table SETUP_TRANS left in HBM (the same values the transforms use).
// ---------------------------------------------------------------------------------------
static int rpnm_of_m(EctHandle* h, int ml, double* out, i64 ld, i64 col0) {
    const EctLegM& lm = h->d->h_legm[ml];
    const int T = h->hp.nsmax, m = lm.m, ndgnh = h->hp.ndgnh;
    std::vector<double> tab;
    for (int par = 0; par < 2; ++par) {
        const int k = par ? lm.ila : lm.ils;
        if (k == 0 || lm.ndglu == 0) continue;
        tab.resize((size_t)k * lm.ndglu);
        int rc = ect_legendre_get_table(h, ml, par, tab.data(), (long long)tab.size());
        if (rc) return rc;
        for (int kk = 0; kk < k; ++kk) {
            const int n = m + par + 2 * kk, p = T + 2 - n;          // 1-based column inside the block of m
            double* col = out + (col0 + p - 1) * ld;
            for (int i = 0; i < lm.ndglu; ++i) col[ndgnh - lm.ndglu + i] = tab[(size_t)kk * lm.ndglu + i];
        }
    }
    return ECT_SUCCESS;
}
extern "C" int ect_inquire_rpnm(int handle, double* rpnm, long long capacity_elems, int* nspolegl, int* npms /* nsmax+1, -1 = not local */) {
    EctHandle* h = get_handle(handle);
    if (!h || !h->d) { ect_set_error("ect_inquire_rpnm: invalid handle"); return ECT_ERR_HANDLE; }
    const EctHostPlan& P = h->hp;
    i64 ncol = 0;
    std::vector<i64> off(P.nump);
    for (int ml = 0; ml < P.nump; ++ml) { off[ml] = ncol; ncol += P.nsmax + 2 - P.myms[ml]; }
    if (nspolegl) *nspolegl = (int)ncol;
    if (npms) {
        for (int m = 0; m <= P.nsmax; ++m) npms[m] = -1;
        for (int ml = 0; ml < P.nump; ++ml) npms[P.myms[ml]] = (int)off[ml];
    }
    if (!rpnm) return ECT_SUCCESS;
    if ((i64)P.ndgnh * ncol > capacity_elems) { ect_set_error("ect_inquire_rpnm: array too small (%lld < %lld)", capacity_elems, (long long)P.ndgnh * ncol); return ECT_ERR_BADARG; }
    ECT_CUDA(cudaSetDevice(h->d->dev));
    memset(rpnm, 0, (size_t)P.ndgnh * ncol * sizeof(double));
    for (int ml = 0; ml < P.nump; ++ml) { int rc = rpnm_of_m(h, ml, rpnm, P.ndgnh, off[ml]); if (rc) return rc; }
    return ECT_SUCCESS;
}
extern "C" int ect_trans_pnm(int handle, int m, double* rpnm, int ld, int ncols) {
    EctHandle* h = get_handle(handle);
    if (!h || !h->d) { ect_set_error("ect_trans_pnm: invalid handle"); return ECT_ERR_HANDLE; }
    const EctHostPlan& P = h->hp;
    if (!rpnm) return ECT_ERR_MISSING;
    if (m < 0 || m > P.nsmax || ld < P.ndgnh || ncols < P.nsmax - m + 2) { ect_set_error("ect_trans_pnm: m, ld or ncols out of range"); return ECT_ERR_BADARG; }
    int ml = -1;
    for (int i = 0; i < P.nump; ++i) if (P.myms[i] == m) ml = i;
    if (ml < 0) { ect_set_error("ect_trans_pnm: wavenumber %d is not held by this task", m); return ECT_ERR_NOTIMPL; }
    ECT_CUDA(cudaSetDevice(h->d->dev));
    for (i64 c = 0; c < ncols; ++c) memset(rpnm + c * ld, 0, (size_t)ld * sizeof(double));
    return rpnm_of_m(h, ml, rpnm, ld, 0);
}

// ---------------------------------------------------------------------------------------
// Legendre polynomial cache files in the reference's format (SETUP_TRANS CDIO_LEGPOL = 'writef' / 'readf',
// cpu/internal/write_legpol_mod.F90:57-170, read_legpol_mod.F90:60-150; plain matrices only -- no FLT butterfly
// structs, no lat-lon section):
//   int32[4]           'LEGP' 'OL  ' NSMAX NDGNH
//   int32[2 NDGNH]     NLOEN(j), NMEN(j) of the northern latitudes, interleaved
//   per local m (MYMS order): real64 RPNMA(IDGLU, ILA) then RPNMS(IDGLU, ILS), column major, column J holding
//                      n = m + 2 (ILA - J) + 1 resp. n = m + 2 (ILS - J)   (n descending)
//   int32[4]           'LEGPOL---EOF-EOF'
// ---------------------------------------------------------------------------------------
extern "C" int ect_write_legpol(int handle, const char* path) {
    EctHandle* h = get_handle(handle);
    if (!h || !h->d) { ect_set_error("ect_write_legpol: invalid handle"); return ECT_ERR_HANDLE; }
    if (!path) return ECT_ERR_MISSING;
    const EctHostPlan& P = h->hp;
    ECT_CUDA(cudaSetDevice(h->d->dev));
    FILE* f = fopen(path, "wb");
    if (!f) { ect_set_error("ect_write_legpol: cannot open %s", path); return ECT_ERR_GENERIC; }
    int32_t hdr[4]; memcpy(hdr, "LEGPOL  ", 8); hdr[2] = P.nsmax; hdr[3] = P.ndgnh;
    bool ok = fwrite(hdr, 4, 4, f) == 4;
    std::vector<int32_t> geo(2 * (size_t)P.ndgnh);
    for (int j = 0; j < P.ndgnh; ++j) { geo[2 * j] = P.nloen[j]; geo[2 * j + 1] = P.nmen[j]; }
    ok = ok && fwrite(geo.data(), 4, geo.size(), f) == geo.size();
    std::vector<double> tab, out;
    for (int ml = 0; ml < P.nump && ok; ++ml) {
        const EctLegM& lm = h->d->h_legm[ml];
        for (int par = 1; par >= 0 && ok; --par) {          // antisymmetric first
            const int k = par ? lm.ila : lm.ils;
            if (k == 0 || lm.ndglu == 0) continue;
            tab.resize((size_t)k * lm.ndglu); out.resize(tab.size());
            int rc = ect_legendre_get_table(h, ml, par, tab.data(), (long long)tab.size());
            if (rc) { fclose(f); return rc; }
            for (int kk = 0; kk < k; ++kk)                   // mine: n ascending; file: column J = k - kk (n descending)
                memcpy(out.data() + (size_t)(k - 1 - kk) * lm.ndglu, tab.data() + (size_t)kk * lm.ndglu, lm.ndglu * sizeof(double));
            ok = fwrite(out.data(), 8, out.size(), f) == out.size();
        }
    }
    ok = ok && fwrite("LEGPOL---EOF-EOF", 1, 16, f) == 16;
    ok = (fclose(f) == 0) && ok;
    if (!ok) { ect_set_error("ect_write_legpol: write to %s failed", path); return ECT_ERR_GENERIC; }
    return ECT_SUCCESS;
}

extern "C" int ect_read_legpol(int handle, const char* path) {
    EctHandle* h = get_handle(handle);
    if (!h || !h->d) { ect_set_error("ect_read_legpol: invalid handle"); return ECT_ERR_HANDLE; }
    if (!path) return ECT_ERR_MISSING;
    const EctHostPlan& P = h->hp;
    ECT_CUDA(cudaSetDevice(h->d->dev));
    FILE* f = fopen(path, "rb");
    if (!f) { ect_set_error("ect_read_legpol: cannot open %s", path); return ECT_ERR_GENERIC; }
    auto fail = [&](const char* msg) { ect_set_error("READ_LEGPOL: %s (%s)", msg, path); fclose(f); return ECT_ERR_BADARG; };
    int32_t hdr[4];
    if (fread(hdr, 4, 4, f) != 4) return fail("SHORT FILE");
    if (memcmp(hdr, "LEGPOL  ", 8) != 0) return fail("WRONG LABEL");
    if (hdr[2] != P.nsmax) return fail("WRONG SPECTRAL TRUNCATION");
    if (hdr[3] != P.ndgnh) return fail("WRONG NO OF GAUSSIAN LATITUDES");
    std::vector<int32_t> geo(2 * (size_t)P.ndgnh);
    if (fread(geo.data(), 4, geo.size(), f) != geo.size()) return fail("SHORT FILE");
    for (int j = 0; j < P.ndgnh; ++j) {
        if (geo[2 * j] != P.nloen[j]) return fail("WRONG NLOEN");
        if (geo[2 * j + 1] != P.nmen[j]) return fail("WRONG NMEN");
    }
    std::vector<double> in, tab;
    for (int ml = 0; ml < P.nump; ++ml) {
        const EctLegM& lm = h->d->h_legm[ml];
        for (int par = 1; par >= 0; --par) {
            const int k = par ? lm.ila : lm.ils;
            if (k == 0 || lm.ndglu == 0) continue;
            in.resize((size_t)k * lm.ndglu); tab.resize(in.size());
            if (fread(in.data(), 8, in.size(), f) != in.size()) return fail("SHORT FILE");
            for (int kk = 0; kk < k; ++kk)
                memcpy(tab.data() + (size_t)kk * lm.ndglu, in.data() + (size_t)(k - 1 - kk) * lm.ndglu, lm.ndglu * sizeof(double));
            int rc = ect_legendre_set_table(h, ml, par, tab.data());
            if (rc) { fclose(f); return rc; }
        }
    }
    char eof[16];
    if (fread(eof, 1, 16, f) != 16 || memcmp(eof, "LEGPOL---EOF-EOF", 16) != 0) return fail("WRONG END LABEL");
    fclose(f);
    h->defer_table = false;
    return ECT_SUCCESS;
}

// ---------------------------------------------------------------------------------------
// V-sets (NPRTRV > 1): INV_TRANS / DIR_TRANS with the fields (levels) of the spectral arrays spread over the V tasks of
// a W-group (KVSETUV / KVSETSC / KVSETSC2 / KVSETSC3A / KVSETSC3B, inv_trans.h:36-58) and grid-point arrays that carry
// every field on the task's eq_regions points.  The W-group transforms its local fields exactly as a W-task run does,
// into / out of the band buffer; TRLTOG / TRGTOL (trltog_mod.F90:213-271: "field -> V-set redistribution") move points
// and fields at once.  Field order of the messages: V-set 0's fields, then V-set 1's, ... each in the Fourier order of
// ftinv_ctl_mod.F90:144-166 restricted to its levels -- which is the order the W-group transform of that V-set uses.
// ---------------------------------------------------------------------------------------
namespace {
struct DevFree { void* p = nullptr; ~DevFree() { if (p) cudaFree(p); } };      // frees a temporary device buffer on every return path
struct VsLists {
    std::vector<int> uv, sc;          // V-set (0-based) of every global vor/div/u/v level and of every global scalar field
    int nuv_g = 0, nsc2_g = 0, lev3a_g = 0, lev3b_g = 0, nsc_g = 0;
    bool mode2_sp = false;
};
}
static int vs_lists(const EctHandle* h, const ect_vset_args* vs, bool mode2, int n3a_fld, int n3b_fld, VsLists& L, const char* who) {
    const int V = h->vs.V;
    auto take = [&](const int* k, int n, std::vector<int>& out, const char* name) -> int {
        if (n < 0 || (n > 0 && !k)) { ect_set_error("%s: %s missing", who, name); return ECT_ERR_MISSING; }
        for (int i = 0; i < n; ++i) {
            if (k[i] < 1 || k[i] > V) { ect_set_error("%s: %s TOO LONG OR CONTAINS VALUES OUTSIDE RANGE", who, name); return ECT_ERR_BADARG; }
            out.push_back(k[i] - 1);
        }
        return ECT_SUCCESS;
    };
    int rc;
    L.mode2_sp = mode2;
    L.nuv_g = vs->nuv_g;
    if ((rc = take(vs->kvsetuv, vs->nuv_g, L.uv, "KVSETUV"))) return rc;
    if (!mode2) { if ((rc = take(vs->kvsetsc, vs->nscalar_g, L.sc, "KVSETSC"))) return rc; }
    else {
        L.nsc2_g = vs->nsc2_g; L.lev3a_g = n3a_fld ? vs->nsc3a_lev_g : 0; L.lev3b_g = n3b_fld ? vs->nsc3b_lev_g : 0;
        if ((rc = take(vs->kvsetsc2, vs->nsc2_g, L.sc, "KVSETSC2"))) return rc;
        for (int j = 0; j < n3a_fld; ++j) if ((rc = take(vs->kvsetsc3a, L.lev3a_g, L.sc, "KVSETSC3A"))) return rc;
        for (int j = 0; j < n3b_fld; ++j) if ((rc = take(vs->kvsetsc3b, L.lev3b_g, L.sc, "KVSETSC3B"))) return rc;
    }
    L.nsc_g = (int)L.sc.size();
    return ECT_SUCCESS;
}
static int vs_count(const std::vector<int>& v, int n, int me) { int c = 0; for (int i = 0; i < n; ++i) c += (v[i] == me); return c; }

// message order of the caller's fields + device table (base, block stride) in that order
static int vs_upload_table(EctDevice* d, const std::vector<int>& gv, int V, const std::vector<double*>& base, const std::vector<i64>& stride,
                           std::vector<int>& nfl, char** tab) {
    const int n = (int)gv.size();
    nfl.assign(V, 0);
    std::vector<double*> pb; std::vector<i64> ps;
    for (int v = 0; v < V; ++v)
        for (int g = 0; g < n; ++g) if (gv[g] == v) { pb.push_back(base[g]); ps.push_back(stride[g]); ++nfl[v]; }
    ECT_CUDA(cudaMalloc(tab, std::max(n, 1) * (sizeof(double*) + sizeof(i64))));
    if (n) {
        ECT_CUDA(cudaMemcpyAsync(*tab, pb.data(), n * sizeof(double*), cudaMemcpyHostToDevice, d->stream));
        ECT_CUDA(cudaMemcpyAsync(*tab + n * sizeof(double*), ps.data(), n * sizeof(i64), cudaMemcpyHostToDevice, d->stream));
        ECT_CUDA(cudaStreamSynchronize(d->stream));     // pb / ps are about to go out of scope
    }
    return ECT_SUCCESS;
}

extern "C" int ect_inv_trans_vset(int handle, const ect_inv_args* a, const ect_vset_args* vs) {
    EctHandle* h = get_handle(handle);
    if (!h) { ect_set_error("ect_inv_trans_vset: invalid handle %d", handle); return ECT_ERR_HANDLE; }
    if (!a || !vs) return ECT_ERR_MISSING;
    if (h->vs.V == 1) return ect_inv_trans(handle, a);
    if (!h->d) { ect_set_error("ect_inv_trans_vset: handle was set up host-only"); return ECT_ERR_CUDA; }
    const EctHostPlan& P = h->hp;
    EctDevice* d = h->d;
    ECT_CUDA(cudaSetDevice(d->dev));
    const int V = h->vs.V, me = h->vs.v, ngp = h->vs.ngptot;
    const int es = h->precision == ECT_PREC_SP ? 4 : 8;
    const bool mode2_sp = (a->spscalar == nullptr) && (a->spsc2 || a->spsc3a || a->spsc3b || vs->nsc2_g || vs->nsc3a_lev_g || vs->nsc3b_lev_g);
    const int f3a = mode2_sp ? a->nsc3a_fld : 0, f3b = mode2_sp ? a->nsc3b_fld : 0;
    VsLists L;
    int rc;
    if ((rc = vs_lists(h, vs, mode2_sp, f3a, f3b, L, "INV_TRANS"))) return rc;
    // local counts must be what the V-set arrays say (inv_trans.F90:236-330)
    const int nuv_l = vs_count(L.uv, L.nuv_g, me);
    if (nuv_l != a->nuv) { ect_set_error("INV_TRANS: PSPVOR holds %d fields, KVSETUV gives this V-set %d", a->nuv, nuv_l); return ECT_ERR_BADARG; }
    if (!mode2_sp && vs_count(L.sc, L.nsc_g, me) != (a->spscalar ? a->nscalar : 0)) { ect_set_error("INV_TRANS: PSPSCALAR field count differs from KVSETSC"); return ECT_ERR_BADARG; }
    if (mode2_sp) {
        std::vector<int> k2(L.sc.begin(), L.sc.begin() + L.nsc2_g);
        if (vs_count(k2, L.nsc2_g, me) != (a->spsc2 ? a->nsc2 : 0)) { ect_set_error("INV_TRANS: PSPSC2 field count differs from KVSETSC2"); return ECT_ERR_BADARG; }
        if (f3a) { std::vector<int> k(L.sc.begin() + L.nsc2_g, L.sc.begin() + L.nsc2_g + L.lev3a_g); if (vs_count(k, L.lev3a_g, me) != a->nsc3a_lev) { ect_set_error("INV_TRANS: PSPSC3A level count differs from KVSETSC3A"); return ECT_ERR_BADARG; } }
        if (f3b) { std::vector<int> k(L.sc.end() - (size_t)f3b * L.lev3b_g, L.sc.end() - (size_t)(f3b - 1) * L.lev3b_g); if (vs_count(k, L.lev3b_g, me) != a->nsc3b_lev) { ect_set_error("INV_TRANS: PSPSC3B level count differs from KVSETSC3B"); return ECT_ERR_BADARG; } }
    }
    const bool scders = a->scders && L.nsc_g > 0, vorgp = a->vorgp && L.nuv_g > 0, divgp = a->divgp && L.nuv_g > 0, uvder = a->uvder && L.nuv_g > 0;
    // global Fourier field list (V-set of every field), ftinv_ctl_mod.F90:144-166
    std::vector<int> gv;
    auto app = [&](const std::vector<int>& x) { gv.insert(gv.end(), x.begin(), x.end()); };
    if (vorgp) app(L.uv);
    if (divgp) app(L.uv);
    if (L.nuv_g) { app(L.uv); app(L.uv); }
    app(L.sc);
    if (scders) app(L.sc);
    if (uvder) { app(L.uv); app(L.uv); }
    if (scders) app(L.sc);
    const int nfg = (int)gv.size();
    if (nfg == 0) return ECT_SUCCESS;
    // the caller's grid-point arrays (all fields): base pointer / block stride per global Fourier field (trltog_mod.F90:579-731)
    const int nproma = (a->nproma > 0 && a->nproma < ngp) ? a->nproma : std::max(ngp, 1);
    const i64 blk = (i64)nproma * ((ngp + nproma - 1) / nproma);
    const bool mode2_gp = (a->gp == nullptr);
    const int nvar_uv = (vorgp ? 1 : 0) + (divgp ? 1 : 0) + 2 + (uvder ? 2 : 0), dfac = scders ? 3 : 1;
    const int n3a_g = f3a * L.lev3a_g, n3b_g = f3b * L.lev3b_g;
    const i64 sz_gp = (i64)nfg * blk, sz_uv = (i64)L.nuv_g * nvar_uv * blk, sz_2 = (i64)L.nsc2_g * dfac * blk, sz_3a = (i64)n3a_g * dfac * blk, sz_3b = (i64)n3b_g * dfac * blk;
    if (mode2_gp && ((L.nuv_g && !a->gpuv) || (L.nsc2_g && !a->gp2) || (n3a_g && !a->gp3a) || (n3b_g && !a->gp3b) || (!mode2_sp && L.nsc_g))) { ect_set_error("INV_TRANS: grid-point output array missing"); return ECT_ERR_MISSING; }
    const bool host = a->memspace == ECT_MEM_HOST;
    double *q_gp = a->gp, *q_uv = a->gpuv, *q_2 = a->gp2, *q_3a = a->gp3a, *q_3b = a->gp3b;
    if (host) {
        const i64 tot = mode2_gp ? sz_uv + sz_2 + sz_3a + sz_3b : sz_gp;
        if ((rc = ensure(d->stage_gp, d->stage_gp_elems, tot, d->stream, false))) return rc;
        double* p = d->stage_gp;
        if (!mode2_gp) q_gp = p;
        else { q_uv = p; p = adv(p, sz_uv, es); q_2 = p; p = adv(p, sz_2, es); q_3a = p; p = adv(p, sz_3a, es); q_3b = p; }
    }
    std::vector<double*> base(nfg); std::vector<i64> stride(nfg);
    if (!mode2_gp) for (int i = 0; i < nfg; ++i) { base[i] = adv(q_gp, (i64)i * nproma, es); stride[i] = (i64)nfg * nproma; }
    else {
        int fi = 0, var = 0;
        auto uvgroup = [&](int vv) { for (int l = 0; l < L.nuv_g; ++l, ++fi) { base[fi] = adv(q_uv, ((i64)vv * L.nuv_g + l) * nproma, es); stride[fi] = (i64)nproma * L.nuv_g * nvar_uv; } };
        auto scgroup = [&](int part) {
            for (int j = 0; j < L.nsc2_g; ++j, ++fi) { base[fi] = adv(q_2, ((i64)part * L.nsc2_g + j) * nproma, es); stride[fi] = (i64)nproma * L.nsc2_g * dfac; }
            for (int j3 = 0; j3 < f3a; ++j3) for (int l = 0; l < L.lev3a_g; ++l, ++fi) { base[fi] = adv(q_3a, (((i64)part * f3a + j3) * L.lev3a_g + l) * nproma, es); stride[fi] = (i64)nproma * L.lev3a_g * f3a * dfac; }
            for (int j3 = 0; j3 < f3b; ++j3) for (int l = 0; l < L.lev3b_g; ++l, ++fi) { base[fi] = adv(q_3b, (((i64)part * f3b + j3) * L.lev3b_g + l) * nproma, es); stride[fi] = (i64)nproma * L.lev3b_g * f3b * dfac; }
        };
        if (vorgp) uvgroup(var++);
        if (divgp) uvgroup(var++);
        if (L.nuv_g) { uvgroup(var++); uvgroup(var++); }
        scgroup(0);
        if (scders) scgroup(1);
        if (uvder) { uvgroup(var++); uvgroup(var++); }
        if (scders) scgroup(2);
    }
    std::vector<int> nfl; char* tab = nullptr;
    DevFree f_tab, f_sp;
    if ((rc = vs_upload_table(d, gv, V, base, stride, nfl, &tab))) return rc;
    f_tab.p = tab;
    if ((rc = gp_exchange_setup(h, nfl[me], nfg))) return rc;
    // ---- the W-group's transform of this V-set's fields into the band buffer (PGP(band points, local fields)) ----
    ect_inv_args la = *a;
    la.memspace = ECT_MEM_DEVICE; la.nproma = 0;
    la.gp = d->gpband; la.gpuv = la.gp2 = la.gp3a = la.gp3b = nullptr;
    char* sp_tmp = nullptr;
    if (host) {       // spectral inputs to the device
        const i64 nsp = P.nspec2;
        const i64 nsc_l = mode2_sp ? (i64)(a->spsc2 ? a->nsc2 : 0) + (a->spsc3a ? (i64)a->nsc3a_lev * a->nsc3a_fld : 0) + (a->spsc3b ? (i64)a->nsc3b_lev * a->nsc3b_fld : 0)
                                   : (a->spscalar ? a->nscalar : 0);
        ECT_CUDA(cudaMalloc(&sp_tmp, (size_t)std::max<i64>((2 * (i64)a->nuv + nsc_l) * nsp * es, 16)));
        f_sp.p = sp_tmp;
        char* p = sp_tmp;
        auto up = [&](const double*& ptr, i64 n) -> int {
            if (!ptr || n == 0) return ECT_SUCCESS;
            ECT_CUDA(cudaMemcpyAsync(p, ptr, (size_t)n * es, cudaMemcpyHostToDevice, d->stream));
            ptr = (const double*)p; p += (size_t)n * es;
            return ECT_SUCCESS;
        };
        if ((rc = up(la.spvor, a->nuv * nsp)) || (rc = up(la.spdiv, a->nuv * nsp)) || (rc = up(la.spscalar, (i64)a->nscalar * nsp)) ||
            (rc = up(la.spsc2, (i64)a->nsc2 * nsp)) || (rc = up(la.spsc3a, (i64)a->nsc3a_lev * a->nsc3a_fld * nsp)) ||
            (rc = up(la.spsc3b, (i64)a->nsc3b_lev * a->nsc3b_fld * nsp))) return rc;
    }
    rc = inv_trans_impl(handle, &la, nullptr, 0);
    // ---- TRLTOG: points and fields to the grid-point tasks ----
    if (!rc) rc = gp_exchange(h, nfl[me], nfg, nfl, es, 1, (double* const*)tab, (const i64*)(tab + nfg * sizeof(double*)), nproma);
    if (!rc && host) {
        if (!mode2_gp) ECT_CUDA(cudaMemcpyAsync(a->gp, q_gp, (size_t)sz_gp * es, cudaMemcpyDeviceToHost, d->stream));
        else {
            if (sz_uv) ECT_CUDA(cudaMemcpyAsync(a->gpuv, q_uv, (size_t)sz_uv * es, cudaMemcpyDeviceToHost, d->stream));
            if (sz_2) ECT_CUDA(cudaMemcpyAsync(a->gp2, q_2, (size_t)sz_2 * es, cudaMemcpyDeviceToHost, d->stream));
            if (sz_3a) ECT_CUDA(cudaMemcpyAsync(a->gp3a, q_3a, (size_t)sz_3a * es, cudaMemcpyDeviceToHost, d->stream));
            if (sz_3b) ECT_CUDA(cudaMemcpyAsync(a->gp3b, q_3b, (size_t)sz_3b * es, cudaMemcpyDeviceToHost, d->stream));
        }
    }
    ECT_CUDA(cudaStreamSynchronize(d->stream));
    return rc;
}

extern "C" int ect_dir_trans_vset(int handle, const ect_dir_args* a, const ect_vset_args* vs) {
    EctHandle* h = get_handle(handle);
    if (!h) { ect_set_error("ect_dir_trans_vset: invalid handle %d", handle); return ECT_ERR_HANDLE; }
    if (!a || !vs) return ECT_ERR_MISSING;
    if (h->vs.V == 1) return ect_dir_trans(handle, a);
    if (!h->d) { ect_set_error("ect_dir_trans_vset: handle was set up host-only"); return ECT_ERR_CUDA; }
    const EctHostPlan& P = h->hp;
    EctDevice* d = h->d;
    ECT_CUDA(cudaSetDevice(d->dev));
    const int V = h->vs.V, me = h->vs.v, ngp = h->vs.ngptot;
    const int es = h->precision == ECT_PREC_SP ? 4 : 8;
    const bool mode2 = (a->gp == nullptr);
    const int f3a = mode2 && a->gp3a ? a->nsc3a_fld : 0, f3b = mode2 && a->gp3b ? a->nsc3b_fld : 0;
    VsLists L;
    int rc;
    if ((rc = vs_lists(h, vs, mode2, f3a, f3b, L, "DIR_TRANS"))) return rc;
    const int nuv_l = vs_count(L.uv, L.nuv_g, me), nsc_l = vs_count(L.sc, L.nsc_g, me);
    if (nuv_l != a->nuv) { ect_set_error("DIR_TRANS: PSPVOR holds %d fields, KVSETUV gives this V-set %d", a->nuv, nuv_l); return ECT_ERR_BADARG; }
    int nsc2_l = 0, lev3a_l = 0, lev3b_l = 0;
    if (mode2) {
        nsc2_l = vs_count(L.sc, L.nsc2_g, me);
        if (f3a) { std::vector<int> k(L.sc.begin() + L.nsc2_g, L.sc.begin() + L.nsc2_g + L.lev3a_g); lev3a_l = vs_count(k, L.lev3a_g, me); }
        if (f3b) { std::vector<int> k(L.sc.end() - (size_t)f3b * L.lev3b_g, L.sc.end() - (size_t)(f3b - 1) * L.lev3b_g); lev3b_l = vs_count(k, L.lev3b_g, me); }
        if (nsc2_l != (a->gp2 ? a->nsc2 : 0) || (f3a && lev3a_l != a->nsc3a_lev) || (f3b && lev3b_l != a->nsc3b_lev)) { ect_set_error("DIR_TRANS: local field counts differ from KVSETSC2 / KVSETSC3A / KVSETSC3B"); return ECT_ERR_BADARG; }
        if ((nsc2_l && !a->spsc2) || (f3a && lev3a_l && !a->spsc3a) || (f3b && lev3b_l && !a->spsc3b)) { ect_set_error("DIR_TRANS: missing spectral output for call mode 2"); return ECT_ERR_MISSING; }
    } else if (nsc_l != a->nscalar) { ect_set_error("DIR_TRANS: PSPSCALAR field count differs from KVSETSC"); return ECT_ERR_BADARG; }
    std::vector<int> gv;
    if (L.nuv_g) { gv.insert(gv.end(), L.uv.begin(), L.uv.end()); gv.insert(gv.end(), L.uv.begin(), L.uv.end()); }
    gv.insert(gv.end(), L.sc.begin(), L.sc.end());
    const int nfg = (int)gv.size();
    if (nfg == 0) return ECT_SUCCESS;
    const int nproma = (a->nproma > 0 && a->nproma < ngp) ? a->nproma : std::max(ngp, 1);
    const i64 blk = (i64)nproma * ((ngp + nproma - 1) / nproma);
    const int n3a_g = f3a * L.lev3a_g, n3b_g = f3b * L.lev3b_g;
    const i64 sz_gp = (i64)nfg * blk, sz_uv = (i64)L.nuv_g * 2 * blk, sz_2 = (i64)L.nsc2_g * blk, sz_3a = (i64)n3a_g * blk, sz_3b = (i64)n3b_g * blk;
    if (mode2 && ((L.nuv_g && !a->gpuv) || (L.nsc2_g && !a->gp2))) { ect_set_error("DIR_TRANS: grid-point input array missing"); return ECT_ERR_MISSING; }
    const bool host = a->memspace == ECT_MEM_HOST;
    const double *q_gp = a->gp, *q_uv = a->gpuv, *q_2 = a->gp2, *q_3a = a->gp3a, *q_3b = a->gp3b;
    if (host) {
        const i64 tot = mode2 ? sz_uv + sz_2 + sz_3a + sz_3b : sz_gp;
        if ((rc = ensure(d->stage_gp, d->stage_gp_elems, tot, d->stream, false))) return rc;
        double* p = d->stage_gp;
        auto up = [&](const double*& ptr, i64 n) -> int {
            if (!ptr || n == 0) return ECT_SUCCESS;
            ECT_CUDA(cudaMemcpyAsync(p, ptr, (size_t)n * es, cudaMemcpyHostToDevice, d->stream));
            ptr = p; p = adv(p, n, es);
            return ECT_SUCCESS;
        };
        if (!mode2) { if ((rc = up(q_gp, sz_gp))) return rc; }
        else if ((rc = up(q_uv, sz_uv)) || (rc = up(q_2, sz_2)) || (rc = up(q_3a, sz_3a)) || (rc = up(q_3b, sz_3b))) return rc;
    }
    std::vector<double*> base(nfg); std::vector<i64> stride(nfg);
    if (!mode2) for (int i = 0; i < nfg; ++i) { base[i] = (double*)adv(q_gp, (i64)i * nproma, es); stride[i] = (i64)nfg * nproma; }
    else {
        int fi = 0;
        for (int vv = 0; vv < (L.nuv_g ? 2 : 0); ++vv) for (int l = 0; l < L.nuv_g; ++l, ++fi) { base[fi] = (double*)adv(q_uv, ((i64)vv * L.nuv_g + l) * nproma, es); stride[fi] = (i64)nproma * L.nuv_g * 2; }
        for (int j = 0; j < L.nsc2_g; ++j, ++fi) { base[fi] = (double*)adv(q_2, (i64)j * nproma, es); stride[fi] = (i64)nproma * L.nsc2_g; }
        for (int j3 = 0; j3 < f3a; ++j3) for (int l = 0; l < L.lev3a_g; ++l, ++fi) { base[fi] = (double*)adv(q_3a, ((i64)j3 * L.lev3a_g + l) * nproma, es); stride[fi] = (i64)nproma * L.lev3a_g * f3a; }
        for (int j3 = 0; j3 < f3b; ++j3) for (int l = 0; l < L.lev3b_g; ++l, ++fi) { base[fi] = (double*)adv(q_3b, ((i64)j3 * L.lev3b_g + l) * nproma, es); stride[fi] = (i64)nproma * L.lev3b_g * f3b; }
    }
    std::vector<int> nfl; char* tab = nullptr;
    DevFree f_tab, f_sp;
    if ((rc = vs_upload_table(d, gv, V, base, stride, nfl, &tab))) return rc;
    f_tab.p = tab;
    if ((rc = gp_exchange_setup(h, nfl[me], nfg))) return rc;
    // ---- TRGTOL: this V-set's fields of the band's points arrive in the band buffer ----
    rc = gp_exchange(h, nfl[me], nfg, nfl, es, 0, (double* const*)tab, (const i64*)(tab + nfg * sizeof(double*)), nproma);
    ECT_CUDA(cudaStreamSynchronize(d->stream));
    if (rc) return rc;
    // ---- the W-group's direct transform: call mode 1 on the band buffer, one (nsc_local, nspec2) scalar array ----
    const i64 nsp = P.nspec2;
    char* sp_tmp = nullptr;
    ECT_CUDA(cudaMalloc(&sp_tmp, (size_t)std::max<i64>((2 * (i64)nuv_l + nsc_l) * nsp * es, 16)));
    f_sp.p = sp_tmp;
    ect_dir_args la = *a;
    la.memspace = ECT_MEM_DEVICE; la.nproma = 0; la.nuv = nuv_l; la.nscalar = nsc_l;
    la.gp = d->gpband; la.gpuv = la.gp2 = la.gp3a = la.gp3b = nullptr;
    la.spsc2 = la.spsc3a = la.spsc3b = nullptr;
    double* t_vor = (double*)sp_tmp; double* t_div = adv(t_vor, nuv_l * nsp, es); double* t_sc = adv(t_div, nuv_l * nsp, es);
    const bool direct_out = !host && !mode2;        // device pointers, call mode 1: the caller's arrays take the results as they come
    if (!direct_out) { la.spvor = t_vor; la.spdiv = t_div; la.spscalar = t_sc; }
    rc = dir_trans_impl(handle, &la, nullptr, 0);
    if (!rc && !direct_out) {
        const cudaMemcpyKind kind = host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
        if (nuv_l) {
            ECT_CUDA(cudaMemcpyAsync(a->spvor, t_vor, (size_t)nuv_l * nsp * es, kind, d->stream));
            ECT_CUDA(cudaMemcpyAsync(a->spdiv, t_div, (size_t)nuv_l * nsp * es, kind, d->stream));
        }
        if (!mode2) { if (nsc_l) ECT_CUDA(cudaMemcpyAsync(a->spscalar, t_sc, (size_t)nsc_l * nsp * es, kind, d->stream)); }
        else {
            // scalars came out as one (nsc_local, nspec2) array in the order [sc2][3a fields x levels][3b ...]
            const size_t pitch = (size_t)nsc_l * es;
            int col = 0;
            if (nsc2_l) { ECT_CUDA(cudaMemcpy2DAsync(a->spsc2, (size_t)nsc2_l * es, (char*)t_sc, pitch, (size_t)nsc2_l * es, nsp, kind, d->stream)); col += nsc2_l; }
            for (int j3 = 0; j3 < f3a; ++j3, col += lev3a_l)
                if (lev3a_l) ECT_CUDA(cudaMemcpy2DAsync((char*)a->spsc3a + (size_t)j3 * lev3a_l * nsp * es, (size_t)lev3a_l * es, (char*)t_sc + (size_t)col * es, pitch, (size_t)lev3a_l * es, nsp, kind, d->stream));
            for (int j3 = 0; j3 < f3b; ++j3, col += lev3b_l)
                if (lev3b_l) ECT_CUDA(cudaMemcpy2DAsync((char*)a->spsc3b + (size_t)j3 * lev3b_l * nsp * es, (size_t)lev3b_l * es, (char*)t_sc + (size_t)col * es, pitch, (size_t)lev3b_l * es, nsp, kind, d->stream));
        }
    }
    ECT_CUDA(cudaStreamSynchronize(d->stream));
    return rc;
}
