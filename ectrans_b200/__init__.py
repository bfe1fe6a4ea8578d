"""ectrans_b200 -- B200-native global spectral transform (SETUP_TRANS / INV_TRANS / DIR_TRANS).

Host-side Python face over the C ABI in ``include/ectrans_b200.h`` (ctypes).  The
compute path is the CUDA library ``ectrans_b200/lib/libectrans_b200.so`` built for
sm_100a; there is no CPU fallback -- if the library is missing or no GPU is present
the calls fail loudly.

Array conventions are the reference's Fortran layouts seen from C:
  spectral  : ``(nspec2, nfld)``  C-contiguous  == Fortran ``PSPEC(nfld, nspec2)``
  gridpoint : ``(ngpblks, nfld, nproma)``       == Fortran ``PGP(nproma, nfld, ngpblks)``
(reference: src/trans/include/ectrans/inv_trans.h:38-76, dir_trans.h:36-61;
transi ``rspscalar[nspec2][nscalar]``, ``rgp[ngpblks][nfld][nproma]`` src/transi/transi.h:986-1010).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.environ.get("ECTRANS_B200_LIB") or os.path.join(_HERE, "lib", "libectrans_b200.so")    # override: A/B builds
_lib = None

ECT_MEM_HOST, ECT_MEM_DEVICE = 0, 1
ECT_SETUP_HOST_ONLY = 1
ECT_SETUP_STREAM_GIVEN = 2
ECT_SETUP_LEGPOL_DEFER = 4
ECT_SETUP_GP_EQ_REGIONS = 8
ECT_SETUP_BANDS_BY_POINTS = 16
ECT_NCCL_UID_BYTES = 128
ECT_PREC_DP, ECT_PREC_SP = 0, 1
(ARR_NLOEN, ARR_NMEN, ARR_NDGLU, ARR_MYMS, ARR_NASM0, ARR_NPROCM, ARR_RMU, ARR_RGW, ARR_LATFIRST,
 ARR_LATCOUNT, ARR_SENDCNT, ARR_RECVCNT, ARR_RACTHE, ARR_MROW0, ARR_LEGRECN, ARR_LEGRECS, ARR_LATROW0,
 ARR_FFTREC, ARR_SENDOFF, ARR_RECVOFF, ARR_LEGDSTRANKN, ARR_LEGDSTRECN, ARR_LEGDSTRANKS, ARR_LEGDSTRECS,
 ARR_FFTDSTRANK, ARR_FFTDSTREC) = range(1, 27)


class EctError(RuntimeError):
    pass


class _SetupOpts(C.Structure):
    _fields_ = [("nsmax", C.c_int), ("ndgl", C.c_int), ("nloen", C.POINTER(C.c_int)), ("nranks", C.c_int),
                ("rank", C.c_int), ("flags", C.c_int), ("device", C.c_int), ("stream", C.c_void_p),
                ("nccl_uid", C.c_void_p), ("precision", C.c_int)]


class Info(C.Structure):
    _fields_ = [("nsmax", C.c_int), ("ndgl", C.c_int), ("ndgnh", C.c_int), ("nranks", C.c_int), ("rank", C.c_int),
                ("nspec2", C.c_int), ("nspec2g", C.c_int), ("ngptot", C.c_int), ("ngptotg", C.c_int),
                ("nump", C.c_int), ("lat0", C.c_int), ("nlat", C.c_int), ("table_bytes", C.c_longlong)]


class _InvArgs(C.Structure):
    _fields_ = [("memspace", C.c_int), ("nproma", C.c_int), ("scders", C.c_int), ("vorgp", C.c_int),
                ("divgp", C.c_int), ("uvder", C.c_int),
                ("spvor", C.c_void_p), ("spdiv", C.c_void_p), ("nuv", C.c_int),
                ("spscalar", C.c_void_p), ("nscalar", C.c_int),
                ("spsc2", C.c_void_p), ("nsc2", C.c_int),
                ("spsc3a", C.c_void_p), ("nsc3a_lev", C.c_int), ("nsc3a_fld", C.c_int),
                ("spsc3b", C.c_void_p), ("nsc3b_lev", C.c_int), ("nsc3b_fld", C.c_int),
                ("gp", C.c_void_p), ("gpuv", C.c_void_p), ("gp2", C.c_void_p), ("gp3a", C.c_void_p),
                ("gp3b", C.c_void_p)]


class _DirArgs(C.Structure):
    _fields_ = [("memspace", C.c_int), ("nproma", C.c_int), ("nuv", C.c_int), ("nscalar", C.c_int),
                ("gp", C.c_void_p), ("gpuv", C.c_void_p),
                ("gp2", C.c_void_p), ("nsc2", C.c_int),
                ("gp3a", C.c_void_p), ("nsc3a_lev", C.c_int), ("nsc3a_fld", C.c_int),
                ("gp3b", C.c_void_p), ("nsc3b_lev", C.c_int), ("nsc3b_fld", C.c_int),
                ("spvor", C.c_void_p), ("spdiv", C.c_void_p), ("spscalar", C.c_void_p),
                ("spsc2", C.c_void_p), ("spsc3a", C.c_void_p), ("spsc3b", C.c_void_p)]


class _VsetArgs(C.Structure):
    _fields_ = [("kvsetuv", C.c_void_p), ("nuv_g", C.c_int), ("kvsetsc", C.c_void_p), ("nscalar_g", C.c_int),
                ("kvsetsc2", C.c_void_p), ("nsc2_g", C.c_int), ("kvsetsc3a", C.c_void_p), ("nsc3a_lev_g", C.c_int),
                ("kvsetsc3b", C.c_void_p), ("nsc3b_lev_g", C.c_int)]


class Timings(C.Structure):
    _fields_ = [("h2d", C.c_float), ("prologue", C.c_float), ("legendre", C.c_float), ("transpose", C.c_float),
                ("fourier", C.c_float), ("epilogue", C.c_float), ("d2h", C.c_float), ("total", C.c_float),
                ("launches", C.c_longlong)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTED_SYMBOLS = [
    "ect_setup", "ect_inquire", "ect_inquire_array", "ect_inv_trans", "ect_dir_trans", "ect_specnorm",
    "ect_get_timings", "ect_synchronize", "ect_release", "ect_finalize", "ect_strerror", "ect_last_error",
    "ect_nccl_unique_id", "ect_host_alloc", "ect_host_free", "ect_debug_get_table", "ect_measure_fp64_peak",
    "ect_gath_grid", "ect_dist_grid", "ect_gath_spec", "ect_dist_spec", "ect_inv_transad", "ect_dir_transad",
    "ect_gpnorm_trans", "ect_vordiv_to_uv", "ect_inquire_rpnm", "ect_trans_pnm", "ect_write_legpol", "ect_read_legpol",
    "ect_gridpoint_partition", "ect_specnorm_met", "ect_inv_trans_vset", "ect_dir_trans_vset", "ect_specnorm_vset",
    "ect_comm_info",
]


def lib():
    """Load the CUDA library (fails loudly if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIBPATH):
            raise EctError(f"{_LIBPATH} not found: build it with `make` (or __graft_entry__.build()); "
                           "there is no CPU fallback")
        L = C.CDLL(_LIBPATH, mode=C.RTLD_GLOBAL)
        L.ect_strerror.restype = C.c_char_p
        L.ect_last_error.restype = C.c_char_p
        L.ect_setup.argtypes = [C.POINTER(_SetupOpts), C.POINTER(C.c_int)]
        L.ect_inquire.argtypes = [C.c_int, C.POINTER(Info)]
        L.ect_inquire_array.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_longlong]
        L.ect_inv_trans.argtypes = [C.c_int, C.POINTER(_InvArgs)]
        L.ect_dir_trans.argtypes = [C.c_int, C.POINTER(_DirArgs)]
        L.ect_inv_transad.argtypes = [C.c_int, C.POINTER(_InvArgs)]
        L.ect_dir_transad.argtypes = [C.c_int, C.POINTER(_DirArgs)]
        L.ect_specnorm.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ect_specnorm_met.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ect_specnorm_vset.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ect_inv_trans_vset.argtypes = [C.c_int, C.POINTER(_InvArgs), C.POINTER(_VsetArgs)]
        L.ect_dir_trans_vset.argtypes = [C.c_int, C.POINTER(_DirArgs), C.POINTER(_VsetArgs)]
        L.ect_get_timings.argtypes = [C.c_int, C.POINTER(Timings)]
        L.ect_comm_info.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_longlong)]
        L.ect_debug_get_table.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong]
        L.ect_measure_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
        L.ect_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_longlong]
        L.ect_host_free.argtypes = [C.c_void_p]
        L.ect_nccl_unique_id.argtypes = [C.c_void_p]
        L.ect_gath_grid.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ect_dist_grid.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ect_gath_spec.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ect_dist_spec.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ect_gpnorm_trans.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ect_vordiv_to_uv.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ect_inquire_rpnm.argtypes = [C.c_int, C.c_void_p, C.c_longlong, C.POINTER(C.c_int), C.c_void_p]
        L.ect_trans_pnm.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.ect_write_legpol.argtypes = [C.c_int, C.c_char_p]
        L.ect_read_legpol.argtypes = [C.c_int, C.c_char_p]
        L.ect_gridpoint_partition.argtypes = [C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_longlong, C.POINTER(C.c_longlong)]
        _lib = L
    return _lib


def _check(rc, what):
    if rc != 0:
        L = lib()
        raise EctError(f"{what} failed: {L.ect_strerror(rc).decode()} ({rc}): {L.ect_last_error().decode()}")


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(ECT_NCCL_UID_BYTES)
    _check(lib().ect_nccl_unique_id(buf), "ect_nccl_unique_id")
    return buf.raw


def measure_fp64_peak(which: int) -> float:
    v = C.c_double(0.0)
    _check(lib().ect_measure_fp64_peak(which, C.byref(v)), "ect_measure_fp64_peak")
    return v.value


def gridpoint_partition(nloen, nproc: int):
    """The reference's default grid-point decomposition (LDEQ_REGIONS=T, LDSPLIT=T) for nproc tasks: (regions per
    band, [per task: array (npieces, 3) of (latitude, first point, count), 0-based, in local point order])."""
    nl = np.ascontiguousarray(nloen, dtype=np.int32)
    nb, ns = C.c_int(0), C.c_longlong(0)
    reg = np.zeros(nproc, dtype=np.int32)
    seg0 = np.zeros(nproc + 1, dtype=np.int32)
    _check(lib().ect_gridpoint_partition(nl.size, nl.ctypes.data, nproc, C.byref(nb), reg.ctypes.data, seg0.ctypes.data,
                                         None, 0, C.byref(ns)), "ect_gridpoint_partition")
    segs = np.zeros((ns.value, 3), dtype=np.int32)
    _check(lib().ect_gridpoint_partition(nl.size, nl.ctypes.data, nproc, None, None, None, segs.ctypes.data, ns.value,
                                         None), "ect_gridpoint_partition")
    return reg[:nb.value].copy(), [segs[seg0[p]:seg0[p + 1]] for p in range(nproc)]


def vordiv_to_uv(nsmax: int, spvor, spdiv):
    """VORDIV_TO_UV without a resolution handle (the reference sets up a temporary spectral-only one): one task,
    double precision, arrays (nspec2g, nfld) with nspec2g = (nsmax+1)(nsmax+2).  Runs on the current CUDA device."""
    spvor = np.ascontiguousarray(spvor, dtype=np.float64); spdiv = np.ascontiguousarray(spdiv, dtype=np.float64)
    if spvor.shape[0] != (nsmax + 1) * (nsmax + 2) or spdiv.shape != spvor.shape:
        raise EctError("vordiv_to_uv: arrays must be (nspec2g, nfld)")
    u, v = np.zeros_like(spvor), np.zeros_like(spvor)
    _check(lib().ect_vordiv_to_uv(0, nsmax, spvor.ctypes.data, spdiv.ctypes.data, u.ctypes.data, v.ctypes.data,
                                  int(spvor.shape[1]), ECT_MEM_HOST), "ect_vordiv_to_uv")
    return u, v


def octahedral_nloen(n: int) -> np.ndarray:
    """O<N> grid (src/programs/ectrans-benchmark.F90:1043-1047)."""
    half = 20 + 4 * np.arange(n, dtype=np.int32)
    return np.concatenate([half, half[::-1]]).astype(np.int32)


def _is_torch(x):
    return x is not None and type(x).__module__.startswith("torch")


def _ptr(x):
    if x is None:
        return None
    if _is_torch(x):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(x.ctypes.data)


class PinnedArray:
    """float64 host array in page-locked memory from ect_host_alloc (the reference benchmark's
    pinned allocation, src/programs/util/ectrans_memory.c)."""

    def __init__(self, shape, dtype=np.float64):
        self.shape = tuple(int(s) for s in shape)
        n = int(np.prod(self.shape)) if self.shape else 1
        dt = np.dtype(dtype)
        self._p = C.c_void_p()
        _check(lib().ect_host_alloc(C.byref(self._p), n * dt.itemsize), "ect_host_alloc")
        buf = (C.c_char * max(n * dt.itemsize, 1)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=dt, count=n).reshape(self.shape)

    def free(self):
        if self._p:
            self.array = None
            lib().ect_host_free(self._p)
            self._p = None


class Transform:
    """One resolution handle (SETUP_TRANS) on one GPU / rank.

    Mirrors the reference's SETUP_TRANS / TRANS_INQ / INV_TRANS / DIR_TRANS / SPECNORM /
    TRANS_RELEASE (src/trans/include/ectrans/*.h).  Inputs may be NumPy arrays (host; the call
    copies H2D/D2H inside, like the reference GPU backend) or torch CUDA tensors (device resident,
    asynchronous on the handle's stream; call ``synchronize()``).
    """

    def __init__(self, nsmax, nloen, nranks=1, rank=0, device=-1, stream=None, nccl_uid=None, host_only=False,
                 precision="dp", legpol_read=None, legpol_write=None, gp_partition="latbands", nprtrv=1, bands="cost"):
        """legpol_read / legpol_write: SETUP_TRANS's CDIO_LEGPOL='readf' / 'writef' with CDLEGPOLFNAME (the reference's
        Legendre-polynomial cache file format); with legpol_read the table is not computed.
        gp_partition: "latbands" (native: the caller's grid points are the task's Fourier latitude band, TRLTOG / TRGTOL
        are local) or "eq_regions" (the reference's default LDEQ_REGIONS=T, LDSPLIT=T decomposition; TRLTOG / TRGTOL are
        NCCL all-to-alls).
        bands: "cost" (default; with more than one rank the Fourier latitude bands are balanced on NLOEN + 0.15 max(NLOEN))
        or "points" (the reference's SUMPLATB point count).
        nprtrv: NPRTRV; > 1: nranks is NPROC = NPRTRW * NPRTRV, rank pe is W-set pe // nprtrv, V-set pe % nprtrv, the
        grid-point arrays follow eq_regions over all tasks and carry every field (use inv_trans_vset / dir_trans_vset)."""
        self.nprtrv = int(nprtrv)
        if gp_partition not in ("latbands", "eq_regions"):
            raise EctError("gp_partition must be 'latbands' or 'eq_regions'")
        L = lib()
        self.precision = precision
        self.dtype = np.float64 if precision == "dp" else np.float32
        nl = np.ascontiguousarray(nloen, dtype=np.int32)
        self._uid = C.create_string_buffer(nccl_uid, ECT_NCCL_UID_BYTES) if nccl_uid else None
        o = _SetupOpts(int(nsmax), int(nl.size), nl.ctypes.data_as(C.POINTER(C.c_int)), int(nranks), int(rank),
                       (ECT_SETUP_HOST_ONLY if host_only else 0) | (ECT_SETUP_STREAM_GIVEN if stream is not None else 0)
                       | (ECT_SETUP_LEGPOL_DEFER if legpol_read else 0)
                       | (ECT_SETUP_GP_EQ_REGIONS if gp_partition == "eq_regions" else 0)
                       | (ECT_SETUP_BANDS_BY_POINTS if bands == "points" else 0)
                       | ((int(nprtrv) & 0xff) << 8 if int(nprtrv) > 1 else 0),
                       int(device), C.c_void_p(stream) if stream else None,
                       C.cast(self._uid, C.c_void_p) if self._uid else None,
                       ECT_PREC_DP if precision == "dp" else ECT_PREC_SP)
        h = C.c_int(0)
        _check(L.ect_setup(C.byref(o), C.byref(h)), "ect_setup")
        self.handle = h.value
        self.rank_world = int(rank)
        if legpol_read:
            rc = L.ect_read_legpol(self.handle, os.fsencode(legpol_read))
            if rc:
                msg = (L.ect_last_error() or b"").decode()
                L.ect_release(self.handle); self.handle = 0
                raise EctError(f"ect_read_legpol failed ({rc}): {msg}")
        if legpol_write:
            _check(L.ect_write_legpol(self.handle, os.fsencode(legpol_write)), "ect_write_legpol")
        self.info = Info()
        _check(L.ect_inquire(self.handle, C.byref(self.info)), "ect_inquire")
        i = self.info
        self.nsmax, self.ndgl, self.nspec2, self.ngptot = i.nsmax, i.ndgl, i.nspec2, i.ngptot
        self.nspec2g, self.ngptotg, self.nump = i.nspec2g, i.ngptotg, i.nump
        self.nranks, self.rank = i.nranks, i.rank
        self.nloen = self._arr(ARR_NLOEN, np.int32, i.ndgl)
        self.nmen = self._arr(ARR_NMEN, np.int32, i.ndgl)
        self.ndglu = self._arr(ARR_NDGLU, np.int32, i.nsmax + 1)
        self.myms = self._arr(ARR_MYMS, np.int32, i.nump)
        self.nasm0 = self._arr(ARR_NASM0, np.int32, i.nsmax + 1)
        self.nprocm = self._arr(ARR_NPROCM, np.int32, i.nsmax + 1)
        self.rmu = self._arr(ARR_RMU, np.float64, i.ndgl)
        self.rgw = self._arr(ARR_RGW, np.float64, i.ndgl)
        self.racthe = self._arr(ARR_RACTHE, np.float64, i.ndgl)
        self.lat_first = self._arr(ARR_LATFIRST, np.int32, i.nranks)
        self.lat_count = self._arr(ARR_LATCOUNT, np.int32, i.nranks)
        self.send_cnt = self._arr(ARR_SENDCNT, np.int64, i.nranks)
        self.recv_cnt = self._arr(ARR_RECVCNT, np.int64, i.nranks)
        self.send_off = self._arr(ARR_SENDOFF, np.int64, i.nranks)
        self.recv_off = self._arr(ARR_RECVOFF, np.int64, i.nranks)
        nseg = int(self._arr(28, np.int32, 1)[0])
        self.gp_segs = self._arr(27, np.int32, 3 * nseg).reshape(nseg, 3)      # (latitude, first point, count) of my grid points
        self.n_regions = self._arr(29, np.int32, i.nranks)

    def record_tables(self):
        """Fourier-buffer record tables of this rank (test / diagnostic access)."""
        i = self.info
        mrow0 = self._arr(ARR_MROW0, np.int64, i.nump + 1)
        latrow0 = self._arr(ARR_LATROW0, np.int64, i.nlat + 1)
        return {"mrow0": mrow0, "latrow0": latrow0,
                "leg_rec_n": self._arr(ARR_LEGRECN, np.int32, int(mrow0[-1])),
                "leg_rec_s": self._arr(ARR_LEGRECS, np.int32, int(mrow0[-1])),
                "fft_rec": self._arr(ARR_FFTREC, np.int32, int(latrow0[-1])),
                "leg_dst_rank_n": self._arr(ARR_LEGDSTRANKN, np.int32, int(mrow0[-1])),
                "leg_dst_rec_n": self._arr(ARR_LEGDSTRECN, np.int32, int(mrow0[-1])),
                "leg_dst_rank_s": self._arr(ARR_LEGDSTRANKS, np.int32, int(mrow0[-1])),
                "leg_dst_rec_s": self._arr(ARR_LEGDSTRECS, np.int32, int(mrow0[-1])),
                "fft_dst_rank": self._arr(ARR_FFTDSTRANK, np.int32, int(latrow0[-1])),
                "fft_dst_rec": self._arr(ARR_FFTDSTREC, np.int32, int(latrow0[-1]))}

    def _arr(self, which, dtype, n):
        out = np.zeros(max(int(n), 1), dtype=dtype)
        _check(lib().ect_inquire_array(self.handle, which, out.ctypes.data, int(n)), "ect_inquire_array")
        return out[:int(n)]

    # ------------------------------------------------------------------
    def gp_fields(self, nuv, nscalar, scders=False, vorgp=False, divgp=False, uvder=False):
        """Number of grid-point output fields of INV_TRANS (inv_trans.F90:356-367)."""
        n = 2 * nuv + nscalar
        if vorgp:
            n += nuv
        if divgp:
            n += nuv
        if scders:
            n += 2 * nscalar
        if uvder:
            n += 2 * nuv
        return n

    def _tdtype(self):
        import torch
        return torch.float64 if self.precision == "dp" else torch.float32

    def _blocks(self, nproma):
        nproma = self.ngptot if (nproma <= 0 or nproma >= self.ngptot) else nproma
        return nproma, (self.ngptot + nproma - 1) // max(nproma, 1)

    def inv_trans(self, spvor=None, spdiv=None, spscalar=None, scders=False, vorgp=False, divgp=False,
                  uvder=False, nproma=0, out=None):
        """INV_TRANS, call mode 1.  Returns gp ``(ngpblks, nfld, nproma)``."""
        dev = _is_torch(spvor) or _is_torch(spscalar)
        nuv = 0 if spvor is None else int(spvor.shape[1])
        nsc = 0 if spscalar is None else int(spscalar.shape[1])
        for a in (spvor, spdiv, spscalar):
            if a is not None:
                assert a.shape[0] == self.nspec2, "spectral arrays are (nspec2, nfld)"
        nfld = self.gp_fields(nuv, nsc, scders, vorgp, divgp, uvder)
        nproma, nblk = self._blocks(nproma)
        if out is None:
            if dev:
                import torch
                ref = spvor if spvor is not None else spscalar
                out = torch.empty((nblk, nfld, nproma), dtype=self._tdtype(), device=ref.device)
            else:
                out = np.empty((nblk, nfld, nproma), dtype=self.dtype)
        if not dev:
            spvor, spdiv, spscalar = (None if a is None else np.ascontiguousarray(a, dtype=self.dtype)
                                      for a in (spvor, spdiv, spscalar))
        a = _InvArgs()
        a.memspace = ECT_MEM_DEVICE if dev else ECT_MEM_HOST
        a.nproma = nproma
        a.scders, a.vorgp, a.divgp, a.uvder = int(scders), int(vorgp), int(divgp), int(uvder)
        a.spvor, a.spdiv, a.nuv = _ptr(spvor), _ptr(spdiv), nuv
        a.spscalar, a.nscalar = _ptr(spscalar), nsc
        a.gp = _ptr(out)
        self._keep = (spvor, spdiv, spscalar, out)
        _check(lib().ect_inv_trans(self.handle, C.byref(a)), "ect_inv_trans")
        return out

    def dir_trans(self, gp, nuv=0, nscalar=0, nproma=0, out=None):
        """DIR_TRANS, call mode 1.  gp ``(ngpblks, 2*nuv+nscalar, nproma)`` ordered u, v, scalars.
        Returns (spvor, spdiv, spscalar) each ``(nspec2, nfld)`` or None."""
        dev = _is_torch(gp)
        nproma, nblk = self._blocks(nproma)
        assert tuple(gp.shape) == (nblk, 2 * nuv + nscalar, nproma), (tuple(gp.shape), (nblk, 2 * nuv + nscalar, nproma))
        if dev:
            import torch
            mk = lambda n: torch.empty((self.nspec2, n), dtype=self._tdtype(), device=gp.device) if n else None
        else:
            gp = np.ascontiguousarray(gp, dtype=self.dtype)
            mk = lambda n: np.empty((self.nspec2, n), dtype=self.dtype) if n else None
        if out is None:
            out = (mk(nuv), mk(nuv), mk(nscalar))
        spvor, spdiv, spsc = out
        a = _DirArgs()
        a.memspace = ECT_MEM_DEVICE if dev else ECT_MEM_HOST
        a.nproma, a.nuv, a.nscalar = nproma, nuv, nscalar
        a.gp = _ptr(gp)
        a.spvor, a.spdiv, a.spscalar = _ptr(spvor), _ptr(spdiv), _ptr(spsc)
        self._keep = (gp, out)
        _check(lib().ect_dir_trans(self.handle, C.byref(a)), "ect_dir_trans")
        return spvor, spdiv, spsc

    def inv_transad(self, gp, nuv=0, nscalar=0, nproma=0, out=None):
        """INV_TRANSAD: adjoint of ``inv_trans`` (no derivative options) for the inner products of the reference's adjoint
        tests.  gp ``(ngpblks, 2*nuv+nscalar, nproma)`` ordered u, v, scalars; the results are ADDED to ``out`` =
        (spvor, spdiv, spscalar) (zero arrays are created when ``out`` is None)."""
        dev = _is_torch(gp)
        nproma, nblk = self._blocks(nproma)
        assert tuple(gp.shape) == (nblk, 2 * nuv + nscalar, nproma)
        if dev:
            import torch
            mk = lambda n: torch.zeros((self.nspec2, n), dtype=self._tdtype(), device=gp.device) if n else None
        else:
            gp = np.ascontiguousarray(gp, dtype=self.dtype)
            mk = lambda n: np.zeros((self.nspec2, n), dtype=self.dtype) if n else None
        if out is None:
            out = (mk(nuv), mk(nuv), mk(nscalar))
        spvor, spdiv, spsc = out
        a = _InvArgs()
        a.memspace = ECT_MEM_DEVICE if dev else ECT_MEM_HOST
        a.nproma, a.nuv, a.nscalar = nproma, nuv, nscalar
        a.gp = _ptr(gp)
        a.spvor, a.spdiv, a.spscalar = _ptr(spvor), _ptr(spdiv), _ptr(spsc)
        self._keep = (gp, out)
        _check(lib().ect_inv_transad(self.handle, C.byref(a)), "ect_inv_transad")
        return spvor, spdiv, spsc

    def dir_transad(self, spvor=None, spdiv=None, spscalar=None, nproma=0, out=None):
        """DIR_TRANSAD: adjoint of ``dir_trans``.  Returns gp ``(ngpblks, 2*nuv+nscalar, nproma)`` ordered u, v, scalars."""
        dev = _is_torch(spvor) or _is_torch(spscalar)
        nuv = 0 if spvor is None else int(spvor.shape[1])
        nsc = 0 if spscalar is None else int(spscalar.shape[1])
        nproma, nblk = self._blocks(nproma)
        if out is None:
            if dev:
                import torch
                ref = spvor if spvor is not None else spscalar
                out = torch.empty((nblk, 2 * nuv + nsc, nproma), dtype=self._tdtype(), device=ref.device)
            else:
                out = np.empty((nblk, 2 * nuv + nsc, nproma), dtype=self.dtype)
        if not dev:
            spvor, spdiv, spscalar = (None if a is None else np.ascontiguousarray(a, dtype=self.dtype)
                                      for a in (spvor, spdiv, spscalar))
        a = _DirArgs()
        a.memspace = ECT_MEM_DEVICE if dev else ECT_MEM_HOST
        a.nproma, a.nuv, a.nscalar = nproma, nuv, nsc
        a.gp = _ptr(out)
        a.spvor, a.spdiv, a.spscalar = _ptr(spvor), _ptr(spdiv), _ptr(spscalar)
        self._keep = (spvor, spdiv, spscalar, out)
        _check(lib().ect_dir_transad(self.handle, C.byref(a)), "ect_dir_transad")
        return out

    def inv_trans_raw(self, **kw):
        """Direct access to every ect_inv_args member (call mode 2 arrays etc.)."""
        a = _InvArgs()
        keep = []
        for k, v in kw.items():
            if hasattr(v, "shape"):
                keep.append(v)
                setattr(a, k, _ptr(v))
            else:
                setattr(a, k, v)
        self._keep = keep
        _check(lib().ect_inv_trans(self.handle, C.byref(a)), "ect_inv_trans")

    def dir_trans_raw(self, **kw):
        a = _DirArgs()
        keep = []
        for k, v in kw.items():
            if hasattr(v, "shape"):
                keep.append(v)
                setattr(a, k, _ptr(v))
            else:
                setattr(a, k, v)
        self._keep = keep
        _check(lib().ect_dir_trans(self.handle, C.byref(a)), "ect_dir_trans")

    def _vset_raw(self, fn, args, kw, vsets):
        keep = []
        for k, v in kw.items():
            if hasattr(v, "shape"):
                keep.append(v); setattr(args, k, _ptr(v))
            else:
                setattr(args, k, v)
        va = _VsetArgs()
        for k, v in vsets.items():                  # kvsetuv, kvsetsc, kvsetsc2, kvsetsc3a, kvsetsc3b: 1-based V-set arrays
            arr = np.ascontiguousarray(v, dtype=np.int32); keep.append(arr)
            setattr(va, k, arr.ctypes.data if arr.size else None)
            setattr(va, {"kvsetuv": "nuv_g", "kvsetsc": "nscalar_g", "kvsetsc2": "nsc2_g", "kvsetsc3a": "nsc3a_lev_g",
                         "kvsetsc3b": "nsc3b_lev_g"}[k], int(arr.size))
        self._keep = keep
        _check(fn(self.handle, C.byref(args), C.byref(va)), "ect_*_trans_vset")

    def inv_trans_vset_raw(self, vsets, **kw):
        """Every ect_inv_args member + the V-set arrays (call mode 2 with NPRTRV > 1); numpy arrays = host memory
        (set memspace=ECT_MEM_HOST), torch CUDA tensors = device pointers."""
        self._vset_raw(lib().ect_inv_trans_vset, _InvArgs(), kw, vsets)

    def dir_trans_vset_raw(self, vsets, **kw):
        self._vset_raw(lib().ect_dir_trans_vset, _DirArgs(), kw, vsets)

    def specnorm_vset(self, spec, kvset, pmet=None):
        """SPECNORM with KVSET: spec (nspec2, local fields of this V-set) -> norms of all len(kvset) global fields."""
        kv = np.ascontiguousarray(kvset, dtype=np.int32)
        spec = np.ascontiguousarray(spec, dtype=self.dtype)
        out = np.zeros(kv.size)
        met = None if pmet is None else np.ascontiguousarray(pmet, dtype=np.float64)
        _check(lib().ect_specnorm_vset(self.handle, spec.ctypes.data if spec.size else None, int(spec.shape[1]), ECT_MEM_HOST,
                                       None if met is None else met.ctypes.data, kv.ctypes.data, int(kv.size), out.ctypes.data),
               "ect_specnorm_vset")
        return out

    def specnorm(self, spec, pmet=None):
        """SPECNORM: spectral L2 norm per field (global over ranks); pmet: optional metric (0:nsmax)."""
        dev = _is_torch(spec)
        nf = int(spec.shape[1])
        if not dev:
            spec = np.ascontiguousarray(spec, dtype=self.dtype)
        out = np.zeros(nf)
        met = None if pmet is None else np.ascontiguousarray(pmet, dtype=np.float64)
        if met is not None and met.size != self.nsmax + 1:
            raise EctError("specnorm: pmet must have nsmax + 1 entries")
        _check(lib().ect_specnorm_met(self.handle, _ptr(spec), nf, ECT_MEM_DEVICE if dev else ECT_MEM_HOST,
                                      None if met is None else met.ctypes.data, out.ctypes.data), "ect_specnorm")
        return out

    # ---- V-sets (NPRTRV > 1): call mode 1, host arrays ----
    def _vset(self, kvsetuv, kvsetsc):
        ku = np.ascontiguousarray(kvsetuv if kvsetuv is not None else [], dtype=np.int32)
        ks = np.ascontiguousarray(kvsetsc if kvsetsc is not None else [], dtype=np.int32)
        va = _VsetArgs(ku.ctypes.data if ku.size else None, int(ku.size), ks.ctypes.data if ks.size else None, int(ks.size),
                       None, 0, None, 0, None, 0)
        return va, (ku, ks)

    def inv_trans_vset(self, spvor, spdiv, spscalar, kvsetuv, kvsetsc, nproma=0, **flags):
        """INV_TRANS with KVSETUV / KVSETSC (1-based V-set per global level / scalar): spvor, spdiv, spscalar hold this
        task's levels (nspec2, local); returns gp (ngpblks, all fields, nproma) on the task's eq_regions points."""
        va, keep = self._vset(kvsetuv, kvsetsc)
        nuv_g, nsc_g = keep[0].size, keep[1].size
        scd = bool(flags.get("scders")) and nsc_g > 0
        nfg = ((nuv_g if flags.get("vorgp") else 0) + (nuv_g if flags.get("divgp") else 0) + 2 * nuv_g + nsc_g
               + (nsc_g if scd else 0) + (2 * nuv_g if flags.get("uvder") else 0) + (nsc_g if scd else 0))
        npr, nblk = self._blocks(nproma)
        gp = np.zeros((nblk, nfg, npr), dtype=self.dtype)
        cv = lambda x: None if x is None or x.shape[1] == 0 else np.ascontiguousarray(x, dtype=self.dtype)
        spvor, spdiv, spscalar = cv(spvor), cv(spdiv), cv(spscalar)
        a = _InvArgs()
        a.memspace = ECT_MEM_HOST; a.nproma = nproma
        a.scders, a.vorgp, a.divgp, a.uvder = (int(bool(flags.get(k))) for k in ("scders", "vorgp", "divgp", "uvder"))
        if spvor is not None:
            a.spvor, a.spdiv, a.nuv = spvor.ctypes.data, spdiv.ctypes.data, spvor.shape[1]
        if spscalar is not None:
            a.spscalar, a.nscalar = spscalar.ctypes.data, spscalar.shape[1]
        a.gp = gp.ctypes.data
        _check(lib().ect_inv_trans_vset(self.handle, C.byref(a), C.byref(va)), "ect_inv_trans_vset")
        return gp

    def dir_trans_vset(self, gp, kvsetuv, kvsetsc, nproma=0):
        """DIR_TRANS with V-sets: gp (ngpblks, 2 nuv_g + nsc_g, nproma) -> (spvor, spdiv, spscalar) of this task's levels."""
        va, keep = self._vset(kvsetuv, kvsetsc)
        me = self.rank_world % self.nprtrv + 1
        nuv_l, nsc_l = int((keep[0] == me).sum()), int((keep[1] == me).sum())
        gp = np.ascontiguousarray(gp, dtype=self.dtype)
        ov, od = np.zeros((self.nspec2, nuv_l), dtype=self.dtype), np.zeros((self.nspec2, nuv_l), dtype=self.dtype)
        os_ = np.zeros((self.nspec2, nsc_l), dtype=self.dtype)
        a = _DirArgs()
        a.memspace = ECT_MEM_HOST; a.nproma = nproma; a.nuv = nuv_l; a.nscalar = nsc_l
        a.gp = gp.ctypes.data
        if nuv_l:
            a.spvor, a.spdiv = ov.ctypes.data, od.ctypes.data
        if nsc_l:
            a.spscalar = os_.ctypes.data
        _check(lib().ect_dir_trans_vset(self.handle, C.byref(a), C.byref(va)), "ect_dir_trans_vset")
        return ov, od, os_

    def gpnorm_trans(self, gp, nproma=0, ave_only=False, pmin=None, pmax=None):
        """GPNORM_TRANS: (average, minimum, maximum) per field of gp (ngpblks, nfld, nproma); global over ranks.
        ave_only: pmin / pmax carry the local extrema in (LDAVE_ONLY)."""
        dev = _is_torch(gp)
        if not dev:
            gp = np.ascontiguousarray(gp, dtype=self.dtype)
        nf = int(gp.shape[1])
        ave = np.zeros(nf)
        mn = np.zeros(nf) if pmin is None else np.ascontiguousarray(pmin, dtype=np.float64).copy()
        mx = np.zeros(nf) if pmax is None else np.ascontiguousarray(pmax, dtype=np.float64).copy()
        _check(lib().ect_gpnorm_trans(self.handle, _ptr(gp), nf, nproma, ECT_MEM_DEVICE if dev else ECT_MEM_HOST,
                                      ave.ctypes.data, mn.ctypes.data, mx.ctypes.data, int(bool(ave_only))), "ect_gpnorm_trans")
        return ave, mn, mx

    def vordiv_to_uv(self, spvor, spdiv):
        """VORDIV_TO_UV on this handle's wavenumbers: (nspec2, nfld) vor, div -> U cos(theta), V cos(theta)."""
        dev = _is_torch(spvor)
        if dev:
            import torch
            u, v = torch.empty_like(spvor), torch.empty_like(spvor)
        else:
            spvor = np.ascontiguousarray(spvor, dtype=self.dtype); spdiv = np.ascontiguousarray(spdiv, dtype=self.dtype)
            u, v = np.zeros_like(spvor), np.zeros_like(spvor)
        _check(lib().ect_vordiv_to_uv(self.handle, self.nsmax, _ptr(spvor), _ptr(spdiv), _ptr(u), _ptr(v), int(spvor.shape[1]),
                                      ECT_MEM_DEVICE if dev else ECT_MEM_HOST), "ect_vordiv_to_uv")
        return u, v

    def legendre_polynomials(self):
        """TRANS_INQ(PRPNM, KPMS): (rpnm (nspolegl, ndgnh) = PRPNM(ndgnh, nspolegl) column major, npms (nsmax+1))."""
        n = C.c_int(0)
        npms = np.zeros(self.nsmax + 1, dtype=np.int32)
        _check(lib().ect_inquire_rpnm(self.handle, None, 0, C.byref(n), npms.ctypes.data), "ect_inquire_rpnm")
        out = np.zeros((n.value, self.ndgl // 2))
        _check(lib().ect_inquire_rpnm(self.handle, out.ctypes.data, out.size, None, None), "ect_inquire_rpnm")
        return out, npms

    def trans_pnm(self, m):
        """TRANS_PNM: polynomials of one wavenumber, (nsmax - m + 3, ndgnh) = PRPNM(ndgnh, nsmax - m + 3) column major."""
        out = np.zeros((self.nsmax - m + 3, self.ndgl // 2))
        _check(lib().ect_trans_pnm(self.handle, m, out.ctypes.data, self.ndgl // 2, self.nsmax - m + 3), "ect_trans_pnm")
        return out

    # ---- GATH_GRID / DIST_GRID / GATH_SPEC / DIST_SPEC (host arrays; collective over the ranks of the handle) ----
    def _owners(self, nfld, owner):
        own = np.zeros(nfld, dtype=np.int32) if owner is None else np.ascontiguousarray(np.broadcast_to(owner, (nfld,)), dtype=np.int32)
        return own, int((own == self.rank).sum())

    def gath_grid(self, gp, kto=None, nproma=0):
        """GATH_GRID: gp (ngpblks, nfld, nproma) local -> (nfld_owned, ngptotg) on the owning ranks (kto: 0-based, default 0)."""
        gp = np.ascontiguousarray(gp, dtype=self.dtype)
        nfld = int(gp.shape[1])
        own, mine = self._owners(nfld, kto)
        out = np.zeros((mine, self.ngptotg), dtype=self.dtype)
        _check(lib().ect_gath_grid(self.handle, gp.ctypes.data, nfld, nproma, own.ctypes.data, out.ctypes.data if mine else None), "ect_gath_grid")
        return out

    def dist_grid(self, gpg, nfld, kfrom=None, nproma=0):
        """DIST_GRID: (nfld_owned, ngptotg) on the owning ranks -> local (ngpblks, nfld, nproma)."""
        own, mine = self._owners(nfld, kfrom)
        gpg = None if gpg is None else np.ascontiguousarray(gpg, dtype=self.dtype)
        npr, nblk = self._blocks(nproma)
        out = np.zeros((nblk, nfld, npr), dtype=self.dtype)
        _check(lib().ect_dist_grid(self.handle, gpg.ctypes.data if mine else None, nfld, nproma, own.ctypes.data, out.ctypes.data), "ect_dist_grid")
        return out

    def gath_spec(self, sp, kto=None):
        """GATH_SPEC: sp (nspec2, nfld) local -> (nspec2g, nfld_owned), m ascending then n ascending, Im(m = 0) zeroed."""
        sp = np.ascontiguousarray(sp, dtype=self.dtype)
        nfld = int(sp.shape[1])
        own, mine = self._owners(nfld, kto)
        out = np.zeros((self.nspec2g, mine), dtype=self.dtype)
        _check(lib().ect_gath_spec(self.handle, sp.ctypes.data, nfld, own.ctypes.data, out.ctypes.data if mine else None), "ect_gath_spec")
        return out

    def dist_spec(self, spg, nfld, kfrom=None):
        """DIST_SPEC: (nspec2g, nfld_owned) on the owning ranks -> local (nspec2, nfld)."""
        own, mine = self._owners(nfld, kfrom)
        spg = None if spg is None else np.ascontiguousarray(spg, dtype=self.dtype)
        out = np.zeros((self.nspec2, nfld), dtype=self.dtype)
        _check(lib().ect_dist_spec(self.handle, spg.ctypes.data if mine else None, nfld, own.ctypes.data, out.ctypes.data), "ect_dist_spec")
        return out

    def legendre_table(self, ml, parity):
        """Test access: P[k, i] for local wavenumber index ml; parity 0: n-m even, 1: odd."""
        m = int(self.myms[ml])
        k = (self.nsmax - m + 3) // 2 if parity == 0 else (self.nsmax - m + 2) // 2
        nl = int(self.ndglu[m])
        out = np.zeros((k, nl))
        _check(lib().ect_debug_get_table(self.handle, ml, parity, out.ctypes.data, out.size), "ect_debug_get_table")
        return out

    def timings(self):
        t = Timings()
        _check(lib().ect_get_timings(self.handle, C.byref(t)), "ect_get_timings")
        return t.as_dict()

    def synchronize(self):
        _check(lib().ect_synchronize(self.handle), "ect_synchronize")

    def comm_info(self):
        """(peer-memory transposition active, consumer-done barriers issued so far) -- ect_comm_info."""
        p2p, nb = C.c_int(0), C.c_longlong(0)
        _check(lib().ect_comm_info(self.handle, C.byref(p2p), C.byref(nb)), "ect_comm_info")
        return {"peer_memory": bool(p2p.value), "entry_barriers": int(nb.value)}

    def release(self):
        if getattr(self, "handle", 0):
            lib().ect_release(self.handle)
            self.handle = 0

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass
