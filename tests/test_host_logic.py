"""Host-side logic of the CUDA library through the C ABI, no GPU needed: the library loads and
exports every symbol include/*.h declares, SETUP_TRANS geometry/decomposition agree with the oracle,
errors come back as codes (never abort), and compute calls fail loudly without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

import ectrans_oracle as eo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def eb(built):
    import ectrans_b200
    return ectrans_b200


def test_exports_match_headers(eb):
    L = eb.lib()
    names = set()
    for h in os.listdir(os.path.join(ROOT, "include")):
        if not h.endswith(".h"):
            continue
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"^\s*(?:const\s+)?(?:int|void|char\s*\*|const char\s*\*|struct\s+\w+\s*\*?|\w+_t\s*\*?)\s+\**((?:ect|trans)_\w+)\s*\(", src, flags=re.M))
    assert len(names) >= 17
    for n in sorted(names):
        assert hasattr(L, n), f"{n} declared in include/ but not exported"
    for n in eb.EXPORTED_SYMBOLS:
        assert n in names


@pytest.mark.parametrize("T,N", [(47, 48), (79, 80), (159, 160), (399, 400)])
def test_geometry_matches_oracle(eb, T, N):
    nloen = eb.octahedral_nloen(N)
    t = eb.Transform(T, nloen, host_only=True)
    s = eo.setup(T, 2 * N, nloen, tables=False)
    np.testing.assert_array_equal(t.nmen, s.nmen)
    np.testing.assert_array_equal(t.ndglu, s.ndglu)
    np.testing.assert_array_equal(t.nasm0, s.nasm0)
    assert (t.nspec2, t.ngptot) == (s.nspec2, s.ngptot)
    assert np.abs(t.rmu - s.rmu).max() < 1e-15
    assert np.abs(t.rgw - s.rw).max() < 1e-15
    assert abs(t.rgw.sum() - 1.0) < 1e-10
    assert np.abs(t.racthe / s.racthe - 1).max() < 1e-14
    t.release()


N32_CLASSIC = [20, 27, 36, 40, 45, 50, 60, 64, 72, 75, 80, 90, 90, 96, 100, 108, 108, 120, 120, 120, 128, 128, 128, 128,
               128, 128, 128, 128, 128, 128, 128, 128]       # classic reduced grid, odd row lengths


@pytest.mark.parametrize("kind,T", [("O48", 95), ("O48", 63), ("F24", 47), ("F24", 30), ("golden150", 99), ("golden150", 60),
                                    ("N32classic", 63), ("N32classic", 42), ("N32classic", 21)])
def test_geometry_other_grids(eb, golden, kind, T):
    """NMEN rules of setup_geom_mod.F90:44-78: linear, quadratic and cubic branches, full and irregular reduced grids."""
    nloen = {"O48": eb.octahedral_nloen(48), "F24": np.full(48, 96, dtype=np.int32),
             "golden150": np.asarray(golden["nloen"], dtype=np.int32),
             "N32classic": np.asarray(N32_CLASSIC + N32_CLASSIC[::-1], dtype=np.int32)}[kind]
    t = eb.Transform(T, nloen, host_only=True)
    s = eo.setup(T, int(nloen.size), nloen, tables=False)
    np.testing.assert_array_equal(t.nmen, s.nmen)
    np.testing.assert_array_equal(t.ndglu, s.ndglu)
    assert (t.nspec2, t.ngptot) == (s.nspec2, s.ngptot)
    t.release()


def test_golden_grid_inquire(eb, golden):
    # tests/test_ectrans4py/test_ectrans4py.py:123-131
    t = eb.Transform(148, golden["nloen"], host_only=True)
    assert (t.ngptot, t.nspec2 // 2) == (33052, 11175)
    np.testing.assert_array_equal(t.nmen, golden["nmen"])
    t.release()


def test_cost_balanced_bands_tco1279(eb, monkeypatch):
    """The band cost model (host_plan.cu: FFT work + records + a constant per latitude, fitted to per-rank stage times on 8
    B200s) at the headline resolution: the partition the fit asks for, every rank agreeing on it, and the one-constant model
    of ECT_BAND_PAD for comparison."""
    T, N = 1279, 1280
    nloen = eb.octahedral_nloen(N)
    monkeypatch.delenv("ECT_BAND_PAD", raising=False)
    cnt = {}
    for P in (4, 8):
        trs = [eb.Transform(T, nloen, nranks=P, rank=r, host_only=True) for r in range(P)]
        cnt[P] = [int(c) for c in trs[0].lat_count]
        assert all([int(c) for c in t.lat_count] == cnt[P] for t in trs) and sum(cnt[P]) == 2 * N
        for t in trs:
            t.release()
    assert cnt[8] == [580, 275, 222, 203, 203, 222, 275, 580]
    assert cnt[4] == [855, 425, 425, 855]
    monkeypatch.setenv("ECT_BAND_PAD", "0.16")
    t = eb.Transform(T, nloen, nranks=8, rank=0, host_only=True)
    assert [int(c) for c in t.lat_count] == [565, 289, 227, 198, 200, 227, 289, 565]
    t.release()


@pytest.mark.parametrize("P", [2, 3, 4, 8])
def test_decomposition(eb, P):
    T, N = 159, 160
    nloen = eb.octahedral_nloen(N)
    trs = [eb.Transform(T, nloen, nranks=P, rank=r, host_only=True, bands="points") for r in range(P)]
    nprocm, myms = eo.suwavedi(T, P)
    first, count = eo.sumplatb_fourier(nloen, P)
    # default: the same SUMPLATB on cost weights (FFT work of a row + a per-latitude constant, host_plan.cu): contiguous
    # bands covering every latitude, symmetric about the equator, fewer polar latitudes per band than the point count gives
    tcs = [eb.Transform(T, nloen, nranks=P, rank=r, host_only=True) for r in range(P)]
    fc, cc = list(tcs[0].lat_first), list(tcs[0].lat_count)
    assert sum(cc) == 2 * N and fc[0] == 0 and all(fc[i + 1] == fc[i] + cc[i] for i in range(P - 1))
    assert all(abs(a - b) <= 2 for a, b in zip(cc, cc[::-1])) and cc[0] <= count[0] and all(list(t.lat_count) == cc for t in tcs)
    assert sum(t.ngptot for t in tcs) == tcs[0].ngptotg
    for t in tcs:
        t.release()
    for r, t in enumerate(trs):
        np.testing.assert_array_equal(t.myms, myms[r])
        np.testing.assert_array_equal(t.nprocm, nprocm)
        assert list(t.lat_first) == first and list(t.lat_count) == count
    assert sum(t.nspec2 for t in trs) == trs[0].nspec2g
    assert sum(t.ngptot for t in trs) == trs[0].ngptotg
    S = np.array([t.send_cnt for t in trs])
    R = np.array([t.recv_cnt for t in trs])
    np.testing.assert_array_equal(S, R.T)          # what a sends to b is what b expects from a
    assert S.sum() == 2 * sum(int(trs[0].ndglu[m]) for m in range(T + 1))
    # record tables: every (m, lat) pair appears exactly once on each side
    for t in trs:
        rt = t.record_tables()
        n = int(t.send_cnt.sum())
        recs = np.concatenate([rt["leg_rec_n"], rt["leg_rec_s"]])
        assert sorted(recs.tolist()) == list(range(n))
        f = rt["fft_rec"]
        assert sorted(f[f >= 0].tolist()) == list(range(int(t.recv_cnt.sum())))
    # fused transposition: what rank a says about the destination of its records is what the destination expects
    tabs = [t.record_tables() for t in trs]
    ndgnh = trs[0].ndgl // 2
    for a, t in enumerate(trs):
        rt = tabs[a]
        for ml, m in enumerate(t.myms):
            nd = int(t.ndglu[m]); isl = ndgnh - nd
            for i in range(0, nd, 7):
                at = rt["mrow0"][ml] + i
                for g, rk, rc in ((isl + i, rt["leg_dst_rank_n"][at], rt["leg_dst_rec_n"][at]),
                                  (t.ndgl - 1 - (isl + i), rt["leg_dst_rank_s"][at], rt["leg_dst_rec_s"][at])):
                    b = trs[rk]
                    l = g - int(b.info.lat0)
                    assert 0 <= l < int(b.info.nlat)
                    assert tabs[rk]["fft_rec"][tabs[rk]["latrow0"][l] + m] == rc
        for l in range(0, int(t.info.nlat), 5):
            g = int(t.info.lat0) + l
            for m in range(0, int(t.nmen[g]) + 1, 3):
                at = rt["latrow0"][l] + m
                rk, rc = rt["fft_dst_rank"][at], rt["fft_dst_rec"][at]
                assert rk == t.nprocm[m]
                ml = list(trs[rk].myms).index(m)
                gn = g if g < ndgnh else t.ndgl - 1 - g
                i = gn - (ndgnh - int(t.ndglu[m]))
                key = "leg_rec_n" if g < ndgnh else "leg_rec_s"
                assert tabs[rk][key][tabs[rk]["mrow0"][ml] + i] == rc
    for t in trs:
        t.release()


def test_error_codes(eb):
    L = eb.lib()
    h = ctypes.c_int(0)
    nl = eb.octahedral_nloen(8)
    o = eb._SetupOpts(7, 15, nl.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), 1, 0, 1, -1, None, None)
    assert L.ect_setup(ctypes.byref(o), ctypes.byref(h)) == -4          # odd ndgl -> bad argument
    assert b"bad arguments" in L.ect_last_error()
    assert L.ect_setup(None, ctypes.byref(h)) == -3                     # missing
    assert L.ect_inquire(12345, ctypes.byref(eb.Info())) == -8          # invalid handle
    assert L.ect_release(12345) == -8
    assert L.ect_strerror(-2) == b"not implemented"
    odd = nl.copy(); odd[0] = odd[-1] = 21                              # odd row lengths are fine (classic reduced grids)
    o = eb._SetupOpts(7, 16, odd.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), 1, 0, 1, -1, None, None)
    assert L.ect_setup(ctypes.byref(o), ctypes.byref(h)) == 0 and L.ect_release(h.value) == 0
    skew = nl.copy(); skew[0] = 24
    o = eb._SetupOpts(7, 16, skew.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), 1, 0, 1, -1, None, None)
    assert L.ect_setup(ctypes.byref(o), ctypes.byref(h)) == -4          # grid not symmetric about the equator
    t = eb.Transform(7, nl, host_only=True)
    a = eb._InvArgs()
    assert L.ect_inv_trans(t.handle, ctypes.byref(a)) == -6             # host-only handle has no device state
    t.release()


def test_no_cpu_fallback(eb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(eb.EctError, match="no CPU fallback"):
        eb.Transform(47, eb.octahedral_nloen(48))


def test_gath_dist_single_rank(built):
    """GATH_GRID / DIST_GRID / GATH_SPEC / DIST_SPEC on one rank (host-only handle): unblocking of PGP(nproma, nfld,
    ngpblks) into PGPG(ngptotg, nfld), identity spectral order, zeroed Im(m = 0) (gath_spec_control_mod.F90:178-184),
    field ownership checks."""
    import ectrans_b200 as eb
    T, N = 21, 24
    tr = eb.Transform(T, eb.octahedral_nloen(N), host_only=True)
    rng = np.random.default_rng(0)
    sp = rng.standard_normal((tr.nspec2, 3))
    g = tr.gath_spec(sp)
    exp = sp.copy(); exp[1:2 * (T + 1):2] = 0
    assert np.array_equal(g, exp)
    assert np.array_equal(tr.dist_spec(g, 3), exp)
    for npr in (0, 100, 7):
        nproma, nblk = tr._blocks(npr)
        gp = rng.standard_normal((nblk, 4, nproma))
        G = tr.gath_grid(gp, nproma=npr)
        flat = gp.transpose(1, 0, 2).reshape(4, -1)[:, :tr.ngptot]
        assert G.shape == (4, tr.ngptotg) and np.array_equal(G, flat)
        back = tr.dist_grid(G, 4, nproma=npr)
        assert np.array_equal(back.transpose(1, 0, 2).reshape(4, -1)[:, :tr.ngptot], flat)
    with pytest.raises(eb.EctError):
        tr.gath_spec(sp, kto=[0, 1, 0])              # rank 1 does not exist
    tr.release()


@pytest.mark.parametrize("kind", ["O16", "O48", "F24", "golden150"])
def test_eq_regions_gridpoint_partition(eb, golden, kind):
    """The reference's default grid-point decomposition (LDEQ_REGIONS=T, LDSPLIT=T): the C++ host code (heap ordered)
    against the oracle's restatement of eq_regions / SUMPLATBEQ / SUSTAONL (point-by-point scan), plus the invariants
    SUSTAONL itself asserts (sustaonl_mod.F90:330-372: every point assigned exactly once) and the balance SUMPLATBEQ
    aims for (tasks differ by at most one point)."""
    nloen = {"O16": eb.octahedral_nloen(16), "O48": eb.octahedral_nloen(48), "F24": np.full(48, 96, dtype=np.int32),
             "golden150": np.asarray(golden["nloen"], dtype=np.int32)}[kind]
    for nproc in (1, 2, 3, 4, 5, 8, 12, 16, 32, 61):
        nreg, segs = eo.gridpoint_partition(nloen, nproc)
        reg, mine = eb.gridpoint_partition(nloen, nproc)
        assert list(reg) == list(nreg) and int(reg.sum()) == nproc
        cover = [np.zeros(int(n), dtype=np.int32) for n in nloen]
        for a, b in zip(segs, mine):
            np.testing.assert_array_equal(np.asarray(a, dtype=np.int32).reshape(-1, 3), b)
            assert (np.diff(b[:, 0]) > 0).all()                     # local order: latitude ascending, one piece per latitude
            for lat, first, cnt in b:
                cover[lat][first:first + cnt] += 1
        assert all((c == 1).all() for c in cover)
        npts = np.array([int(b[:, 2].sum()) for b in mine])
        assert npts.max() - npts.min() <= 1 and npts.sum() == int(np.sum(nloen))


def test_eq_regions_known_partitions():
    """Leopardi's recursive zonal equal-area partition of the sphere (what eq_regions_mod.F90 codes): polar caps are
    single regions, the collars of an even partition are symmetric about the equator."""
    assert eo.eq_regions(1) == [1] and eo.eq_regions(2) == [1, 1] and eo.eq_regions(3) == [1, 1, 1]
    assert eo.eq_regions(4) == [1, 2, 1] and eo.eq_regions(8) == [1, 6, 1] and eo.eq_regions(10) == [1, 4, 4, 1]
    assert eo.eq_regions(32) == [1, 6, 9, 9, 6, 1] and eo.eq_regions(100) == [1, 6, 11, 15, 17, 17, 15, 11, 6, 1]
    for n in range(1, 300):
        r = eo.eq_regions(n)
        assert sum(r) == n and (n % 2 == 1 or r == r[::-1]) and (n < 3 or (r[0] == 1 and r[-1] == 1))


@pytest.mark.parametrize("world,V", [(2, 2), (4, 2), (8, 2), (6, 3)])
def test_trltog_tables_vsets(eb, world, V):
    """NPRTRV > 1: tasks form a W x V grid (w = pe // V, v = pe % V); the grid-point partition is eq_regions over all
    tasks, the latitude bands belong to the W-groups.  Band-owner tables run over all tasks, task tables over the W owners."""
    T, N = 47, 48
    nloen = eb.octahedral_nloen(N)
    W = world // V
    trs = [eb.Transform(T, nloen, nranks=world, rank=r, host_only=True, nprtrv=V) for r in range(world)]
    ref_w = [eb.Transform(T, nloen, nranks=W, rank=w, host_only=True) for w in range(W)]
    reg, segs = eb.gridpoint_partition(nloen, world)
    latoff = np.concatenate([[0], np.cumsum(nloen)])
    local_global = []
    for p, t in enumerate(trs):
        w = p // V
        assert t.info.nranks == W and t.info.rank == w                      # the W-group's wavenumbers and latitude band
        np.testing.assert_array_equal(t.myms, ref_w[w].myms)
        assert (t.info.lat0, t.info.nlat) == (ref_w[w].info.lat0, ref_w[w].info.nlat)
        np.testing.assert_array_equal(t.gp_segs, segs[p])
        local_global.append(np.concatenate([latoff[l] + f + np.arange(c) for l, f, c in t.gp_segs]))
        assert t.ngptot == local_global[-1].size
    assert sum(t.ngptot for t in trs) == int(nloen.sum())
    for r, t in enumerate(trs):
        w = r // V
        band0 = latoff[t.info.lat0]
        nband = int(latoff[t.info.lat0 + t.info.nlat] - band0)
        xb_idx = t._arr(30, np.int32, nband); xb_off = t._arr(31, np.int64, world + 1)
        assert sorted(xb_idx.tolist()) == list(range(nband))
        for p, tp in enumerate(trs):
            xg_off = tp._arr(32, np.int64, W + 1)
            np.testing.assert_array_equal(band0 + xb_idx[xb_off[p]:xb_off[p + 1]], local_global[p][xg_off[w]:xg_off[w + 1]])
    for t in trs + ref_w:
        t.release()


@pytest.mark.parametrize("world", [2, 3, 5, 8])
def test_trltog_tables_eq_regions(eb, world):
    """TRLTOG / TRGTOL with the reference's grid-point decomposition, host logic only: the message a band owner packs
    for a task (band points XBIDX[XBOFF[p] ..]) must be, point for point, what the task expects from that owner
    (its local points [XGOFF[r], XGOFF[r+1])); every band point travels exactly once."""
    T, N = 47, 48
    nloen = eb.octahedral_nloen(N)
    trs = [eb.Transform(T, nloen, nranks=world, rank=r, host_only=True, gp_partition="eq_regions") for r in range(world)]
    reg, segs = eb.gridpoint_partition(nloen, world)
    latoff = np.concatenate([[0], np.cumsum(nloen)])
    assert sum(t.ngptot for t in trs) == int(nloen.sum())
    local_global = []          # task -> global index of its local points
    for p, t in enumerate(trs):
        np.testing.assert_array_equal(t.gp_segs, segs[p])
        np.testing.assert_array_equal(t.n_regions[:len(reg)], reg)
        local_global.append(np.concatenate([latoff[l] + f + np.arange(c) for l, f, c in t.gp_segs]) if len(t.gp_segs) else np.zeros(0, int))
        assert t.ngptot == local_global[-1].size
    for r, t in enumerate(trs):
        band0 = latoff[t.info.lat0]
        nband = int(latoff[t.info.lat0 + t.info.nlat] - band0)
        xb_idx = t._arr(30, np.int32, nband); xb_off = t._arr(31, np.int64, world + 1)
        assert sorted(xb_idx.tolist()) == list(range(nband))
        for p, tp in enumerate(trs):
            xg_off = tp._arr(32, np.int64, world + 1)
            sent = band0 + xb_idx[xb_off[p]:xb_off[p + 1]]
            np.testing.assert_array_equal(sent, local_global[p][xg_off[r]:xg_off[r + 1]])
    for t in trs:
        t.release()


def test_bench_reference_arm_runs_without_gpu():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) must work on a box without a GPU and
    print one JSON line carrying the contract's keys."""
    import json, subprocess, sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "T79_O80_L10",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "ms" and line["higher_is_better"] is False
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0


def test_gridpoint_partition_random_grids(eb):
    """Random symmetric reduced grids (even and odd row lengths, plateaus) x random task counts: the heap-ordered C++
    SUSTAONL must reproduce the oracle's point-by-point scan, ties included (equal angles -> northernmost latitude)."""
    rng = np.random.default_rng(2024)
    checked = 0
    for _ in range(40):
        nh = int(rng.integers(4, 28))
        half = np.sort(rng.integers(8, 200, size=nh))
        if rng.random() < 0.5:
            half = (half // 4) * 4 + 4                      # many equal angles between rows
        nloen = np.concatenate([half, half[::-1]]).astype(np.int32)
        nproc = int(rng.integers(2, 41))
        try:
            nreg, segs = eo.gridpoint_partition(nloen, nproc)
        except Exception:
            continue                                         # too many tasks for this grid: both sides refuse (below)
        reg, mine = eb.gridpoint_partition(nloen, nproc)
        assert list(reg) == list(nreg)
        for a, b in zip(segs, mine):
            np.testing.assert_array_equal(np.asarray(a, dtype=np.int32).reshape(-1, 3), b)
        checked += 1
    assert checked >= 25
    # SUMPLATBEQ's "NPROC TOO BIG FOR THIS RESOLUTION" (sumplatbeq_mod.F90:94-99) is an error code here, not an abort
    with pytest.raises(eb.EctError):
        eb.gridpoint_partition(np.full(4, 8, dtype=np.int32), 32)


def test_eq_regions_cxx_matches_reference_arithmetic(eb):
    """For odd task counts the ideal collar sizes come in exact-tie pairs, so eq_regions hangs on the last bits of the
    reference's arithmetic (its own gamma polynomial, 4*pi*sin(s/2)**2 association): the C++ must give the oracle's
    restatement for every count, e.g. eq_regions(31) = [1 6 8 9 6 1], not [1 6 9 8 6 1]."""
    L = eb.lib()
    nl = np.full(8192, 16, dtype=np.int32)
    for n in list(range(1, 300)) + [511, 777, 1001, 1023, 2047]:
        nb, ns = ctypes.c_int(0), ctypes.c_longlong(0)
        reg = np.zeros(n, dtype=np.int32); seg0 = np.zeros(n + 1, dtype=np.int32)
        assert L.ect_gridpoint_partition(8192, nl.ctypes.data, n, ctypes.byref(nb), reg.ctypes.data, seg0.ctypes.data, None, 0, ctypes.byref(ns)) == 0
        assert list(reg[:nb.value]) == eo.eq_regions(n), n
    # a band that would fit inside what is left of one latitude cannot be described: an error code, not a crash
    with pytest.raises(eb.EctError, match="NPROC TOO BIG"):
        eb.gridpoint_partition(np.full(64, 4096, dtype=np.int32), 1024)


def test_host_plan_fuzz_against_oracle(eb):
    """Random symmetric reduced grids (odd and even row lengths), truncations and task counts: NMEN / NDGLU / Gaussian
    latitudes, the wavenumber distribution (SUWAVEDI) and the Fourier latitude bands (SUMPLATB) of the C++ host plan
    against the oracle's restatements."""
    rng = np.random.default_rng(11)
    for it in range(120):
        nh = int(rng.integers(2, 50))
        half = np.sort(rng.integers(4, 400, size=nh))
        nloen = np.concatenate([half, half[::-1]]).astype(np.int32)
        T = int(rng.integers(1, 3 * nh))
        W = int(rng.integers(1, min(2 * nh, 12) + 1))
        trs = [eb.Transform(T, nloen, nranks=W, rank=r, host_only=True, bands="points") for r in range(W)]
        s = eo.setup(T, 2 * nh, nloen, tables=False)
        t = trs[0]
        np.testing.assert_array_equal(t.nmen, s.nmen)
        np.testing.assert_array_equal(t.ndglu, s.ndglu)
        assert np.abs(t.rmu - s.rmu).max() < 1e-15 and np.abs(t.rgw - s.rw).max() < 1e-15
        first, count = eo.sumplatb_fourier(nloen, W)[:2]
        np.testing.assert_array_equal(t.lat_first, np.asarray(first)); np.testing.assert_array_equal(t.lat_count, np.asarray(count))
        wave = eo.suwavedi(T, W)
        np.testing.assert_array_equal(t.nprocm, np.asarray(wave[0] if isinstance(wave, tuple) else wave))
        assert sum(x.nspec2 for x in trs) == (T + 1) * (T + 2) and sum(x.ngptot for x in trs) == int(nloen.sum())
        for x in trs:
            x.release()


def test_ectrans4py_model_order_permutation():
    """LREORDER=True of the reference's ectrans4py (sp2gp_gauss4py.F90:82-108): the 'model' order <-> ecTrans order maps,
    restated literally from the Fortran loops and compared with the vectorised helpers."""
    from ectrans_b200 import ectrans4py as e4
    T = 7
    size = (T + 1) * (T + 2)
    rng = np.random.default_rng(0)
    pspec = rng.standard_normal(size)
    nasm0, ji = {}, 1
    for n in range(T + 1):
        nasm0[n] = ji
        ji = ji + 1 + n + (n + 1)
    buf, ji = np.zeros(size), 0
    for m in range(T + 1):
        for n in range(m, T + 1):
            buf[ji] = pspec[nasm0[n] + m - 1]; ji += 1
            buf[ji] = 0.0 if m == 0 else pspec[nasm0[n] - m - 1]; ji += 1
    got = e4._from_model_order(T, pspec)
    assert np.array_equal(got, buf)
    back = e4._to_model_order(T, got, size)
    used = np.zeros(size, bool)
    re, im = e4._model_order(T)
    used[re] = True; used[im[im >= 0]] = True
    assert used.sum() == (T + 1) ** 2 and np.array_equal(back[used], pspec[used]) and np.all(back[~used] == 0)
