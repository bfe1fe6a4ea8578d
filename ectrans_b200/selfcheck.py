"""Oracle-free parity self-check of a (possibly distributed) run, for ``bench.py --gpus N`` and the multi-GPU tools.

Nothing here imports ``oracle/``: the checks are anchored on
  * the reference's own golden vectors (tests/golden: ectrans4py's ``tl149-c24-s1t@sp{,2gp}.npy``, abs tol 1e-10,
    /root/reference/tests/test_ectrans4py/test_ectrans4py.py:16,143-160),
  * bit identity between the N-rank transform and the one-rank transform of the same global fields on the same GPU
    (the property behind the reference's --dump-checksums harness, src/programs/ectrans-benchmark.F90:1455-1638,
    tests/compare_checksums.py:22-46),
  * the benchmark's analytic input (single harmonic Re psi(4,19) = 1, ectrans-benchmark.F90:1389-1415) and its
    round-trip criterion (<= 100 eps relative spectral-norm error, :847-871).
The stress part issues several transforms of the SAME direction back to back on device pointers (no host
synchronisation, one rank delayed) and through the chunked host path -- the call sequences that need the
consumer-done barrier of the peer-memory transposition (csrc/api.cu ect_transpose_enter).
"""
from __future__ import annotations

import os

import numpy as np

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _goff(T, m):
    """0-based offset of wavenumber m in the global m-major spectral vector (NASM0 of a one-task run)."""
    return 2 * (m * (T + 1) - m * (m - 1) // 2)


def _spec_index(tr):
    """Global spectral indices of this rank's coefficients, in local (MYMS) order."""
    T = tr.nsmax
    if tr.nump == 0:
        return np.zeros(0, dtype=np.int64)
    return np.concatenate([_goff(T, int(m)) + np.arange(2 * (T - int(m) + 1)) for m in tr.myms])


def _grid_index(tr):
    """Global grid-point indices of this rank's points (native latitude-band partition)."""
    off = np.concatenate([[0], np.cumsum(tr.nloen.astype(np.int64))])
    if len(tr.gp_segs) == 0:
        return np.zeros(0, dtype=np.int64)
    return np.concatenate([off[l] + f + np.arange(c) for l, f, c in tr.gp_segs])


def run(eb, world, rank, local, uid_fn, stress_T=399, skew_ms=30.0, verbose=False):
    """Returns a dict of parity figures for this rank; combine over ranks with ``reduce``."""
    import torch
    dev = torch.device("cuda", local)
    res = {}
    mk = lambda T, nloen, **kw: eb.Transform(T, nloen, nranks=world, rank=rank, device=local,
                                             nccl_uid=uid_fn() if world > 1 else None, **kw)
    # ---- 1. reference golden vectors through the distributed transform ----
    nl = np.load(os.path.join(_GOLDEN, "lon_number_by_lat.npy")).astype(np.int32)
    sp = np.load(os.path.join(_GOLDEN, "tl149-c24-s1t@sp.npy"))
    gpl = np.load(os.path.join(_GOLDEN, "tl149-c24-s1t@sp2gp.npy"))
    gpref = np.concatenate([gpl[i, :nl[i]] for i in range(nl.size)])
    tr = mk(148, nl)
    si, gi = _spec_index(tr), _grid_index(tr)
    gp = tr.inv_trans(spscalar=np.ascontiguousarray(sp[si][:, None]))
    res["golden_inv_maxabs"] = float(np.abs(gp[0, 0] - gpref[gi]).max()) if gi.size else 0.0
    _, _, so = tr.dir_trans(np.ascontiguousarray(gpref[gi][None, None, :]), 0, 1)
    res["golden_dir_maxabs"] = float(np.abs(so[:, 0] - sp[si]).max()) if si.size else 0.0
    res["peer_memory"] = bool(tr.comm_info()["peer_memory"]) if world > 1 else False
    tr.release()

    # ---- 2. T159 / O160 with every derivative option: N ranks == one rank, bit for bit; benchmark harmonic ----
    T, N = 159, 160
    nloen = eb.octahedral_nloen(N)
    tr, tr1 = mk(T, nloen), eb.Transform(T, nloen, device=local)
    si, gi = _spec_index(tr), _grid_index(tr)
    rng = np.random.default_rng(7)
    nuv, nsc = 3, 4
    def spec(n):
        a = rng.uniform(-0.1, 0.1, size=(tr1.nspec2, n))
        a[1:2 * (T + 1):2] = 0.0            # Im(m = 0)
        return a
    vor, div, sc = spec(nuv), spec(nuv), spec(nsc)
    vor[0:2] = 0.0; div[0:2] = 0.0          # (0, 0) of vorticity / divergence
    opts = dict(scders=True, vorgp=True, divgp=True, uvder=True)
    g1 = tr1.inv_trans(vor, div, sc, **opts)
    gN = tr.inv_trans(*(np.ascontiguousarray(a[si]) for a in (vor, div, sc)), **opts)
    bit = bool(np.array_equal(gN[0], g1[0][:, gi]))
    nv = 2 * nuv          # u, v follow vor, div in the output list
    uvsc1 = np.ascontiguousarray(g1[:, nv:nv + 2 * nuv + nsc])
    o1 = tr1.dir_trans(uvsc1, nuv, nsc)
    oN = tr.dir_trans(np.ascontiguousarray(uvsc1[:, :, gi]), nuv, nsc)
    bit = bit and all(np.array_equal(a, b[si]) for a, b in zip(oN, o1))
    res["t159_bit_identical"] = bit
    # round trip of the one-rank result
    n_in = tr1.specnorm(sc); n_out = tr1.specnorm(o1[2])
    res["t159_roundtrip_norm_rel"] = float(np.abs(n_out / n_in - 1.0).max())          # white random spectrum: <= 1e-12
    # benchmark input: Re psi(m=4, n=19) = 1; its round trip must keep the spectral norm to 100 eps (--check 100)
    h = np.zeros((tr1.nspec2, 1)); h[_goff(T, 4) + 2 * (19 - 4), 0] = 1.0
    gh = tr1.inv_trans(spscalar=h)
    _, _, hb = tr1.dir_trans(gh, 0, 1)
    res["harmonic_roundtrip_maxabs"] = float(np.abs(hb - h).max())
    res["harmonic_roundtrip_norm_err_eps"] = float(abs(tr1.specnorm(hb)[0] / tr1.specnorm(h)[0] - 1.0) / np.finfo(np.float64).eps)
    tr.release(); tr1.release()

    # ---- 3. stress: same-direction transforms back to back on device pointers, one rank delayed; chunked host path ----
    T = N = None
    T, N = stress_T, stress_T + 1
    nloen = eb.octahedral_nloen(N)
    stream = torch.cuda.current_stream().cuda_stream
    tr = eb.Transform(T, nloen, nranks=world, rank=rank, device=local, stream=stream, nccl_uid=uid_fn() if world > 1 else None)
    tr1 = eb.Transform(T, nloen, device=local, stream=stream)
    si = torch.from_numpy(_spec_index(tr)).to(dev)
    gi = torch.from_numpy(_grid_index(tr)).to(dev)
    nuv, nsc = 12, 20                       # 44 Legendre fields: two chunks on the host path
    g = torch.Generator(device=dev); g.manual_seed(99)          # same global fields on every rank
    rnd = lambda n: (torch.rand((tr1.nspec2, n), generator=g, device=dev, dtype=torch.float64) - 0.5) * 0.2
    ins = [(rnd(nuv), rnd(nuv), rnd(nsc)) for _ in range(3)]
    ref_gp = [tr1.inv_trans(*x).clone() for x in ins]
    tr1.synchronize()
    loc_in = [tuple(a[si].contiguous() for a in x) for x in ins]
    outs = [torch.empty((1, 2 * nuv + nsc, tr.ngptot), dtype=torch.float64, device=dev) for _ in range(3)]
    late = world > 1 and rank == world - 1
    clk = 1.9e6 * skew_ms                   # ~ SM cycles
    torch.cuda.synchronize()
    for k in range(3):
        if late and k == 1:
            torch.cuda._sleep(int(clk))     # this rank reaches the 2nd transform late: its peers run ahead
        if (not late) and world > 1 and k == 2 and rank == 0:
            torch.cuda._sleep(int(clk / 2))
        tr.inv_trans(*loc_in[k], out=outs[k])
    tr.synchronize(); torch.cuda.synchronize()
    ok_inv = all(bool(torch.equal(outs[k][0], ref_gp[k][0][:, gi])) for k in range(3))
    ref_sp = [tuple(t.clone() for t in tr1.dir_trans(ref_gp[k], nuv, nsc)) for k in range(3)]
    tr1.synchronize()
    loc_gp = [ref_gp[k][:, :, gi].contiguous() for k in range(3)]
    souts = []
    torch.cuda.synchronize()
    for k in range(3):
        if late and k == 1:
            torch.cuda._sleep(int(clk))
        souts.append(tr.dir_trans(loc_gp[k], nuv, nsc))
    tr.synchronize(); torch.cuda.synchronize()
    ok_dir = all(bool(torch.equal(a, b[si])) for k in range(3) for a, b in zip(souts[k], ref_sp[k]))
    res["stress_inv3_bit_identical"] = ok_inv
    res["stress_dir3_bit_identical"] = ok_dir
    # chunked host path (numpy arrays in, several inverse sub-calls per call): N ranks vs one rank, same chunking
    h_in = [a.cpu().numpy() for a in ins[0]]
    gh1 = tr1.inv_trans(*h_in)
    ghN = tr.inv_trans(*(np.ascontiguousarray(a[si.cpu().numpy()]) for a in h_in))
    gin = gi.cpu().numpy()
    ok_h = bool(np.array_equal(ghN[0], gh1[0][:, gin]))
    sh1 = tr1.dir_trans(gh1, nuv, nsc)
    shN = tr.dir_trans(np.ascontiguousarray(gh1[:, :, gin]), nuv, nsc)
    sin = si.cpu().numpy()
    ok_h = ok_h and all(np.array_equal(a, b[sin]) for a, b in zip(shN, sh1))
    res["stress_hostpath_bit_identical"] = ok_h
    res["entry_barriers"] = int(tr.comm_info()["entry_barriers"]) if world > 1 else 0
    tr.release(); tr1.release()
    res["ok"] = bool(res["golden_inv_maxabs"] < 1e-10 and res["golden_dir_maxabs"] < 1e-10 and res["t159_bit_identical"]
                     and res["t159_roundtrip_norm_rel"] <= 1e-12 and res["harmonic_roundtrip_norm_err_eps"] <= 100.0
                     and res["harmonic_roundtrip_maxabs"] < 1e-13
                     and ok_inv and ok_dir and ok_h)
    if verbose:
        print(f"[selfcheck rank {rank}] {res}", flush=True)
    return res


def reduce(res, world, dev):
    """Worst case over ranks (max of errors, AND of flags)."""
    if world == 1:
        return res
    import torch
    import torch.distributed as dist
    keys = sorted(res)
    num = torch.tensor([float(res[k]) if not isinstance(res[k], bool) else (0.0 if res[k] else 1.0) for k in keys],
                       dtype=torch.float64, device=dev)
    dist.all_reduce(num, op=dist.ReduceOp.MAX)
    out = {}
    for k, v in zip(keys, num.cpu().tolist()):
        out[k] = (v == 0.0) if isinstance(res[k], bool) else (int(v) if isinstance(res[k], int) else v)
    return out
