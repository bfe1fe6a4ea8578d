"""World-size-2 (and 3) check of the transposition plan on CPU with the gloo backend: each rank builds
its host plan through the C ABI, fills the Legendre-side Fourier buffer with (m, latitude) tags using
its record tables, exchanges with all_to_all_single exactly as ect_transpose() does with NCCL
(TRMTOL, reference cpu/internal/trmtol_mod.F90:101-141), and verifies through the FFT-side table that
every record landed where the Fourier stage will look for it; then the reverse (TRLTOM)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, T, N, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import ectrans_b200 as eb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        t = eb.Transform(T, eb.octahedral_nloen(N), nranks=world, rank=rank, host_only=True)
        rt = t.record_tables()
        nleg, nfft = int(t.send_cnt.sum()), int(t.recv_cnt.sum())
        ndgl, ndgnh = t.ndgl, t.ndgl // 2
        leg = np.full((nleg, 2), -1, dtype=np.int64)
        for ml, m in enumerate(t.myms):
            nd = int(t.ndglu[m]); isl = ndgnh - nd
            for i in range(nd):
                leg[rt["leg_rec_n"][rt["mrow0"][ml] + i]] = (m, isl + i)
                leg[rt["leg_rec_s"][rt["mrow0"][ml] + i]] = (m, ndgl - 1 - (isl + i))
        assert (leg >= 0).all()
        fft = torch.full((nfft, 2), -7, dtype=torch.int64)
        dist.all_to_all_single(fft, torch.from_numpy(leg), output_split_sizes=[int(c) for c in t.recv_cnt],
                               input_split_sizes=[int(c) for c in t.send_cnt])
        fft = fft.numpy()
        lat0 = int(t.info.lat0)
        for l in range(int(t.info.nlat)):
            g = lat0 + l
            for m in range(int(t.nmen[g]) + 1):
                r = rt["fft_rec"][rt["latrow0"][l] + m]
                assert tuple(fft[r]) == (m, g), (rank, l, m, fft[r])
        # reverse direction (TRLTOM)
        back = torch.full((nleg, 2), -9, dtype=torch.int64)
        dist.all_to_all_single(back, torch.from_numpy(fft), output_split_sizes=[int(c) for c in t.send_cnt],
                               input_split_sizes=[int(c) for c in t.recv_cnt])
        assert np.array_equal(back.numpy(), leg)
        q.put((rank, "ok", nleg, nfft))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e), 0, 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,T,N", [(2, 63, 64), (3, 47, 48)])
def test_transposition_plan_gloo(built, world, T, N):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, T, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
    assert sum(r[2] for r in res) == sum(r[3] for r in res)


def _worker_gp(rank, world, V, port, T, N, q):
    """TRLTOG / TRGTOL of the eq_regions partition and of V-sets, emulated with gloo send/recv: the message layout
    ([fields of the source V-set][points of the pair], caller fields in V-set-major message order) and the offsets are the
    ones gp_exchange() in csrc/api.cu uses with NCCL."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import ectrans_b200 as eb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nloen = eb.octahedral_nloen(N)
        kw = dict(nprtrv=V) if V > 1 else dict(gp_partition="eq_regions")
        t = eb.Transform(T, nloen, nranks=world, rank=rank, host_only=True, **kw)
        W, w, v = world // V, rank // V, rank % V
        latoff = np.concatenate([[0], np.cumsum(nloen)])
        band0 = int(latoff[t.info.lat0]); nband = int(latoff[t.info.lat0 + t.info.nlat]) - band0
        xb_idx = t._arr(30, np.int32, nband); xb_off = t._arr(31, np.int64, world + 1); xg_off = t._arr(32, np.int64, W + 1)
        nfg = 5                                              # global fields, V-set of field f = f % V
        fields = [[f for f in range(nfg) if f % V == vv] for vv in range(V)]
        nfl = [len(x) for x in fields]
        mine = fields[v]
        # band buffer B[local field][band point] = tag(global field, global point)
        tag = lambda f, g: f * 10_000_000 + g
        B = np.array([[tag(f, band0 + i) for i in range(nband)] for f in mine], dtype=np.int64).reshape(len(mine), nband)
        # pack (k_gp_band_msg): message to task p = [local field][points xb_idx[xb_off[p]:xb_off[p+1]]]
        sends = [np.ascontiguousarray(B[:, xb_idx[xb_off[p]:xb_off[p + 1]]]) for p in range(world)]
        ngp = t.ngptot
        recv_buf = np.full(ngp * nfg, -1, dtype=np.int64)   # [block of owner w'][field slot][points of w']
        reqs = []
        for p in range(world):
            if sends[p].size and p != rank:                 # gloo has no self pair: the own message is copied below
                reqs.append(dist.isend(torch.from_numpy(sends[p].reshape(-1)), p))
        views = []
        for wp in range(W):
            cnt = int(xg_off[wp + 1] - xg_off[wp]); pre = 0
            for vp in range(V):
                n = cnt * nfl[vp]
                at = int(xg_off[wp]) * nfg + cnt * pre
                pre += nfl[vp]
                if n and wp * V + vp == rank:
                    views.append((at, torch.from_numpy(sends[rank].reshape(-1).copy())))
                elif n:
                    buf = torch.empty(n, dtype=torch.int64)
                    reqs.append(dist.irecv(buf, wp * V + vp)); views.append((at, buf))
        for r in reqs:
            r.wait()
        for at, buf in views:
            recv_buf[at:at + buf.numel()] = buf.numpy()
        # unpack (k_gp_user_msg): field slot s of the message order <-> global field order[s]
        order = [f for vv in range(V) for f in fields[vv]]
        gidx = np.concatenate([latoff[l] + f0 + np.arange(c) for l, f0, c in t.gp_segs]) if len(t.gp_segs) else np.zeros(0, int)
        user = np.full((nfg, ngp), -1, dtype=np.int64)
        for g in range(ngp):
            wp = int(np.searchsorted(xg_off, g, side="right") - 1)
            o, cnt = int(xg_off[wp]), int(xg_off[wp + 1] - xg_off[wp])
            for s_, f in enumerate(order):
                user[f, g] = recv_buf[o * nfg + s_ * cnt + (g - o)]
        want = np.array([[tag(f, int(gi)) for gi in gidx] for f in range(nfg)], dtype=np.int64).reshape(nfg, ngp)
        assert np.array_equal(user, want)
        q.put((rank, "ok", ngp, nband))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e), 0, 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,V", [(2, 1), (3, 1), (2, 2), (4, 2)])
def test_gridpoint_exchange_gloo(built, world, V):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 1000) + 10 * world + V
    procs = [ctx.Process(target=_worker_gp, args=(r, world, V, port, 31, 32, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
    assert sum(r[2] for r in res) == 2 * sum(20 + 4 * i for i in range(32))          # every grid point of O32 on exactly one task
