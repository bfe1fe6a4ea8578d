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
