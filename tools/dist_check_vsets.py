"""Run under torchrun: INV_TRANS / DIR_TRANS with NPRTRV > 1 (fields spread over V-sets, eq_regions grid-point tasks,
TRLTOG / TRGTOL redistributing points and fields) against the oracle and against one rank.  Across NPRTRV the results
agree to rounding, not bit for bit: two real fields share one complex FFT and the V-sets change which fields pair up
(across NPRTRW, where the pairs stay the same, tools/dist_check.py demands bit identity).
ECT_DIST_V = NPRTRV (default 2); world must be a multiple of it."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import torch.distributed as dist
import ectrans_b200 as eb
import ectrans_oracle as eo

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
V = int(os.environ.get("ECT_DIST_V", "2"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
buf = torch.zeros(eb.ECT_NCCL_UID_BYTES, dtype=torch.uint8, device=dev)
if rank == 0:
    buf.copy_(torch.frombuffer(bytearray(eb.nccl_unique_id()), dtype=torch.uint8))
dist.broadcast(buf, 0)
uid = bytes(buf.cpu().numpy().tobytes())
T, N, nuv, nsc = 63, 64, 5, 7
nloen = eb.octahedral_nloen(N)
tr = eb.Transform(T, nloen, nranks=world, rank=rank, device=local, nccl_uid=uid, nprtrv=V)
v = rank % V
s = eo.setup(T, 2 * N, nloen)
vor = eo.random_spectral(s, nuv, 1, zero00=True); div = eo.random_spectral(s, nuv, 2, zero00=True); sc = eo.random_spectral(s, nsc, 3)
kvuv = np.arange(nuv) % V + 1                    # the benchmark's round-robin V-sets (ectrans-benchmark.F90:480-507)
kvsc = (np.arange(nsc) + 1) % V + 1
opts = dict(scders=True, uvder=True, vorgp=True)
ref = eo.inv_trans(s, vor, div, sc, scders=True, uvder=True, vorgp=True, divgp=False)
idx = np.concatenate([np.arange(s.nasm0[m], s.nasm0[m] + 2 * (T - m + 1)) for m in tr.myms]) if tr.nump else np.zeros(0, int)
luv, lsc = np.where(kvuv == v + 1)[0], np.where(kvsc == v + 1)[0]
loc = lambda a, sel: np.ascontiguousarray(a[sel][:, idx].T)
gidx = np.concatenate([s.latoff[l] + f + np.arange(c) for l, f, c in tr.gp_segs]) if len(tr.gp_segs) else np.zeros(0, int)
rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
gp = tr.inv_trans_vset(loc(vor, luv), loc(div, luv), loc(sc, lsc), kvuv, kvsc, **opts)
e_inv = max(rel(gp[0][i], ref[i][gidx]) for i in range(ref.shape[0]))
# blocked arrays carry the same numbers
gpb = tr.inv_trans_vset(loc(vor, luv), loc(div, luv), loc(sc, lsc), kvuv, kvsc, nproma=41, **opts)
same = float(np.array_equal(gpb.transpose(1, 0, 2).reshape(gpb.shape[1], -1)[:, :tr.ngptot], gp[0]))
gin = np.ascontiguousarray(gp[:, nuv:nuv + 2 * nuv + nsc])          # u, v, scalars (vorticity precedes them)
ov, od, os_ = tr.dir_trans_vset(gin, kvuv, kvsc)
rv, rd, rs = eo.dir_trans(s, ref[nuv:nuv + 2 * nuv + nsc], nuv, nsc)
e_dir = 0.0
if tr.nump:
    if len(luv): e_dir = max(e_dir, rel(ov.T, rv[luv][:, idx]), rel(od.T, rd[luv][:, idx]))
    if len(lsc): e_dir = max(e_dir, rel(os_.T, rs[lsc][:, idx]))
# one rank, same input: bit identity
tr1 = eb.Transform(T, nloen, device=local)
T_ = lambda a: np.ascontiguousarray(a.T)
g1 = tr1.inv_trans(T_(vor), T_(div), T_(sc), **opts)
e_one = rel(gp[0], g1[0][:, gidx])
v1, d1, s1 = tr1.dir_trans(np.ascontiguousarray(g1[:, nuv:nuv + 2 * nuv + nsc]), nuv, nsc)
if tr.nump:
    e_one = max(e_one, rel(ov, v1[idx][:, luv]), rel(od, d1[idx][:, luv]), rel(os_, s1[idx][:, lsc]))
tr1.release()
t = torch.tensor([e_inv, e_dir, e_one, 1.0 - same], device=dev, dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("vset check (W=%d V=%d): inv %.2e dir %.2e vs-one-rank %.2e blocked-differs %g" % ((world // V, V) + tuple(t.cpu().tolist())))
    print("VSET_CHECK_OK" if float(t[0]) < 1e-12 and float(t[1]) < 1e-12 and float(t[2]) < 1e-13 and float(t[3]) == 0.0 else "VSET_CHECK_FAIL")
tr.release()
dist.destroy_process_group()
