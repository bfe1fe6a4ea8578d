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
# ---- call mode 2, the benchmark's shape (ectrans-benchmark.F90:473-527): PSPVOR/PSPDIV(lev_l, nspec2), PSPSC3A(lev_l, nspec2,
# nfld), PSPSC2(1, nspec2) -> PGPUV(nproma, lev, 2 + 2 ders, blk), PGP3A(nproma, lev, 3 nfld, blk), PGP2(nproma, 3, blk); device
# pointers for the inverse, host arrays for the direct transform ----
nlev, nfld3 = nuv, 2
sc3 = eo.random_spectral(s, nlev * nfld3, 7).reshape(nfld3, nlev, -1)          # [fld][lev][nspec2]
sc2 = eo.random_spectral(s, 1, 8)
kv3a = kvuv.copy(); kv2 = np.array([1 % V + 1])
l3 = np.where(kv3a == v + 1)[0]; l2 = np.where(kv2 == v + 1)[0]
ref2 = eo.inv_trans(s, vor, div, np.concatenate([sc2, sc3.reshape(nfld3 * nlev, -1)]), scders=True, uvder=True)
npr = tr.ngptot
dv_ = torch.from_numpy(loc(vor, luv)).to(dev); dd_ = torch.from_numpy(loc(div, luv)).to(dev)
d3a = torch.from_numpy(np.ascontiguousarray(sc3[:, l3][:, :, idx].transpose(0, 2, 1))).to(dev)      # (fld, nspec2, lev_l) = PSPSC3A(lev_l, nspec2, fld)
d2 = torch.from_numpy(loc(sc2, l2)).to(dev)
guv = torch.zeros((1, 4, nlev, npr), dtype=torch.float64, device=dev)             # PGPUV(nproma, lev, 4, 1): u v du dv
g3a = torch.zeros((1, 3 * nfld3, nlev, npr), dtype=torch.float64, device=dev)
g2 = torch.zeros((1, 3, npr), dtype=torch.float64, device=dev)
vsets = dict(kvsetuv=kvuv, kvsetsc2=kv2, kvsetsc3a=kv3a)
kw = dict(memspace=eb.ECT_MEM_DEVICE, scders=1, uvder=1, nuv=len(luv), gpuv=guv, gp2=g2, gp3a=g3a, nsc3a_fld=nfld3, nsc3a_lev=len(l3), nsc2=len(l2))
if len(luv): kw.update(spvor=dv_, spdiv=dd_)
if len(l3): kw.update(spsc3a=d3a)
if len(l2): kw.update(spsc2=d2)
tr.inv_trans_vset_raw(vsets, **kw)
torch.cuda.synchronize()
# oracle field order: u v | sc2, 3a(fld, lev) | nsd of those | du dv | ewd of those
nscg = 1 + nfld3 * nlev
R = ref2[:, gidx]
e_m2 = max(rel(guv[0, 0].cpu().numpy(), R[0:nlev]), rel(guv[0, 1].cpu().numpy(), R[nlev:2 * nlev]),
           rel(guv[0, 2].cpu().numpy(), R[2 * nlev + 2 * nscg:3 * nlev + 2 * nscg]), rel(guv[0, 3].cpu().numpy(), R[3 * nlev + 2 * nscg:4 * nlev + 2 * nscg]))
o_sc, o_ns, o_ew = 2 * nlev, 2 * nlev + nscg, 4 * nlev + 2 * nscg
for part, o in enumerate((o_sc, o_ns, o_ew)):
    e_m2 = max(e_m2, rel(g2[0, part].cpu().numpy(), R[o]))
    for j3 in range(nfld3):
        e_m2 = max(e_m2, rel(g3a[0, part * nfld3 + j3].cpu().numpy(), R[o + 1 + j3 * nlev:o + 1 + (j3 + 1) * nlev]))
# direct, host arrays: PGPUV(nproma, lev, 2, 1), PGP3A(nproma, lev, nfld, 1), PGP2(nproma, 1, 1)
huv = np.ascontiguousarray(np.stack([R[0:nlev], R[nlev:2 * nlev]])[None])
h3a = np.ascontiguousarray(R[o_sc + 1:o_sc + 1 + nfld3 * nlev].reshape(nfld3, nlev, -1)[None])
h2 = np.ascontiguousarray(R[o_sc:o_sc + 1][None])
o_v = np.zeros((tr.nspec2, len(luv))); o_d = np.zeros_like(o_v)
o_3a = np.zeros((nfld3, tr.nspec2, len(l3))); o_2 = np.zeros((tr.nspec2, len(l2)))
kw = dict(memspace=eb.ECT_MEM_HOST, nuv=len(luv), gpuv=huv, gp2=h2, gp3a=h3a, nsc3a_fld=nfld3, nsc3a_lev=len(l3), nsc2=len(l2))
if len(luv): kw.update(spvor=o_v, spdiv=o_d)
if len(l3): kw.update(spsc3a=o_3a)
if len(l2): kw.update(spsc2=o_2)
tr.dir_trans_vset_raw(vsets, **kw)
rv2, rd2, rs2 = eo.dir_trans(s, np.concatenate([ref2[0:2 * nlev], ref2[o_sc:o_sc + nscg]]), nlev, nscg)
if tr.nump:
    if len(luv): e_m2 = max(e_m2, rel(o_v.T, rv2[luv][:, idx]), rel(o_d.T, rd2[luv][:, idx]))
    if len(l2): e_m2 = max(e_m2, rel(o_2.T, rs2[0:1][:, idx]))
    for j3 in range(nfld3):
        if len(l3): e_m2 = max(e_m2, rel(o_3a[j3].T, rs2[1 + j3 * nlev:1 + (j3 + 1) * nlev][l3][:, idx]))
e_dir = max(e_dir, e_m2)
# SPECNORM with KVSET: the norms of all global fields on every task, bit identical with one rank (m sums added in order)
nrm = tr.specnorm_vset(loc(sc, lsc), kvsc)
e_nrm = rel(nrm, eo.specnorm(s, sc))
e_dir = max(e_dir, e_nrm * 1e-1)             # 1e-13 bound folded into the 1e-12 one
# one rank, same input
tr1 = eb.Transform(T, nloen, device=local)
T_ = lambda a: np.ascontiguousarray(a.T)
g1 = tr1.inv_trans(T_(vor), T_(div), T_(sc), **opts)
e_one = rel(gp[0], g1[0][:, gidx])
v1, d1, s1 = tr1.dir_trans(np.ascontiguousarray(g1[:, nuv:nuv + 2 * nuv + nsc]), nuv, nsc)
if tr.nump:
    e_one = max(e_one, rel(ov, v1[idx][:, luv]), rel(od, d1[idx][:, luv]), rel(os_, s1[idx][:, lsc]))
same = min(same, float(np.array_equal(nrm, tr1.specnorm(T_(sc)))))
tr1.release()
t = torch.tensor([e_inv, e_dir, e_one, 1.0 - same], device=dev, dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("vset check (W=%d V=%d): inv %.2e dir %.2e vs-one-rank %.2e blocked-differs %g" % ((world // V, V) + tuple(t.cpu().tolist())))
    print("VSET_CHECK_OK" if float(t[0]) < 1e-12 and float(t[1]) < 1e-12 and float(t[2]) < 1e-13 and float(t[3]) == 0.0 else "VSET_CHECK_FAIL")
tr.release()
dist.destroy_process_group()
