"""Run under torchrun: distributed INV_TRANS/DIR_TRANS (m over ranks, latitude bands, NCCL all-to-all)
against the oracle on the same global input."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import torch.distributed as dist
import ectrans_b200 as eb
import ectrans_oracle as eo

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
def fresh_uid():                      # one NCCL id per communicator (= per distributed handle)
    buf = torch.zeros(eb.ECT_NCCL_UID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(eb.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy().tobytes())
uid = fresh_uid()
T, N, nuv, nsc = 79, 80, 3, 4
nloen = eb.octahedral_nloen(N)
GP = os.environ.get("ECT_DIST_GP", "latbands")     # "eq_regions": the reference's grid-point decomposition, TRLTOG / TRGTOL as all-to-alls
tr = eb.Transform(T, nloen, nranks=world, rank=rank, device=local, nccl_uid=uid, gp_partition=GP)
s = eo.setup(T, 2 * N, nloen)
vor = eo.random_spectral(s, nuv, 1, zero00=True); div = eo.random_spectral(s, nuv, 2, zero00=True); sc = eo.random_spectral(s, nsc, 3)
ref = eo.inv_trans(s, vor, div, sc, scders=True)
# local spectral slices in MYMS order
idx = np.concatenate([np.arange(s.nasm0[m], s.nasm0[m] + 2 * (T - m + 1)) for m in tr.myms]) if tr.nump else np.zeros(0, int)
loc = lambda a: np.ascontiguousarray(a[:, idx].T)
gp = tr.inv_trans(loc(vor), loc(div), loc(sc), scders=True)
gidx = np.concatenate([s.latoff[l] + f + np.arange(c) for l, f, c in tr.gp_segs]) if len(tr.gp_segs) else np.zeros(0, int)
assert gidx.size == tr.ngptot
refloc = ref[:, gidx]                                # this task's grid points (pieces of latitudes)
rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
e_inv = rel(gp[0], refloc)
# NPROMA-blocked arrays (ragged last block) must carry the same numbers
npr = 37
gpb = tr.inv_trans(loc(vor), loc(div), loc(sc), scders=True, nproma=npr)
unb = gpb.transpose(1, 0, 2).reshape(gpb.shape[1], -1)[:, :tr.ngptot]
blocked_same = float(np.array_equal(unb, gp[0]))
ov, od, os_ = tr.dir_trans(np.ascontiguousarray(gp[:, :2 * nuv + nsc]), nuv, nsc)
obv, obd, obs = tr.dir_trans(np.ascontiguousarray(gpb[:, :2 * nuv + nsc]), nuv, nsc, nproma=npr)
blocked_same = min(blocked_same, float(np.array_equal(obv, ov) and np.array_equal(obd, od) and np.array_equal(obs, os_)))
rv, rd, rs = eo.dir_trans(s, ref[:2 * nuv + nsc], nuv, nsc)
e_dir = max(rel(ov.T, rv[:, idx]), rel(od.T, rd[:, idx]), rel(os_.T, rs[:, idx])) if tr.nump else 0.0
nrm = tr.specnorm(loc(sc))
e_nrm = rel(nrm, eo.specnorm(s, sc))
tim = tr.timings()
# ---- GATH_GRID / GATH_SPEC / DIST_* and reproducibility across decompositions (the property behind the reference's
# --dump-checksums harness, ectrans-benchmark.F90:1455-1638): the gathered results must equal, bit for bit, those of
# one rank transforming the same global input ----
nfg = gp.shape[1]
kto = np.arange(nfg, dtype=np.int32) % world                       # field f gathered on rank f % world
gg = tr.gath_grid(gp, kto=kto)
gv, gd, gs = tr.gath_spec(ov, kto=0), tr.gath_spec(od, kto=0), tr.gath_spec(os_, kto=np.arange(nsc) % world)
T_ = lambda a: np.ascontiguousarray(a.T)
dv = tr.dist_spec(T_(vor) if rank == 0 else None, nuv, kfrom=0)
e_dist = float(np.abs(dv - loc(vor)).max()) if tr.nump else 0.0
dg = tr.dist_grid(ref[:nfg] if rank == 0 else None, nfg, kfrom=0)
e_dist = max(e_dist, float(np.abs(dg[0] - refloc[:nfg]).max()))
bit = blocked_same
tr1 = eb.Transform(T, nloen, device=local)                          # the same transform on one rank
g1 = tr1.inv_trans(T_(vor), T_(div), T_(sc), scders=True)
mine = [f for f in range(nfg) if kto[f] == rank]
bit = min(bit, float(np.array_equal(gg, g1[0][mine])))
e_gath = rel(gg, ref[mine]) if mine else 0.0
v1, d1, s1 = tr1.dir_trans(np.ascontiguousarray(g1[:, :2 * nuv + nsc]), nuv, nsc)
def z(a):                                                            # GATH_SPEC zeroes Im(m = 0)
    b = a.copy(); b[1:2 * (T + 1):2] = 0
    return b
if rank == 0:
    bit = min(bit, float(np.array_equal(gv, z(v1))), float(np.array_equal(gd, z(d1))))
smine = [f for f in range(nsc) if f % world == rank]
bit = min(bit, float(np.array_equal(gs, z(s1)[:, smine])))
# ---- GPNORM_TRANS (bit identical across decompositions: latitudes are added in global order), VORDIV_TO_UV on the
# task's wavenumbers, Legendre cache file per task ----
bit = min(bit, float(np.array_equal(tr.specnorm(os_), tr1.specnorm(s1))))        # SPECNORM: m sums added in order
a_n, lo_n, hi_n = tr.gpnorm_trans(gp)
a_1, lo_1, hi_1 = tr1.gpnorm_trans(g1)
bit = min(bit, float(np.array_equal(a_n, a_1) and np.array_equal(lo_n, lo_1) and np.array_equal(hi_n, hi_1)))
un, vn = tr.vordiv_to_uv(loc(vor), loc(div))
u1, v1_ = tr1.vordiv_to_uv(T_(vor), T_(div))
if tr.nump:
    bit = min(bit, float(np.array_equal(un, u1[idx]) and np.array_equal(vn, v1_[idx])))
import tempfile
path = os.path.join(tempfile.gettempdir(), "legpol_%d_of_%d.bin" % (rank, world))
tr.release()
trw = eb.Transform(T, nloen, nranks=world, rank=rank, device=local, nccl_uid=fresh_uid(), legpol_write=path, gp_partition=GP)
trw.release()
tr = eb.Transform(T, nloen, nranks=world, rank=rank, device=local, nccl_uid=fresh_uid(), legpol_read=path, gp_partition=GP)
gp2 = tr.inv_trans(loc(vor), loc(div), loc(sc), scders=True)
bit = min(bit, float(np.array_equal(gp2, gp)))
os.remove(path)
tr1.release()
t = torch.tensor([e_inv, e_dir, e_nrm, e_dist, e_gath, 1.0 - bit], device=dev, dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("dist errors inv %.2e dir %.2e norm %.2e dist %.2e gath %.2e not-bit-identical %g" % tuple(t.cpu().tolist()), tim)
    ok = bool((t[:3] < 1e-12).all()) and float(t[3]) == 0.0 and float(t[4]) < 1e-12 and float(t[5]) == 0.0
    print("DIST_CHECK_OK" if ok else "DIST_CHECK_FAIL")
tr.release()
dist.destroy_process_group()
