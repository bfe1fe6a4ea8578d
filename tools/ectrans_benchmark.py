#!/usr/bin/env python
"""ectrans-benchmark on the B200 backend: same flags, same synthetic input, same pass criterion and the same
timing blocks as the reference driver (src/programs/ectrans-benchmark.F90: options :1141-1186, field layout
:450-527, input Re psi(4,19) = 1 :1389-1415, time-step loop :619-724, error norms :790-871, stats :878-943).

  python tools/ectrans_benchmark.py -t 79 -g O80 -l 10 -f 1 -n 10 --niter-warmup 3 --norms --check 100
  python tools/ectrans_benchmark.py -t 1279 -g O1280 -l 137 -f 1 --device-resident

Host arrays by default (the timed INV_TRANS / DIR_TRANS calls include H2D / D2H, as with the reference GPU
backend; pinned unless --no-pinning); --device-resident keeps the fields in HBM.
Single task (the reference's LDMPOFF mode); multi-GPU goes through bench.py / torchrun.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ectrans_b200 as eb  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-t", "--truncation", type=int, default=79)
    ap.add_argument("-g", "--grid", default=None, help="O<N> (octahedral) or F<N> (full); default O<T+1>")
    ap.add_argument("-l", "--nlev", type=int, default=1)
    ap.add_argument("-f", "--nfld", type=int, default=1)
    ap.add_argument("-n", "--niter", type=int, default=10)
    ap.add_argument("--niter-warmup", type=int, default=3)
    ap.add_argument("--nproma", type=int, default=0)
    ap.add_argument("--vordiv", action="store_true")
    ap.add_argument("--scders", action="store_true")
    ap.add_argument("--uvders", action="store_true")
    ap.add_argument("--norms", action="store_true")
    ap.add_argument("--check", type=int, default=0, metavar="NCHECK")
    ap.add_argument("--dump-checksums", default=None, metavar="FILE",
                    help="running 64-bit checksum of every gathered field after each direct transform, in the layout of "
                         "the reference's dump_checksums (ectrans-benchmark.F90:1455-1638; the reference's crc64 comes from "
                         "fiat, which is not available here: the value is zlib crc32 | adler32 << 32, good for comparing "
                         "runs of this backend across decompositions, not against reference dumps)")
    ap.add_argument("--no-pinning", action="store_true")
    ap.add_argument("--device-resident", action="store_true")
    ap.add_argument("--precision", default="dp", choices=["dp", "sp"])
    ap.add_argument("-v", action="count", default=0)
    a = ap.parse_args()

    T = a.truncation
    grid = a.grid or f"O{T + 1}"
    n = int(grid[1:])
    nloen = eb.octahedral_nloen(n) if grid[0].upper() == "O" else np.full(2 * n, 4 * n, dtype=np.int32)
    t0 = time.perf_counter()
    tr = eb.Transform(T, nloen, precision=a.precision)
    print("======= Start of runtime parameters =======\n")
    print(f"nsmax      {T}\ngrid       {grid}\nndgl       {2 * n}\nnlev       {a.nlev}\nnflds      {a.nfld}")
    print(f"ngptot     {tr.ngptot}\nnspec2     {tr.nspec2}\nprecision  {a.precision}\nsetup (s)  {time.perf_counter() - t0:8.4f}")
    print("\n======= End of runtime parameters =======\n")
    nuv, nsc = a.nlev, a.nlev * a.nfld + 1
    dt = tr.dtype
    ia = int(tr.nasm0[4]) + 2 * (19 - 4) if T >= 19 else None

    def alloc(shape):
        if a.device_resident:
            import torch
            return torch.zeros(shape, dtype=tr._tdtype(), device="cuda")
        if a.no_pinning:
            return np.zeros(shape, dtype=dt)
        return eb.PinnedArray(shape, dtype=dt).array

    spvor, spdiv, spsc = alloc((tr.nspec2, nuv)), alloc((tr.nspec2, nuv)), alloc((tr.nspec2, nsc))
    for arr in (spvor, spdiv, spsc):
        arr[...] = 0
        if ia is not None:
            arr[ia, :] = 1.0
    nfld_gp = tr.gp_fields(nuv, nsc, a.scders, a.vordiv, a.vordiv, a.uvders)
    nproma, nblk = tr._blocks(a.nproma)
    gp = alloc((nblk, nfld_gp, nproma))
    iu = 2 * nuv if a.vordiv else 0
    opts = dict(scders=a.scders, vorgp=a.vordiv, divgp=a.vordiv, uvder=a.uvders, nproma=a.nproma)
    n0 = [tr.specnorm(x) for x in (spvor, spdiv, spsc)]

    def sync():
        if a.device_resident:
            tr.synchronize()

    def dump_checksums(jstep):
        import zlib
        host = lambda x: x.cpu().numpy() if a.device_resident else np.asarray(x)
        lines = ["====================", f"iteration {jstep}", "===================="]
        for name, arr, gath in (("zgp", host(gp), tr.gath_grid), ("zspvor", host(spvor), tr.gath_spec),
                                ("zspdiv", host(spdiv), tr.gath_spec), ("zspscalar", host(spsc), tr.gath_spec)):
            icrc = 0
            nf = arr.shape[1]
            for jf in range(nf):
                g = gath(np.ascontiguousarray(arr[:, jf:jf + 1]), nproma=a.nproma) if name == "zgp" else gath(np.ascontiguousarray(arr[:, jf:jf + 1]))
                buf = np.ascontiguousarray(g).tobytes()
                icrc = zlib.crc32(buf, icrc & 0xffffffff) | (zlib.adler32(buf, (icrc >> 32) or 1) << 32)
                lines.append(f"{name} ({jf + 1}) = {icrc:016X}")
        with open(a.dump_checksums, "a" if jstep > 1 else "w") as fh:
            fh.write("\n".join(lines) + "\n")

    print("======= Start of spectral transforms  =======\n")
    print(f"Running for {a.niter} iterations with {a.niter_warmup} extra warm-up iterations\n")
    t_inv, t_dir, t_step = [], [], []
    gin = None
    tloop0 = time.perf_counter()
    for jstep in range(1, a.niter + a.niter_warmup + 1):
        sync(); t1 = time.perf_counter()
        tr.inv_trans(spvor, spdiv, spsc, out=gp, **opts)
        sync(); t2 = time.perf_counter()
        src = gp[:, iu:iu + 2 * nuv + nsc]
        if gin is None or iu or nfld_gp != 2 * nuv + nsc:
            if a.device_resident:
                gin = src.contiguous()
            else:
                if gin is None:
                    gin = alloc((nblk, 2 * nuv + nsc, nproma))
                gin[...] = src
        else:
            gin = gp
        sync(); t3 = time.perf_counter()
        tr.dir_trans(gin, nuv, nsc, nproma=a.nproma, out=(spvor, spdiv, spsc))
        sync(); t4 = time.perf_counter()
        if jstep > a.niter_warmup:
            t_inv.append(t2 - t1); t_dir.append(t4 - t3); t_step.append((t2 - t1) + (t4 - t3))
        line = f"time step {jstep:6d} took{(t2 - t1) + (t4 - t3):8.4f}"
        if a.dump_checksums:
            dump_checksums(jstep)
        if a.norms:
            errs = [float(np.abs(tr.specnorm(x) / n - 1).max()) for x, n in zip((spvor, spdiv, spsc), n0)]
            line += f" | zspvor max err={errs[0]:10.3e} | zspdiv max err={errs[1]:10.3e} | zspscalar max err={errs[2]:10.3e}"
        print(line)
    tloop = time.perf_counter() - tloop0
    print("\n======= End of spectral transforms  =======\n")
    rc = 0
    if a.norms or a.check:
        errs = [float(np.abs(tr.specnorm(x) / n - 1).max()) for x, n in zip((spvor, spdiv, spsc), n0)]
        print(f"max error zspvor(1:nlev,:)    = {errs[0]:10.3e}")
        print(f"max error zspdiv(1:nlev,:)    = {errs[1]:10.3e}")
        print(f"max error zspscalar(1:nlev,:,1) = {errs[2]:10.3e}\n")
        print(f"max error combined =          = {max(errs):10.3e}\n")
        if a.check:
            tol = a.check * float(np.finfo(dt).eps)
            if max(errs) > tol:
                print("*******************************\nCorrectness test failed")
                print(f"Maximum spectral norm error = {max(errs):9.2e}\nError tolerance = {tol:9.2e}\n*******************************")
                rc = 1
    print("======= Start of time step stats =======\n")
    for title, v in (("Inverse transforms", t_inv), ("Direct transforms", t_dir), ("Inverse-direct transforms", t_step)):
        v = np.array(v)
        print(title); print("-" * len(title))
        print(f"avg  (s): {v.mean():8.4f}\nmin  (s): {v.min():8.4f}\nmax  (s): {v.max():8.4f}\nmed  (s): {np.median(v):8.4f}")
        if title.startswith("Inverse-direct"):
            print(f"loop (s): {tloop:8.4f}")
        print(" ")
    print("======= End of time step stats =======\n")
    tr.release()
    return rc


if __name__ == "__main__":
    sys.exit(main())
