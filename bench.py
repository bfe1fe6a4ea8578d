#!/usr/bin/env python
"""bench.py -- ms per INV_TRANS + DIR_TRANS step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our CUDA path
  python bench.py --impl reference ...                     the reference algorithm on the host cores
                                                           (oracle port: NumPy/OpenBLAS + pocketfft)

One "step" = one inverse + one direct transform of the whole field set
(vor/div on nlev levels + nlev*nfld + 1 scalars), like one iteration of
ectrans-benchmark (src/programs/ectrans-benchmark.F90:619-724).

N = 1 workload: TCo1279 / O1280, 137 levels, dp -- the configuration the metric is quoted on
(it fits one 180 GB B200).  For N > 1 the same global problem is split over the ranks
(m over ranks for the Legendre stage, latitude bands for the Fourier stage, NCCL all-to-all
between them): strong scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (T, N of O-grid, nlev, nfld)
    "T79_O80_L10": (79, 80, 10, 1),
    "T159_O160_L137": (159, 160, 137, 3),
    "TCo399_O400_L137": (399, 400, 137, 1),
    "TCo1279_O1280_L137": (1279, 1280, 137, 1),
    "TCo2559_O2560_L137": (2559, 2560, 137, 1),
}
# BASELINE.json configs 2 and 4 are single precision builds (JPRB = real32)
PRECISION = {"TCo399_O400_L137": "sp", "TCo2559_O2560_L137": "sp"}


def legendre_flops(T, ndglu, nfields):
    """sum_m 2 c(m) NDGLU(m) (floor((T-m+2)/2) + floor((T-m+3)/2)), per direction."""
    tot = 0
    for m in range(T + 1):
        c = nfields if m == 0 else 2 * nfields
        tot += 2 * c * int(ndglu[m]) * ((T - m + 2) // 2 + (T - m + 3) // 2)
    return float(tot)


def fft_bytes(nloen, nfields, size=8):
    return float(2 * sum(2 * (int(n) // 2 + 1) for n in nloen) * nfields * size)


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.maxmhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.maxmhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.maxmhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.maxmhz, "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------------------------
#  CPU arm: the oracle port on a bounded sample of the workload
# ---------------------------------------------------------------------------------------------
def cpu_sample_step(T, N, nuv, nsc, stride, rng_seed=0):
    """Times the reference algorithm (oracle port) on every `stride`-th zonal wavenumber (Legendre
    stage incl. prologue/epilogue arithmetic) and every `stride`-th latitude (Fourier stage) of the
    workload, all fields, inverse + direct, and scales each part by the sampled share of the Legendre
    flops / grid points.  Returns (estimated ms per full step, detail)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ectrans_oracle as eo
    import scipy.fft as sfft
    cores = use_all_host_threads()
    nloen = eo.octahedral_nloen(N)
    ms = list(range(0, T + 1, stride))
    key = (T, N, stride)
    if key not in _CPU_SETUP:
        _CPU_SETUP[key] = eo.setup(T, 2 * N, nloen, ms=ms)
    s = _CPU_SETUP[key]
    rng = np.random.default_rng(rng_seed)
    nf = 2 * nuv + nsc
    # spectral input only for the sampled m (same layout, other coefficients are never touched)
    mk = lambda n: rng.uniform(-0.1, 0.1, size=(n, s.nspec2)) if n else None
    vor, div, sc = mk(nuv), mk(nuv), mk(nsc)
    t_leg_inv = t_leg_dir = 0.0
    fl_s = 0.0
    for m in ms:
        ndglu = int(s.ndglu[m])
        if ndglu == 0:
            continue
        t0 = time.perf_counter()
        north, south = eo.ltinv_m(s, m, vor, div, sc)
        t1 = time.perf_counter()
        _ = eo.ledir_m(s, m, north, south, nuv)
        t2 = time.perf_counter()
        t_leg_inv += t1 - t0
        t_leg_dir += t2 - t1
        c = nf if m == 0 else 2 * nf
        fl_s += 2 * c * ndglu * ((T - m + 2) // 2 + (T - m + 3) // 2)
    fl = legendre_flops(T, s.ndglu, nf)
    lats = list(range(0, 2 * N, stride))
    t_ft_inv = t_ft_dir = 0.0
    pts = 0.0
    for j in lats:
        nlon = int(nloen[j]); km = int(s.nmen[j])
        half = np.zeros((nf, nlon // 2 + 1), dtype=np.complex128)
        half[:, :km + 1] = rng.standard_normal((nf, km + 1)) + 1j * rng.standard_normal((nf, km + 1))
        t0 = time.perf_counter()
        row = sfft.irfft(half, n=nlon, axis=1, workers=cores) * nlon
        t1 = time.perf_counter()
        _ = (sfft.rfft(row, axis=1, workers=cores) / nlon)[:, :km + 1]
        t2 = time.perf_counter()
        t_ft_inv += t1 - t0
        t_ft_dir += t2 - t1
        pts += nlon
    pts_all = float(nloen.sum())
    leg_inv, leg_dir = t_leg_inv * fl / fl_s, t_leg_dir * fl / fl_s
    ft_inv, ft_dir = t_ft_inv * pts_all / pts, t_ft_dir * pts_all / pts
    total_ms = 1e3 * (leg_inv + leg_dir + ft_inv + ft_dir)
    detail = {"legendre_inv_s": leg_inv, "legendre_dir_s": leg_dir, "fourier_inv_s": ft_inv, "fourier_dir_s": ft_dir,
              "sampled_cpu_s": t_leg_inv + t_leg_dir + t_ft_inv + t_ft_dir}
    return total_ms, detail


_CPU_SETUP = {}
CPU_STRIDE = {79: 1, 159: 1, 399: 4, 1279: 16, 2559: 32}     # sampling of the CPU arm: 1 = the whole workload, nothing extrapolated


def host_cores():
    """Cores this process may use (cgroup / affinity aware) -- NOT the OMP_NUM_THREADS a launcher exported:
    torch.distributed.run sets OMP_NUM_THREADS=1 for its workers, which made the round-1 reference arm 2.5x slower at N > 1."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def use_all_host_threads():
    """BLAS (OpenBLAS inside NumPy) and pocketfft on every core, whatever the environment says."""
    n = host_cores()
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ["OPENBLAS_NUM_THREADS"] = str(n)
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=n)
    except Exception:
        pass
    return n


def cpu_sample_text(stride, nf):
    if stride == 1:
        return (f"the whole workload (every zonal wavenumber and latitude, all {nf} fields), nothing extrapolated; oracle port "
                "(NumPy + OpenBLAS GEMM, scipy pocketfft), all host threads")
    return (f"every {stride}th zonal wavenumber (Legendre stage) and every {stride}th latitude (Fourier stage), all {nf} fields, "
            "EXTRAPOLATED to the whole workload by the sampled share of the Legendre flops / grid points; oracle port "
            "(NumPy + OpenBLAS GEMM, scipy pocketfft), all host threads")


def run_reference(args):
    """The reference algorithm on the host cores.  ectrans-benchmark-cpu itself cannot be built (no Fortran compiler, fiat,
    ecbuild, FFTW in this image -- SURVEY 0.2), so this is the oracle port (kind "port"), with every host thread, on the
    same configuration keys as the GPU arm.  T79 / T159 run the whole workload; larger configurations a stated sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    T, N, nlev, nfld = CONFIGS[args.config]
    nuv, nsc = nlev, nlev * nfld + 1
    nf = 2 * nuv + nsc
    stride = args.cpu_stride or CPU_STRIDE.get(T, 16)
    cores = use_all_host_threads()
    prec = args.precision or PRECISION.get(args.config, "dp")
    vals = []
    t_run = time.time()
    for i in range(args.warmup + args.steps):
        ms, detail = cpu_sample_step(T, N, nuv, nsc, stride, rng_seed=i)
        if i >= args.warmup:
            vals.append(ms)
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": "ms per INV_TRANS+DIR_TRANS step", "value": v, "unit": "ms",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": v,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 (the port computes in double for both precisions)", "data": "synthetic",
            "config": {"workload": args.config, "truncation": T, "grid": f"O{N}", "levels": nlev, "fields": nf,
                       "decomposition": f"host, {cores} threads", "l2": "n/a (host)", "precision": prec},
            "extrapolated": stride != 1, "wall_s": time.time() - t_run,
            "cpu_baseline": {"value": v, "unit": "ms", "cores": cores, "kind": "port", "sample": cpu_sample_text(stride, nf),
                             "detail": detail},
            "e2e": {"value": v, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
#  GPU arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import ectrans_b200 as eb
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    uid = None

    def fresh_uid():                      # one NCCL id per communicator (= per distributed handle)
        import torch.distributed as dist
        buf = torch.zeros(eb.ECT_NCCL_UID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(eb.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        return bytes(buf.cpu().numpy().tobytes())

    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    # ---- parity gate before anything is timed: reference golden vectors through the N-rank transform, N ranks ==
    # one rank bit for bit (T159 with all derivative options), back-to-back same-direction transforms with one rank
    # delayed, chunked host path.  No oracle involved (ectrans_b200/selfcheck.py). ----
    parity = None
    if not args.no_parity:
        from ectrans_b200 import selfcheck
        parity = selfcheck.reduce(selfcheck.run(eb, world, rank, local, fresh_uid), world, dev)
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"metric": "ms per INV_TRANS+DIR_TRANS step", "n_gpus": world, "parity": parity,
                                  "error": "parity check failed: nothing was timed"}), flush=True)
            sys.exit(3)
    if world > 1:
        uid = fresh_uid()
    T, N, nlev, nfld = CONFIGS[args.config]
    nuv, nsc = nlev, nlev * nfld + 1
    nf = 2 * nuv + nsc
    stream = torch.cuda.current_stream().cuda_stream
    t0 = time.time()
    prec = args.precision or PRECISION.get(args.config, "dp")
    tdt, ndt, esz = (torch.float32, np.float32, 4) if prec == "sp" else (torch.float64, np.float64, 8)
    tr = eb.Transform(T, eb.octahedral_nloen(N), nranks=world, rank=rank, device=local, stream=stream, nccl_uid=uid,
                      precision=prec)
    setup_s = time.time() - t0
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    mk = lambda n: (torch.rand((tr.nspec2, n), generator=g, device=dev, dtype=tdt) - 0.5) * 0.2
    spvor, spdiv, spsc = mk(nuv), mk(nuv), mk(nsc)
    gp = torch.empty((1, nf, tr.ngptot), dtype=tdt, device=dev)
    o_vor, o_div, o_sc = (torch.empty_like(spvor), torch.empty_like(spdiv), torch.empty_like(spsc))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        tr.inv_trans(spvor, spdiv, spsc, out=gp)
        ti = tr.timings() if args.stage_timings else None
        tr.dir_trans(gp, nuv, nsc, out=(o_vor, o_div, o_sc))
        td = tr.timings() if args.stage_timings else None
        return ti, td

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1) / args.steps
    # per-stage timings (separate pass: reading them synchronises)
    args.stage_timings = True
    stages = {"inv": [], "dir": []}
    launches = 0
    for _ in range(max(2, min(args.steps, 3))):
        ti, td = step_device()
        stages["inv"].append(ti); stages["dir"].append(td)
        launches = int(ti["launches"] + td["launches"])
    args.stage_timings = False
    med = lambda d, k: float(np.median([x[k] for x in stages[d]]))
    leg_ms = med("inv", "legendre") + med("dir", "legendre")
    ft_ms = med("inv", "fourier") + med("dir", "fourier")
    tp_ms = med("inv", "transpose") + med("dir", "transpose")
    clocks = sampler.summary() if rank == 0 else None
    sampler.stop_flag = True
    t = torch.tensor([ms_dev, leg_ms, ft_ms, tp_ms], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        import torch.distributed as dist
        # per-rank stage times (the transposition time of a rank is mostly the wait for the slowest producer)
        mine = torch.tensor([med("inv", "legendre"), med("dir", "legendre"), med("inv", "fourier"), med("dir", "fourier"),
                             med("inv", "transpose"), med("dir", "transpose")], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {k: [round(float(a[i]), 3) for a in allr] for i, k in enumerate(
            ("legendre_inv", "legendre_dir", "fourier_inv", "fourier_dir", "transpose_inv", "transpose_dir"))}
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, leg_ms, ft_ms, tp_ms = [float(x) for x in t.cpu()]

    # ---- end to end through the host API: pinned host arrays, H2D + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        del gp, o_vor, o_div, o_sc
        torch.cuda.empty_cache()
        h_in = [eb.PinnedArray((tr.nspec2, n), dtype=ndt) for n in (nuv, nuv, nsc)]
        for hp, dv in zip(h_in, (spvor, spdiv, spsc)):
            hp.array[...] = dv.cpu().numpy()
        del spvor, spdiv, spsc
        torch.cuda.empty_cache()
        h_gp = eb.PinnedArray((1, nf, tr.ngptot), dtype=ndt)
        h_out = [eb.PinnedArray((tr.nspec2, n), dtype=ndt) for n in (nuv, nuv, nsc)]

        def step_host():
            tr.inv_trans(h_in[0].array, h_in[1].array, h_in[2].array, out=h_gp.array)
            tr.dir_trans(h_gp.array, nuv, nsc, out=tuple(x.array for x in h_out))
            return float(h_out[2].array[0, 0])

        ne = max(1, args.e2e_steps or args.steps)
        for _ in range(max(1, min(args.warmup, 2))):
            step_host()
        barrier()
        e0.record()
        for _ in range(ne):
            step_host()
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1) / ne
        te = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        bytes_in = esz * (tr.nspec2 * nf + tr.ngptot * nf)
        bytes_out = esz * (tr.ngptot * nf + tr.nspec2 * nf)
        e2e = {"value": float(te.cpu()[0]), "unit": "ms", "h2d_bytes_per_step": int(bytes_in),
               "d2h_bytes_per_step": int(bytes_out), "steps": ne}
    if rank != 0:
        tr.release()
        return
    # ---- roofline of the dominant kernel pair (k_leinv + k_ledir): FP64 tensor (DMMA) ----
    fl = 2.0 * legendre_flops(T, tr.ndglu, nf)       # inverse + direct, whole job
    peak_dmma = eb.measure_fp64_peak(0)
    peak_dfma = eb.measure_fp64_peak(1)
    ach = fl / world / (leg_ms * 1e-3) / 1e12          # per GPU
    hbm_peak, hbm_measured = None, True
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm_peak, hbm_measured = 6650.0, False
    traffic = None
    try:   # DRAM bytes of one k_leinv launch from the committed ncu --set full capture of this workload
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if tj.get("workload") == args.config and world == 1:
            traffic = {"k_leinv_dram_bytes_per_launch": tj["dram_bytes_per_launch"],
                       "k_ledir_dram_bytes_per_launch": tj.get("k_ledir_dram_bytes_per_launch"),
                       "algorithmic_bytes_per_launch": int(tr.info.table_bytes + 8 * (nf * 2) * (sum(T - m + 2 for m in range(T + 1)) + 2 * sum(int(x) for x in tr.ndglu))),
                       "source": tj["source"], "static": True}
    except Exception:
        traffic = None
    fb = 2.0 * fft_bytes(tr.nloen, nf) * esz / 8
    ft_ach = fb / world / (ft_ms * 1e-3) / 1e9
    ft_traffic = None
    try:   # static: DRAM bytes of the largest Fourier launch in each direction, from the committed ncu --set full captures
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if tj.get("workload") == args.config and world == 1 and prec == "dp":
            ft_traffic = dict(tj["fourier"], static=True)
    except Exception:
        ft_traffic = None
    line = {
        "metric": "ms per INV_TRANS+DIR_TRANS step", "value": ms_dev, "unit": "ms", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None,
        "dtype": "f64" if prec == "dp" else "f32 (Fourier stage in float, Legendre contraction 3xTF32 on tcgen05 with fp32 accumulation, m = 0 in f64)", "data": "synthetic",
        "config": {"workload": args.config, "truncation": T, "grid": f"O{N}", "levels": nlev, "fields": nf,
                   "decomposition": f"nprtrw={world},nprtrv=1", "l2": "inputs (GBs) larger than L2, no flush needed"},
        "stages_ms": {"legendre": leg_ms, "fourier": ft_ms, "transpose": tp_ms,
                      "prologue": med("inv", "prologue"), "epilogue": med("dir", "epilogue")},
        # the dominant kernel of the step is the Fourier stage (half of it): its roofline is the headline one
        "roofline": {"bound": "hbm", "kernel": "k_fourier<inv>+k_fourier<dir> (all launches of the Fourier stage)", "achieved": ft_ach,
                     "peak": hbm_peak, "unit": "GB/s", "frac": ft_ach / hbm_peak, "traffic": ft_traffic,
                     "algorithmic_bytes_per_step": fb, "share_of_step": ft_ms / max(ms_dev, 1e-9),
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if hbm_measured else "fallback 6650 GB/s (of fallback)",
                     "note": "arithmetic (FP64) bound, not HBM bound: see DESIGN.md section 4 and profiles/r02_fourier_cz.md"},
        "roofline_legendre": {"bound": "tensor", "kernel": ("k_leinv+k_ledir (FP64 DMMA m8n8k4)" if prec == "dp" else
                                                            "k_leg_tc (tcgen05 kind::tf32, 3xTF32) + m = 0 on DMMA + operand split kernels"),
                              "achieved": ach, "peak": peak_dmma if prec == "dp" else None, "unit": "TFLOP/s",
                              "frac": ach / peak_dmma if prec == "dp" else None, "traffic": traffic,
                              "peak_source": "ect_measure_fp64_peak(DMMA) measured live in this run (MEASURED_PEAKS.json has no FP64 figure)",
                              "dfma_peak": peak_dfma, "share_of_step": leg_ms / max(ms_dev, 1e-9)},
        "gpu_launches": launches * args.steps,
        "clocks": clocks, "setup_s": setup_s,
    }
    if e2e:
        line["e2e"] = e2e
    if parity is not None:
        line["parity"] = parity
    if per_rank is not None:
        line["stages_ms_per_rank"] = per_rank
    if world == 1 and not args.no_cpu:
        stride = args.cpu_stride or CPU_STRIDE.get(T, 16)
        v, detail = cpu_sample_step(T, N, nuv, nsc, stride)
        line["cpu_baseline"] = {"value": v, "unit": "ms", "cores": host_cores(), "kind": "port", "extrapolated": stride != 1,
                                "sample": cpu_sample_text(stride, nf), "detail": detail}
    print(json.dumps(line), flush=True)
    tr.release()


def _shutdown():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="TCo1279_O1280_L137", choices=list(CONFIGS))
    ap.add_argument("--precision", default=None, choices=["dp", "sp"], help="default: the config's (TCo399, TCo2559: sp)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity gate that runs before the timed region")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed end-to-end steps (default: --steps)")
    ap.add_argument("--cpu-stride", type=int, default=0)
    ap.add_argument("--stage-timings", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        try:
            run_ours(args)
        finally:
            _shutdown()          # no "destroy_process_group() was not called" warning after the JSON line


if __name__ == "__main__":
    main()
