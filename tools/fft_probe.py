"""Phase timing of the Fourier kernel per shared-memory class (debug probe: block 0, first pair)."""
import os, sys, ctypes, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    import ectrans_b200 as eb
    L = eb.lib()
    T, N, nf = int(os.environ.get("PROBE_T", 1279)), int(os.environ.get("PROBE_N", 1280)), 32
    prec = os.environ.get("PROBE_PREC", "dp")
    tdt = torch.float32 if prec == "sp" else torch.float64
    tr = eb.Transform(T, eb.octahedral_nloen(N), stream=torch.cuda.current_stream().cuda_stream, precision=prec)
    sc = (torch.rand((tr.nspec2, nf), device="cuda", dtype=tdt) - 0.5)
    gp = torch.empty((1, nf, tr.ngptot), dtype=tdt, device="cuda")
    for direction in ("inv", "dir"):
        for it in range(2):
            if direction == "inv": tr.inv_trans(spscalar=sc, out=gp)
            else: tr.dir_trans(gp, 0, nf, out=(None, None, sc))
        torch.cuda.synchronize()
        t = tr.timings()
        buf = (ctypes.c_longlong * 64)()
        L.ect_debug_fft_probe(buf)
        v = np.array(buf[:])
        d = {"load": int(v[1] - v[0])}
        prev = v[1]
        for i in range(2, 10):
            if v[i] > prev: d[f"dif{i-2}"] = int(v[i] - prev); prev = v[i]
        if v[10] > prev: d["middle"] = int(v[10] - prev); prev = v[10]
        for i in range(11, 20):
            if v[i] > prev: d[f"dit{i-10}"] = int(v[i] - prev); prev = v[i]
        d["store"] = int(v[20] - prev); d["total"] = int(v[20] - v[0])
        print("bucket", os.environ.get("ECT_FFT_ONLY_BUCKET"), direction, "fourier ms %.3f" % t["fourier"], d, flush=True)
else:
    for b in range(14):
        env = dict(os.environ, ECT_FFT_ONLY_BUCKET=str(b))
        subprocess.run([sys.executable, __file__, "child"], env=env)
