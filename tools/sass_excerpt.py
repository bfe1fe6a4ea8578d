"""SASS evidence per hot kernel: counts of the mnemonics that prove the hardware path (B200_PROFILING.md) and their first
occurrences.  python tools/sass_excerpt.py > profiles/r02_sass_excerpts.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "ectrans_b200", "lib", "libectrans_b200.so")], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
print("# cuobjdump -sass ectrans_b200/lib/libectrans_b200.so (sm_100a): instruction counts per kernel and first occurrences of the")
print("# tensor / TMA / cluster instructions.  DMMA = mma.sync f64; UTCHMMA = tcgen05.mma; UTMALDG = cp.async.bulk.tensor;")
print("# LDTM = tcgen05.ld; LDGSTS = cp.async; UCGABAR = barrier.cluster.\n")
keys = ["DMMA", "UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "LDGSTS", "UBLKCP", "SYNCS", "UCGABAR", "DFMA", "DADD", "DMUL", "FFMA", "LDS", "STS", "BAR.SYNC"]
want = ("k_leinv", "k_ledir", "k_leg_tc", "k_fourier", "k_tc_", "k_ltinv", "k_ltdir")
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    if not any(w in name for w in want):
        continue
    c, first, n = collections.Counter(), {}, 0
    for line in f.split("\n"):
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        n += 1
        op = m.group(1)
        for k in keys:
            if op.startswith(k):
                c[k] += 1
                if k in ("DMMA", "UTCHMMA", "UTMALDG", "LDTM", "UCGABAR", "UBLKCP") and k not in first:
                    first[k] = re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", line.strip())[:120]
    if n:
        print(f"{name}: {n} instructions; " + ", ".join(f"{k} {c[k]}" for k in keys if c[k]))
        for v in first.values():
            print("      " + v)
