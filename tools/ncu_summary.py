"""Summarise ncu CSV exports (launch list or --page raw) into small text tables for profiles/."""
import csv, sys

def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
    tot = {}
    for r in rows[1:]:
        if len(r) <= iv: continue
        try: v = float(r[iv].replace(",", ""))
        except ValueError: continue
        k = r[ik][:70]; t = tot.setdefault(k, [0.0, 0]); t[0] += v; t[1] += 1
    s = sum(v[0] for v in tot.values())
    print("%-72s %4s %12s %7s" % ("kernel", "n", "ms (sum)", "share"))
    for k, v in sorted(tot.items(), key=lambda x: -x[1][0]):
        print("%-72s %4d %12.3f %6.1f%%" % (k, v[1], v[0] / 1e6, 100 * v[0] / s))
    print("total %.3f ms" % (s / 1e6))

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit", "launch__grid_size", "launch__block_size",
        "smsp__average_warp", "smsp__warp_issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe", "lts__t_sector_hit_rate.pct", "launch__shared_mem_per_block", "sm__cycles_active.avg", "smsp__issue_active.avg.pct",
        "smsp__cycles_active.avg", "launch__waves_per_multiprocessor", "smsp__average_warps_issue_stalled"]

def raw(path, filt=None):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]; units = rows[1]
    ik = hdr.index("Kernel Name")
    for r in rows[2:]:
        if filt and filt not in r[ik]: continue
        print("=== %s  id=%s" % (r[ik][:90], r[0]))
        for i, h in enumerate(hdr):
            if any(h.startswith(k) for k in KEYS):
                print("   %-95s %18s %s" % (h, r[i], units[i]))

if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](*sys.argv[2:])
