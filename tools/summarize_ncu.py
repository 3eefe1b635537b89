"""Summarises ncu outputs brought back in gpurun_out/ into small text files under profiles/ (tracked).
  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r1a_launches.txt
  python tools/summarize_ncu.py full gpurun_out/prof_r1a.ncu-rep profiles/r1a_ncu_full.txt
"""
import collections
import csv
import subprocess
import sys

mode, src, dst = sys.argv[1:4]
out = []
if mode == "launches":
    rows = list(csv.reader(open(src)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[h], rows[h + 1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) > vi:
            agg.setdefault(r[ki].split("(")[0].replace("void p2v::", "").replace("void ", ""), []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    out.append("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    out.append("%-60s %6s %12s %10s %7s" % ("kernel", "n", "total_ms", "avg_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append("%-60s %6d %12.3f %10.1f %7.3f" % (k[:60], len(v), sum(v) / 1e6, sum(v) / len(v) / 1e3, sum(v) / tot))
else:
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]
    want += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in data:
        out.append("---- " + r[idx["Kernel Name"]][:110])
        for w in want:
            if w in idx and r[idx[w]] not in ("", "0"):
                out.append("  %-88s %s %s" % (w, r[idx[w]], units[idx[w]]))
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
