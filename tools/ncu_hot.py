"""Top stall-sample SASS lines of one kernel of an ncu --set full --import-source on report.
  python tools/ncu_hot.py gpurun_out/prof_r1l.ncu-rep <launch id> [n]"""
import csv
import subprocess
import sys

rep, kid = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, seen = [], set()      # the csv repeats most kernels twice; keep the first copy of each launch (keyed by its first address)
for b in raw.split('"Kernel Name",')[1:]:
    key = b[:4000]
    if key not in seen:
        seen.add(key)
        blocks.append(b)
lines = blocks[int(kid)].splitlines()
print(lines[0][:100])
rows = list(csv.reader(lines[1:]))
hdr = rows[0]
ia, isrc, iall, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[1:] if len(r) > iall and r[iall].isdigit()]
tot = sum(int(r[iall]) for r in data)
print("total samples", tot, "instructions", sum(int(r[iex]) for r in data))
agg = {}
for r in data:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
idx = {id(r): k for k, r in enumerate(data)}
for r in sorted(data, key=lambda r: -int(r[iall]))[:n]:
    top = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print("%5d %5.1f%% #%-5d ex=%-8s %-60s %s" % (int(r[iall]), 100.0 * int(r[iall]) / tot, idx[id(r)], r[iex], r[isrc].strip()[:60], top))
