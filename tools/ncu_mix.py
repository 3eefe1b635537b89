"""Executed-instruction mix of one kernel of an ncu --set full --import-source on report, per `units` outputs:
  python tools/ncu_mix.py gpurun_out/prof_r1n.ncu-rep <launch id> <units>      e.g. units = M*N/32 for a GEMM (warp-level outputs)
Comparing the per-output counts with what the source should need is how the wasted instructions of r1n-r1p were found."""
import collections
import csv
import subprocess
import sys

rep, kid, units = sys.argv[1], int(sys.argv[2]), float(eval(sys.argv[3]))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, seen = [], set()
for b in raw.split('"Kernel Name",')[1:]:
    if b[:4000] not in seen:
        seen.add(b[:4000])
        blocks.append(b)
rows = list(csv.reader(blocks[kid].splitlines()[1:]))
hdr, data = rows[0], rows[1:]
isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
c, tot = collections.Counter(), 0
for r in data:
    if len(r) > iex and r[iex].isdigit():
        op = r[isrc].split()
        c[op[1] if op[0].startswith("@") else op[0]] += int(r[iex])
        tot += int(r[iex])
print(blocks[kid].splitlines()[0][:100])
print("total %d warp instructions, %.2f per unit" % (tot, tot / units))
for k, v in c.most_common(40):
    print("  %-30s %12d %7.2f" % (k, v, v / units))
