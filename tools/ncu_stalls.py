"""Warp-stall picture of the kernels in an `ncu --set full --import-source on` report: per kernel the sampled stall reasons
(share of all warp samples; mbarrier sleeps counted separately from the instructions that issue them), executed warp
instructions, shared-memory wavefronts against the ideal count, and the instructions with the most samples.
  python tools/ncu_stalls.py gpurun_out/prof_r2.ncu-rep [kernel-name regex] > profiles/r2_stalls.txt"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
seen = set()
for b in raw.split('"Kernel Name",')[1:]:
    lines = b.splitlines()
    name = lines[0].strip().strip('",')
    short = (name[:name.index(">(") + 1] if ">(" in name else re.sub(r"\(.*", "", name)).replace("void p2v::", "").replace("(int)", "").replace("(bool)", "")
    if short in seen or (pat and not pat.search(short)):
        continue
    seen.add(short)
    rows = [r for r in csv.reader(lines[1:]) if len(r) > 10]
    hdr, data = rows[0], rows[1:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    instr = sum(int(r[ix["Instructions Executed"]]) for r in data)
    c = collections.Counter()
    for r in data:
        if "NANOSLEEP" in r[ix["Source"]]:
            c["mbarrier sleep"] += int(r[ix["# Samples"]])
            continue
        for k in stalls:
            c[k[6:]] += int(r[ix[k]])
    wf = sum(int(r[ix["L1 Wavefronts Shared"]] or 0) for r in data) if "L1 Wavefronts Shared" in ix else 0
    wfi = sum(int(r[ix["L1 Wavefronts Shared Ideal"]] or 0) for r in data) if "L1 Wavefronts Shared Ideal" in ix else 0
    print("== %s" % short)
    print("   %d warp samples, %d executed warp instructions, shared-memory wavefronts %d (ideal %d)" % (tot, instr, wf, wfi))
    print("   " + "  ".join("%s %.1f%%" % (k, 100.0 * v / max(tot, 1)) for k, v in c.most_common(10)))
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:12]:
        why = {k[6:]: int(r[ix[k]]) for k in stalls if int(r[ix[k]])}
        top = max(why, key=why.get) if why else "-"
        print("   %5d samples  x%-9s %-58s %s" % (int(r[ix["# Samples"]]), r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:58], top))
