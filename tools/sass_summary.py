"""Opcode evidence per kernel of the shipped library: counts of the Blackwell-native SASS mnemonics (UTC*MMA = tcgen05.mma,
UTMALDG / UTMASTG = TMA, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit) and of the legacy paths (IDP.4A = dp4a, HMMA).
  python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "p2vit_b200", "csrc", "libp2vit_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UTCIMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "ACQBULK", "IDP.4A", "IDP.2A", "HMMA", "IMMA", "FFMA", "MUFU", "I2F", "F2I", "FRND", "REDUX", "ATOM", "RED"]
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                counts[kern][w] += 1
print("library:", os.path.relpath(lib, ROOT), " arch:", "sm_100a" if "sm_100a" in sass else "?")
print("%-78s %7s  %s" % ("kernel", "instrs", "watched opcodes"))
agg = collections.Counter()
for k, c in counts.items():
    agg.update(c)
    print("%-78s %7d  %s" % (k[:78], total[k], " ".join("%s:%d" % (w, c[w]) for w in WATCH if c[w])))
print("\nTOTAL", " ".join("%s:%d" % (w, agg[w]) for w in WATCH if agg[w]))
