"""Per-kernel SASS opcode census of the shipped libskb.so (cuobjdump -sass): how many tcgen05 MMAs (UTCIMMA / UTCHMMA ...),
TMEM loads (LDTM), FP64 tensor MMAs (DMMA), bulk-copy TMA instructions (UBLKCP), mbarrier waits (SYNCS) each kernel holds.
Usage: python tools/sass_opcodes.py > profiles/rNN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "scikit-optimize_b200", "csrc", "libskb.so")
OPS = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCCP", "DMMA", "UBLKCP", "UTMALDG", "SYNCS", "DFMA", "DADD",
       "DMUL", "MUFU", "I2F", "F2I", "REDUX", "SHFL", "BAR"]
sass = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True, check=True).stdout
demangle = lambda s: subprocess.run(["c++filt", s], stdout=subprocess.PIPE, text=True).stdout.strip()   # noqa: E731
counts, order, cur = {}, [], None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if cur and m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for o in OPS:
            if op.startswith(o):
                counts[cur][o] += 1
                break
print("# %s: %d kernels; columns: instructions, then the opcodes with a non-zero count" % (os.path.relpath(so, ROOT), len(order)))
tot = collections.Counter()
for fn in order:
    c = counts[fn]
    tot.update(c)
    name = re.sub(r"\((?!anonymous).*", "", demangle(fn))
    print("%-70s %6d  %s" % (name, c["_total"], "  ".join("%s=%d" % (o, c[o]) for o in OPS if c[o])))
print("%-70s %6d  %s" % ("TOTAL", tot["_total"], "  ".join("%s=%d" % (o, tot[o]) for o in OPS if tot[o])))
