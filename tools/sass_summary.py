"""Per-kernel SASS census of libgapb200.so (cuobjdump -sass): the instructions that show which hardware paths a kernel uses.
  DMMA.8x8x4      FP64 tensor-core MMA (mma.sync.m8n8k4.f64)
  UTMALDG         TMA load (cp.async.bulk.tensor)           SYNCS   mbarrier operations
  ACQBULK/PREEXIT griddepcontrol.wait / .launch_dependents (programmatic dependent launch)
  REDG/ATOMG      global atomics (FP64 force scatter)        ATOMS   shared-memory atomics
usage: python tools/sass_summary.py [path/to/lib.so] > profiles/rNN_sass_summary.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "quip_b200/libgapb200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
WATCH = ["DMMA", "DFMA", "UTMALDG", "SYNCS", "ACQBULK", "PREEXIT", "REDG", "ATOMG", "ATOMS", "LDGSTS", "BAR"]
cur, rows = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        rows[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        rows[cur]["_total"] += 1
        op = m.group(1)
        if op in WATCH:
            rows[cur][op] += 1
print(__doc__)
print("%-9s" % "instr" + "".join("%9s" % w for w in WATCH) + "  kernel")
for k, c in rows.items():
    name = re.sub(r"gapb200::|\(anonymous namespace\)::|<unnamed>::", "", demangle(k))
    name = re.sub(r"\(.*", "", name)
    print("%-9d" % c["_total"] + "".join("%9d" % c[w] for w in WATCH) + "  " + name[:110])
