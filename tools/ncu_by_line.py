"""Dev tool: join the SASS rows of an `ncu --page source --csv` export with the line table of
`nvdisasm -g` and aggregate samples / shared-memory wavefronts / global requests per source line.

  ncu -i rep.ncu-rep --page source --csv > src.csv
  cuobjdump -xelf all lib.so; nvdisasm -g -c x.cubin > all.sass      (cut to the kernel)
  python tools/ncu_by_line.py src.csv kern.sass [source.cuh]
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass = sys.argv[1], sys.argv[2]
source = sys.argv[3] if len(sys.argv) > 3 else None
# line table: offset -> line
line_of, cur = {}, None
for ln in open(sass):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = int(m.group(2))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
# the export holds one table per profiled launch: take the LAST launch whose kernel name contains
# $NCU_KERNEL (default: the last table)
import os
want = os.environ.get("NCU_KERNEL", "")
tables, name = [], ""
for i, r in enumerate(rows):
    if r and r[0] == "Kernel Name":
        name = r[1]
    if "Address" in r and "Source" in r:
        tables.append((name, i))
sel = [t for t in tables if want in t[0]] or tables
hdr, start = rows[sel[-1][1]], sel[-1][1] + 1
end = min([t[1] for t in tables if t[1] > sel[-1][1]] + [len(rows)])
rows = rows[:end]
print("kernel:", sel[-1][0][:100])
ix = {n: i for i, n in enumerate(hdr)}
base = None
agg = defaultdict(lambda: defaultdict(float))
cols = ["# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal",
        "L1 Tag Requests Global", "stall_long_sb", "stall_barrier", "stall_short_sb", "stall_wait",
        "stall_mio", "stall_lg", "stall_math", "stall_not_selected"]
cols = [c for c in cols if c in ix]
for r in rows[start:]:
    if len(r) != len(hdr):
        continue
    a = int(r[ix["Address"]], 16)
    if base is None:
        base = a
    line = line_of.get(a - base)
    for c in cols:
        try:
            agg[line][c] += float(r[ix[c]])
        except ValueError:
            pass
text = {}
if source:
    for n, l in enumerate(open(source), 1):
        text[n] = l.rstrip()
tot = {c: sum(agg[l][c] for l in agg) for c in cols}
print("totals:", {c: int(v) for c, v in tot.items()})
key = sys.argv[4] if len(sys.argv) > 4 else "# Samples"
print("%5s " % "line" + " ".join("%9s" % c.replace("L1 Wavefronts Shared", "shWF").replace("Instructions Executed", "inst")
                                  .replace("L1 Tag Requests Global", "gTag").replace("stall_", "")[:9] for c in cols))
for l in sorted(agg, key=lambda l: -agg[l][key])[:60]:
    print("%5s " % l + " ".join("%9d" % agg[l][c] for c in cols) + "  | " + text.get(l, "")[:90])
