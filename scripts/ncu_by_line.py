#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS table of one kernel with `nvdisasm -gi` line info of the cubin:
stall samples and executed instructions per source line of the KERNEL BODY (the outermost inlined-at line), so that a
hot helper is charged to the call site that reached it.
usage: ncu_by_line.py <ncu_source.csv> <nvdisasm_gi.sass> <mangled-kernel-substring> [top]"""
import csv, re, sys
from collections import defaultdict

src_csv, sass, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# offset -> (outer line, inner line)
line_of = {}
cur_inner = cur_outer = None
insec = False
pend = []
for ln in open(sass):
    if ln.lstrip().startswith(".section"):
        insec = (".text." in ln and kname in ln)
        continue
    if not insec:
        continue
    m = re.search(r'//## File "[^"]*/([^"/]+)", line (\d+)( inlined at "[^"]*/([^"/]+)", line (\d+))?', ln)
    if m:
        pend.append((m.group(1), int(m.group(2)), m.group(4), int(m.group(5)) if m.group(5) else None))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', ln)
    if m:
        if pend:
            # the first record is the innermost location; the last record's own line is the outermost frame
            cur_inner = (pend[0][0], pend[0][1])
            cur_outer = (pend[-1][0], pend[-1][1]) if pend[-1][2] is None else (pend[-1][2], pend[-1][3])
            pend = []
        line_of[int(m.group(1), 16)] = (cur_outer, cur_inner, ln.split("*/", 1)[1].strip()[:60])
rows = list(csv.reader(open(src_csv)))
for i in range(2, len(rows)):          # several launches in one export: keep the first
    if rows[i] and rows[i][0] == "Kernel Name":
        rows = rows[:i]
        break
hdr = rows[1]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[2][ia], 16)
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
inner = defaultdict(lambda: [0, 0])
tot_s = tot_i = 0
for r in rows[2:]:
    off = int(r[ia], 16) - base
    lo = line_of.get(off)
    if lo is None:
        continue
    s, ex = int(r[isamp] or 0), int(r[iex] or 0)
    a = agg[lo[0]]
    a[0] += s
    a[1] += ex
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            a[2][hdr[i]] += v
    k = inner[(lo[0], lo[1])]
    k[0] += s
    k[1] += ex
    tot_s += s
    tot_i += ex
print(f"total samples {tot_s}, warp instructions {tot_i}")
print("== by kernel-body line (outermost frame)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ", ".join(f"{n[6:]} {v}" for n, v in sorted(a[2].items(), key=lambda kv: -kv[1])[:4])
    print(f"{k[0]}:{k[1]:4d}  samples {a[0]:6d} ({100.0*a[0]/max(tot_s,1):4.1f}%)  inst {a[1]:8d} ({100.0*a[1]/max(tot_i,1):4.1f}%)  {st}")
print("== by (body line, innermost line)")
for k, a in sorted(inner.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0][0]}:{k[0][1]:4d} <- {k[1][0]}:{k[1][1]:4d}  samples {a[0]:6d}  inst {a[1]:8d}")
