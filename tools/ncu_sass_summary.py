"""Summarise `ncu --page source --csv --print-source sass` of one kernel launch: executed warp instructions and stall
samples per opcode, and the hottest SASS lines.  Usage: python tools/ncu_sass_summary.py file.csv [top]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {k: hdr.index(k) for k in ("Source", "# Samples", "Instructions Executed")}
ops = collections.Counter(); samp = collections.Counter(); tot = 0; tots = 0
lines = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    src = r[ci["Source"]].strip()
    m = re.match(r"(@!?U?P[0-9T]\s+)?([A-Z0-9_]+)((\.[A-Z0-9_]+)*)", src)
    op = m.group(2) if m else "?"
    if op == "IMAD" and ".MOV" in (m.group(3) or ""): op = "IMAD.MOV"
    n = int(r[ci["Instructions Executed"]] or 0); s = int(r[ci["# Samples"]] or 0)
    ops[op] += n; samp[op] += s; tot += n; tots += s
    lines.append((s, n, src))
print(f"total warp instructions {tot:,}  samples {tots:,}")
for op, n in ops.most_common(22):
    print(f"  {op:10s} inst {n:14,} {100*n/tot:5.1f}%   samples {samp[op]:8,} {100*samp[op]/max(1,tots):5.1f}%")
print("hottest lines by samples:")
for s, n, src in sorted(lines, reverse=True)[:top]:
    print(f"  {s:7d} {n:12,}  {src[:90]}")
