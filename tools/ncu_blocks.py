"""Basic-block view of `ncu --page source --csv --print-source sass`: consecutive SASS lines with the same executed count,
with their share of executed instructions / stall samples and the top stall reasons.  Usage: ncu_blocks.py file.csv [min%]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.2
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; S = hdr.index("Source"); N = hdr.index("Instructions Executed"); P = hdr.index("# Samples")
stall_cols = [(k, hdr.index(k)) for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
body = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
blocks = []; cur = None
for i, r in enumerate(body):
    n = int(r[N] or 0)
    if cur is None or cur["n"] != n:
        cur = {"n": n, "start": i, "samples": 0, "len": 0, "st": collections.Counter(), "fp": 0, "first": r[S].strip()[:40]}; blocks.append(cur)
    cur["samples"] += int(r[P] or 0); cur["len"] += 1
    if re.search(r"\b(DFMA|DMUL|DADD)\b", r[S]): cur["fp"] += 1
    for k, c in stall_cols: cur["st"][k] += int(r[c] or 0)
tots = sum(b["samples"] for b in blocks); tot = sum(b["n"] * b["len"] for b in blocks)
allst = collections.Counter()
for b in blocks: allst.update(b["st"])
print("total warp instructions", f"{tot:,}", "samples", tots)
print({k.replace("stall_", ""): round(100 * v / tots, 1) for k, v in allst.most_common(12)})
for b in blocks:
    if b["samples"] > tots * thr / 100 or b["n"] * b["len"] > tot * thr / 100:
        print(f"@{b['start']:5d} len {b['len']:4d} fp {b['fp']:4d} exec {b['n']:>11,} inst% {100*b['n']*b['len']/tot:5.1f} samp% {100*b['samples']/tots:5.1f}",
              {k.replace("stall_", ""): round(100 * v / tots, 1) for k, v in b["st"].most_common(3)}, b["first"])
