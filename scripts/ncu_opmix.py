"""Summarise an `ncu --page source --csv` export: executed instructions per opcode, top stall lines."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
ops = collections.Counter(); thr = collections.Counter(); samples = collections.Counter()
total = 0; lines = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.rstrip(";")
    base = op.split(".")[0]
    n = int(r[ix["Instructions Executed"]]); t = int(r[ix["Thread Instructions Executed"]])
    s = int(r[ix["# Samples"]])
    ops[base] += n; thr[base] += t; samples[base] += s; total += n
    lines.append((s, n, src))
print("total warp instructions executed:", total)
for k, v in ops.most_common(40):
    print("%-10s %12d %6.2f%%  samples %7d" % (k, v, 100.0 * v / total, samples[k]))
print("\ntop stall lines:")
for s, n, src in sorted(lines, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("%7d %10d  %s" % (s, n, src))
