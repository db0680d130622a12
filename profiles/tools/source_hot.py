"""Reads `ncu -i X.ncu-rep --page source --csv --print-source cuda` style CSV on stdin (one table per kernel) and prints the
source lines with the most warp-stall samples / instructions.  usage: source_hot.py [top_n]"""
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
hdr_i = [i for i, r in enumerate(rows) if "Source" in r and any("Sampl" in c for c in r)]
for t, hi in enumerate(hdr_i):
    names = rows[hi]
    end = hdr_i[t + 1] if t + 1 < len(hdr_i) else len(rows)
    si = names.index("Source")
    samp = next(i for i, c in enumerate(names) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)" or "Sampling (All" in c)
    inst = next((i for i, c in enumerate(names) if c == "Instructions Executed" or c.startswith("Instructions Executed")), None)
    body = [r for r in rows[hi + 1:end] if len(r) == len(names)]
    def num(x):
        try:
            return float(x.replace(",", ""))
        except ValueError:
            return 0.0
    tot = sum(num(r[samp]) for r in body) or 1.0
    print(f"== table {t}: {len(body)} lines, {tot:.0f} samples; columns: {names[samp]!r} / {names[inst] if inst is not None else None!r}")
    for r in sorted(body, key=lambda r: -num(r[samp]))[:top]:
        print(f"{100 * num(r[samp]) / tot:5.1f}%  inst={r[inst] if inst is not None else '':>12}  {r[si].strip()[:150]}")
