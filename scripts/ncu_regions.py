"""dev: instruction / stall-sample share per marked region of a source file, from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`.
    python scripts/ncu_regions.py src.csv file.cu "name=substring of the first line" ...
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
fname = sys.argv[2]
hdr = ix = cur = None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        cur = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {n: i for i, n in enumerate(hdr)}
        continue
    if hdr is None or len(r) < len(hdr) - 2 or r[0] == "":
        continue
    try:
        inst = int(r[ix["Instructions Executed"]] or 0)
        s = int(r[ix["# Samples"]] or 0)
        th = int(r[ix["Thread Instructions Executed"]] or 0)
    except Exception:
        continue
    a = agg.setdefault((cur.split("/")[-1], int(r[0])), [0, 0, 0])
    a[0] += inst
    a[1] += s
    a[2] += th
tot = sum(a[0] for a in agg.values())
tots = sum(a[1] for a in agg.values())
print(f"total warp instructions {tot}, samples {tots}, avg threads/inst {sum(a[2] for a in agg.values()) / max(tot, 1):.1f}")
src = open(fname).read().split("\n")
base = fname.split("/")[-1]
marks = []
for spec in sys.argv[3:]:
    name, sub = spec.split("=", 1)
    marks.append((name, next(i + 1 for i, l in enumerate(src) if sub in l)))
marks.sort(key=lambda m: m[1])
marks.append(("end", len(src) + 1))
for i, (name, l0) in enumerate(marks[:-1]):
    l1 = marks[i + 1][1]
    sel = [a for (f, ln), a in agg.items() if f == base and l0 <= ln < l1]
    inst, s, th = sum(a[0] for a in sel), sum(a[1] for a in sel), sum(a[2] for a in sel)
    print(f"{name:26s} lines {l0:4d}-{l1:4d}: inst {100 * inst / tot:5.1f}%  samples {100 * s / tots:5.1f}%  threads/inst {th / max(inst, 1):4.1f}")
for f in sorted(set(k[0] for k in agg)):
    if f != base:
        print(f"{f:26s} inst {100 * sum(a[0] for (ff, ln), a in agg.items() if ff == f) / tot:5.1f}%")
