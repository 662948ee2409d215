"""dev: aggregate warp-stall samples per CUDA source line from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want_file = sys.argv[2] if len(sys.argv) > 2 else None
cur_file, hdr, ix = None, None, None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {n: i for i, n in enumerate(hdr)}
        continue
    if hdr is None or r[0] == "Function Name" or len(r) < len(hdr) - 2:
        continue
    if r[0] != "":  # a source line row (aggregated over its SASS)
        try:
            s = int(r[ix["# Samples"]])
        except ValueError:
            continue
        key = (cur_file.split("/")[-1], int(r[0]), r[1].strip())
        stalls = {n: int(r[i] or 0) for n, i in ix.items() if n.startswith("stall_") and "Not Issued" not in n and i < len(r)}
        inst = int(r[ix["Instructions Executed"]] or 0)
        a = agg.setdefault(key, [0, 0, {}])
        a[0] += s
        a[1] += inst
        for k, v in stalls.items():
            a[2][k] = a[2].get(k, 0) + v
tot = sum(a[0] for a in agg.values())
print("total samples", tot)
for (f, ln, src), (s, inst, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[: int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    if want_file and want_file not in f:
        continue
    top = sorted(((v, k.replace("stall_", "")) for k, v in st.items()), reverse=True)[:3]
    print(f"{f}:{ln:4d} {s:6d} {100 * s / max(tot, 1):5.1f}% inst {inst:8d}  {src[:80]:80s} {top}")
