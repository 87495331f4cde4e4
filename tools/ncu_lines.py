"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line: samples, instructions, top stall reasons.
usage: ncu_lines.py dump.csv [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.OrderedDict()
fname = None
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]; continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    key = (fname, line)
    a = agg.setdefault(key, {"src": r[1], "samples": 0, "inst": 0, "stalls": collections.Counter()})
    si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
    try:
        a["samples"] += int(r[si] or 0); a["inst"] += int(r[ii] or 0)
    except ValueError:
        pass
    for j, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            try: a["stalls"][h[6:]] += int(r[j] or 0)
            except ValueError: pass
tot = sum(a["samples"] for a in agg.values()) or 1
print(f"total samples {tot}")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = ", ".join(f"{k}={v}" for k, v in a["stalls"].most_common(3))
    print(f"{100*a['samples']/tot:5.1f}%  inst={a['inst']:9d}  {f}:{l:<5d} {a['src'].strip()[:110]}   [{st}]")
