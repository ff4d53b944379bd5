"""Print a per-launch table from an `ncu --metrics gpu__time_duration.sum,... --csv` log."""
import csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
gi, bi, ii = hdr.index("Grid Size"), hdr.index("Block Size"), hdr.index("ID")
out = {}
for r in rows[1:]:
    d = out.setdefault(r[ii], {"k": r[ki][:58], "grid": r[gi], "block": r[bi]})
    d[r[mi]] = (r[vi], r[ui])
for i, d in out.items():
    t = d.get("gpu__time_duration.sum", ("?", ""))
    rd = d.get("dram__bytes_read.sum", ("?", "")); wr = d.get("dram__bytes_write.sum", ("?", ""))
    print(i, d["k"], d["grid"], d["block"], "time", *t, "rd", *rd, "wr", *wr)
