"""Condense `ncu --page source --csv` into the 40 hottest source lines (warp stall samples) with their dominant stall
reasons.  Usage: python tools/ncu_hot_lines.py source.csv"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = None
for i, r in enumerate(rows):
    if "Source" in r and any("Sampling" in c or "Samples" in c for c in r):
        hdr = i
        break
if hdr is None:
    print("no source page header found; first rows:")
    for r in rows[:5]:
        print(r[:12])
    sys.exit(0)
H = rows[hdr]
col = {c: j for j, c in enumerate(H)}
samp = next((c for c in H if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)" or "Sampling (All" in c), None)
stall_cols = [c for c in H if c.startswith("stall_") or c.startswith("Stall")]
agg = defaultdict(lambda: [0.0, defaultdict(float), ""])
total = 0.0
for r in rows[hdr + 1:]:
    if len(r) < len(H):
        continue
    try:
        s = float(r[col[samp]] or 0) if samp else 0.0
    except ValueError:
        continue
    key = r[col.get("Address", 0)] if "Source" not in col else r[col["Source"]][:110]
    a = agg[key]
    a[0] += s
    total += s
    for c in stall_cols:
        try:
            a[1][c] += float(r[col[c]] or 0)
        except ValueError:
            pass
print(f"total samples {total:.0f}; columns: {samp}")
for key, (s, st, _) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{100 * s / max(total, 1):5.1f}%  {key}   " + ", ".join(f"{k}={v:.0f}" for k, v in top if v > 0))
