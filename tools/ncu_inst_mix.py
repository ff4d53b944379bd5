"""From `ncu --page source --csv`: executed thread instructions per SASS opcode and the 30 most executed lines.
Usage: python tools/ncu_inst_mix.py source.csv"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Source" in r and any("Instructions Executed" in c for c in r))
H = rows[hdr]
ci = H.index("Source")
ce = next(j for j, c in enumerate(H) if c.strip() == "# Instructions Executed" or c.strip() == "Instructions Executed")
ct = next((j for j, c in enumerate(H) if "Thread Instructions Executed" in c and "Pred" not in c), ce)
ops, lines, tot = collections.Counter(), [], 0
for r in rows[hdr + 1:]:
    if len(r) < len(H):
        continue
    try:
        n = float(r[ct] or 0)
    except ValueError:
        continue
    src = r[ci].strip()
    m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
    op = m.group(1) if m else "?"
    ops[op] += n
    tot += n
    lines.append((n, src[:90]))
print(f"thread instructions executed: {tot:.4g}")
for op, n in ops.most_common(28):
    print(f"{100 * n / tot:5.1f}%  {op}")
print("--- most executed lines")
for n, src in sorted(lines, reverse=True)[:30]:
    print(f"{100 * n / tot:5.2f}%  {src}")
