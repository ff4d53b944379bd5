"""profiles/sass_markers.txt: per kernel family, the SASS mnemonics that show what the hardware path is
(UBLKCP = cp.async.bulk / TMA bulk copy, SYNCS = mbarrier, LDTM / STTM = tcgen05.ld / tcgen05.st (tensor memory used as
staging storage by the whole-axis kernel), LDGSTS = cp.async, LDG/STG .128 = 128-bit global accesses,
DFMA/DADD/DMUL = fp64 pipe, FFMA = fp32; UTC*MMA / UTMALDG would be tensor-core MMA / tensor-map TMA: an FFT has no
contraction, none is expected).  Usage: python tools/sass_markers.py > profiles/sass_markers.txt"""
import collections, glob, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fam = collections.OrderedDict()
for obj in sorted(glob.glob(os.path.join(ROOT, "impulse_b200", "csrc", "build", "*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            d = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            d = re.sub(r"^void ", "", d)
            d = re.sub(r"\(.*$", "", d)
            tm = re.match(r"(impulse::\w+)<(\w+)", d)
            name = (os.path.basename(obj), f"{tm.group(1)}<{tm.group(2)},...>" if tm else d)
            fam.setdefault(name, [0, collections.Counter()])[0] += 1
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            c = fam[name][1]
            base = op.split(".")[0]
            c[base] += 1
            if base in ("LDG", "STG") and ".128" in op:
                c[base + ".128"] += 1
            if base in ("LDS", "STS") and ".128" in op:
                c[base + ".128"] += 1
print(__doc__.split("Usage")[0].strip())
print()
keys = ["UBLKCP", "UTMALDG", "SYNCS", "LDTM", "STTM", "LDGSTS", "LDG", "LDG.128", "STG", "STG.128", "LDS.128", "STS.128", "DFMA", "DADD", "DMUL", "FFMA", "FADD", "BAR", "ATOMG", "REDG", "MEMBAR", "ERRBAR"]
print(f"{'object':24s} {'kernel family':48s} {'inst':>5s} " + " ".join(f"{k:>8s}" for k in keys))
for (obj, k), (n, c) in fam.items():
    print(f"{obj:24s} {k:48s} {n:5d} " + " ".join(f"{c[x]:8d}" for x in keys))
tot = collections.Counter()
for (_, _), (n, c) in fam.items():
    tot.update(c)
print()
print("any tensor-core MMA (UTC*MMA/HMMA):", sum(v for k, v in tot.items() if "MMA" in k), "| tensor-map TMA (UTMALDG/UTMASTG):", tot["UTMALDG"] + tot["UTMASTG"],
      "| bulk TMA copies (UBLKCP):", tot["UBLKCP"], "| mbarrier ops (SYNCS):", tot["SYNCS"], "| cp.async (LDGSTS):", tot["LDGSTS"],
      "| tensor-memory loads / stores (LDTM / STTM):", tot["LDTM"], "/", tot["STTM"])
