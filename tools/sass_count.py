"""Static SASS statistics per kernel of an object file: total instructions and a few mnemonic classes.
Usage: python tools/sass_count.py build/fast_kernels.o [regex]"""
import re, subprocess, sys, collections
obj = sys.argv[1]
rx = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
name, stats = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        stats[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1).split(".")[0]
        stats[name]["total"] += 1
        stats[name][op] += 1
for n, c in stats.items():
    d = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    if rx and not rx.search(d):
        continue
    keys = ["total", "FFMA", "FADD", "FMUL", "FFMA2", "FADD2", "FMUL2", "DFMA", "DADD", "DMUL", "LDS", "STS", "LDG", "STG", "IMAD", "IADD3", "MOV", "LEA", "UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "BAR"]
    print(d[:110], " ".join(f"{k}={c[k]}" for k in keys if c[k]))
