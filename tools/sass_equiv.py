"""Compare the SASS of every kernel in two cubins (e.g. the last GPU-validated build and the current one) after a
refactoring that should not change generated code:

    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -cubin -o old.cubin <old>/fast_kernels.cu
    nvcc ... -cubin -o new1.cubin fast_kernels.cu ; nvcc ... -cubin -o new2.cubin fast3_kernels.cu
    python tools/sass_equiv.py old.cubin new1.cubin new2.cubin

Kernels are matched by mangled name; trailing boolean template parameters that the new build added with value false
(`ELb0` before the closing `EEEv`) are stripped.  Used at the end of round 1 after the register kernels moved into
headers and a translation unit of their own: 352 kernels, 352 identical instruction streams; and again when the fused
Bluestein launcher moved from fast_kernels.cu to fastblue_kernels.cu (old fast_kernels.cu vs the two new files: 232 of
232 identical)."""
import re
import subprocess
import sys


def parse(cubin):
    sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    fn, d = None, {}
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            d[fn] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and fn:
            d[fn].append(re.sub(r"\s+", " ", m.group(1)))
    return d


def variants(name):
    yield name
    n = name
    for _ in range(3):   # up to three added trailing `false` parameters
        n2 = re.sub(r"ELb0(EEEv)", r"\1", n, count=1)
        if n2 == n:
            break
        n = n2
        yield n


def main():
    old = parse(sys.argv[1])
    new = {}
    for p in sys.argv[2:]:
        new.update(parse(p))
    index = {}
    for k in new:
        for v in variants(k):
            index.setdefault(v, k)
    same = diff = missing = 0
    for k, v in old.items():
        if k not in index:
            missing += 1
            print("MISSING", k)
        elif new[index[k]] == v:
            same += 1
        else:
            diff += 1
            print("DIFFERENT", len(v), len(new[index[k]]), k)
    print(f"old kernels {len(old)}: identical {same}, different {diff}, missing {missing}")
    return 0 if diff == 0 and missing == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
