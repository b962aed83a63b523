#!/usr/bin/env python
"""SASS mnemonic counts per kernel of miso_b200/libmiso_b200.so (cuobjdump -sass | c++filt): the evidence that the hot
path uses tcgen05 / TMEM / vector reductions.   python tools/sass_counts.py > profiles/rNN_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "miso_b200", "libmiso_b200.so")
WANT = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "REDG.E.ADD.F32x4", "REDG", "ATOMG", "LDG.E.128", "STG.E.128", "LDS.128", "STS.128",
        "SHFL", "BAR.SYNC", "FFMA", "FFMA2", "ATOMS", "UTMALDG", "UTMASTG", "UTMAREDG"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                           text=True, check=True).stdout.split("\n")
    counts, cur, it = collections.OrderedDict(), None, iter(names)
    for line in sass.split("\n"):
        if "Function : " in line:
            cur = re.sub(r"\(.*", "", next(it))
            counts[cur] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
        if cur is None or not m:
            continue
        op = m.group(1)
        for w in WANT:
            if op == w or op.startswith(w + "."):
                counts[cur][w] += 1
                break
    print("# SASS mnemonic counts per kernel of miso_b200/libmiso_b200.so (cuobjdump -sass, sm_100a; tools/sass_counts.py)")
    print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UTCBAR = tcgen05.commit, REDG.E.ADD.F32x4 = red.global.add.v4.f32")
    print("# no UTMA* in the product: gathers/scatters are irregular per sample (TMA tensor reduce was probed and rejected,")
    print("# profiles/r02_tma_reduce_probe.json)\n")
    for k, c in counts.items():
        if c:
            print(f"{k}: " + ", ".join(f"{w}={c[w]}" for w in WANT if c[w]))


if __name__ == "__main__":
    main()
