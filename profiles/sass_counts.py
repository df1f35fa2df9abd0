#!/usr/bin/env python
"""Counts the tcgen05 / TMEM / TMA SASS mnemonics per kernel of libpvsr.so (cuobjdump -sass), as evidence that the
convolution launches are tcgen05 code (B200_PROFILING.md: UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA
load, UTCBAR = tcgen05.commit / mbarrier arrive, HMMA = mma.sync).  Usage: python profiles/sass_counts.py > profiles/r02/sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "efficient-and-phase-aware-video-super-resolution-for-cardiac-mri_b200", "csrc", "libpvsr.so")
MNEMONICS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMALDG.2CTA", "UTMASTG", "UTCBAR", "UTCBAR.2CTA",
             "SYNCS", "HMMA", "LDSM", "LDGSTS", "RED", "ATOMG", "MUFU.EX2", "MUFU.RCP", "STG", "LDG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    names = sorted(set(re.findall(r"Function : (\S+)", out)))
    dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.strip().splitlines()
    for n, d in zip(names, dm):
        demangle[n] = d.replace("pvsr::", "").replace("(pvsr::ConvMaps, pvsr::ConvParams)", "").replace("void ", "")
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = demangle.get(m.group(1), m.group(1))
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for key in MNEMONICS:
            if "." in key:
                base, suf = key.split(".", 1)
                if op.startswith(base) and ("." + suf) in op:
                    counts[cur][key] += 1
            elif op.split(".")[0] == key:
                counts[cur][key] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS mnemonic counts per kernel (cuobjdump -sass; profiles/sass_counts.py)")
    print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA) load,")
    print("# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = warp-level mma.sync, LDSM = ldmatrix, RED = red.global")
    total = collections.Counter()
    for k, c in counts.items():
        total.update(c)
        row = "  ".join(f"{m}={c[m]}" for m in MNEMONICS if c[m])
        print(f"{k[:110]:110s} {row}")
    print("TOTAL " + "  ".join(f"{m}={total[m]}" for m in MNEMONICS if total[m]))


if __name__ == "__main__":
    main()
