"""Samples at every mbarrier wait / MMA / TMA site of an .ncu-rep: python profiles/barrier_waits.py rep.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
body = [r for r in rows[2:] if len(r) == len(hdr)]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[isamp] or 0) for r in body)
print("total samples", tot)
for k, r in enumerate(body):
    s = r[isrc]
    if any(t in s for t in ("TRYWAIT", "UTCHMMA", "UTMALDG", "UTCBAR", "LDTM", "ARRIVE", "BAR.SYNC", "UCGABAR")):
        n = int(r[isamp] or 0) + int(body[k + 1][isamp] or 0) if k + 1 < len(body) else 0
        if n > 0 or "TRYWAIT" in s:
            print(f"{k:5d} samp(+next)={n:7d} ex={r[iex]:>10s}  {s.strip()[:100]}")
