"""Top sampled SASS lines of an .ncu-rep (source page): python profiles/top_stalls.py rep.ncu-rep [n]"""
import csv
import subprocess
import sys


def main(path, n=25):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    tot = sum(int(r[isamp] or 0) for r in body)
    print(f"total samples {tot}")
    idx = {id(r): k for k, r in enumerate(body)}
    for r in sorted(body, key=lambda r: -int(r[isamp] or 0))[:n]:
        st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        print(f"{idx[id(r)]:5d} {int(r[isamp]):7d} {100.0 * int(r[isamp]) / tot:5.1f}% ex={r[iex]:>9s} {r[isrc].strip()[:70]:70s} {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
