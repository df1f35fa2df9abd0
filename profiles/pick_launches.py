"""Reads an ncu launch list (gpu__time_duration.sum per launch) and prints, for every distinct kernel of the second
half of the run (the second eager pass of profiles/ncu_target.py), one representative launch id (median duration):
`<id> <tag>` per line.  Only our own kernels (no at:: / nccl)."""
import collections
import csv
import re
import sys


def main(path, top=8):
    rows = []
    for r in csv.reader(open(path)):
        if len(r) >= 15 and r[0].isdigit() and r[12] == "gpu__time_duration.sum":
            rows.append((int(r[0]), r[4], float(r[14].replace(",", ""))))
    half = rows[len(rows) // 2:]
    by = collections.defaultdict(list)
    for i, name, ns in half:
        if name.startswith("void at::") or "nccl" in name or "at::native" in name:
            continue
        by[name].append((ns, i))
    tot = sorted(by.items(), key=lambda kv: -sum(x[0] for x in kv[1]))[:top]
    for name, lst in tot:
        lst.sort()
        ns, i = lst[len(lst) // 2]
        tag = re.sub(r"[^A-Za-z0-9]+", "_", name.split("(")[0].replace("void ", "").replace("pvsr::", ""))[:60].strip("_")
        print(i, tag)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 8)
