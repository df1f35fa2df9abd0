"""Aggregates an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_*]` launch list per kernel name:
python profiles/aggregate_launches.py launches.csv [first_id last_id]"""
import collections
import csv
import sys


def main(path, lo=0, hi=10 ** 9):
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for r in csv.reader(open(path)):
        if len(r) < 15 or not r[0].isdigit() or not (lo <= int(r[0]) <= hi):
            continue
        name = r[4].split('(')[0].replace('void ', '').replace('pvsr::', '')[:58]
        metric, unit, val = r[12], r[13], float(r[14].replace(',', ''))
        a = agg[name]
        if metric == 'gpu__time_duration.sum':
            a[0] += 1
            a[1] += val / 1e3 if unit == 'ns' else val * (1e3 if unit == 'ms' else 1.0)
        elif metric.startswith('dram__bytes_read'):
            a[2] += val * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(unit, 0)
        elif metric.startswith('dram__bytes_write'):
            a[3] += val * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(unit, 0)
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':60s} {'n':>5s} {'us':>11s} {'share':>6s} {'rd MB':>9s} {'wr MB':>9s} {'GB/s':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        bw = (v[2] + v[3]) / 1e3 / (v[1] / 1e6) if v[1] > 0 else 0
        print(f"{k:60s} {v[0]:5d} {v[1]:11.1f} {v[1] / tot:6.3f} {v[2]:9.1f} {v[3]:9.1f} {bw:7.0f}")
    print(f"{'total':60s} {sum(v[0] for v in agg.values()):5d} {tot:11.1f}")


if __name__ == '__main__':
    main(sys.argv[1], *(int(x) for x in sys.argv[2:4]))
