"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: python tools/launch_summary.py launches.csv [second_half]"""
import csv, sys, collections, re

def main():
    rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
    hdr = rows[0]
    ci = {n: i for i, n in enumerate(hdr)}
    data = []
    for r in rows[1:]:
        if len(r) < len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        grid = r[ci["Grid Size"]] if "Grid Size" in ci else ""
        data.append((re.sub(r"\(.*", "", r[ci["Kernel Name"]]), v, grid))
    if len(sys.argv) > 2:
        data = data[len(data) // 2:]
    agg = collections.OrderedDict()
    for n, v, g in data:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:70]:70s} n={a[0]:5d} total={a[1]:10.1f} us  share={100*a[1]/tot:5.1f}%  avg={a[1]/a[0]:8.1f} us")
    print(f"TOTAL {tot:.1f} us over {len(data)} launches")

main()
