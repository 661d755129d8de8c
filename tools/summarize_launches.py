"""Aggregate an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`)
per kernel: python tools/summarize_launches.py X.csv [first_id last_id]  ->  markdown table on stdout."""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) >= 15]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, im, iv, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    agg = collections.OrderedDict()
    for r in rows:
        if r is hdr or r[im] != "gpu__time_duration.sum" or not r[ii].isdigit() or not lo <= int(r[ii]) <= hi:
            continue
        name = r[ik].split("(")[0].replace("void ", "")[:56]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.3f | %.1f | %.1f%% |" % (k, v[0], v[1] / 1e6, v[1] / v[0] / 1e3, 100 * v[1] / tot))
    print("\nTotal %.3f ms over %d launches" % (tot / 1e6, sum(v[0] for v in agg.values())))


if __name__ == "__main__":
    main()
