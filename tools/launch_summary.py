"""Summarise an ncu launch list (gpu__time_duration.sum per launch, --csv) by kernel name."""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("ustrun::", "")
                v = float(d["Metric Value"].replace(",", ""))
                u = d["Metric Unit"]
                v = v / 1e3 if u in ("nsecond", "ns") else (v * 1e3 if u in ("msecond", "ms") else v)
                agg[name][0] += 1
                agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:56s} {v[0]:5d} {v[1]:10.1f} us {100 * v[1] / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
