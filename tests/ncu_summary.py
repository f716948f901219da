"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total ms, share."""
import collections
import csv
import sys


def summarise(path):
    rows = list(csv.reader(open(path)))
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if hdr is None:
            if len(r) > 5 and r[0] == "ID":
                hdr = r
            continue
        if len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = d["Kernel Name"].split("(")[0]
        v = float(d["Metric Value"].replace(",", ""))
        unit = d["Metric Unit"]
        if unit == "ns":
            v /= 1e6
        elif unit == "us":
            v /= 1e3
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values()) or 1.0
    out = ["kernel,launches,total_ms,share"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%s,%d,%.3f,%.3f" % (k, v[0], v[1], v[1] / tot))
    return out


if __name__ == "__main__":
    print("\n".join(summarise(sys.argv[1])))
