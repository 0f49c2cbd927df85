"""Summarise an `ncu --csv` launch list (plain or .gz): per-kernel count, total/avg duration, DRAM bytes, GB/s."""
import collections
import csv
import gzip
import sys


def main(fn, skip_first=0):
    op = gzip.open if fn.endswith(".gz") else open
    rows = list(csv.reader(l for l in op(fn, "rt") if l.startswith('"')))
    hdr = rows[0]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = per.setdefault(int(r[ii]), {"name": r[ki]})
        try:
            d[r[mi]] = float(r[vi].replace(",", ""))
        except ValueError:
            pass
    agg = collections.OrderedDict()
    for i, d in per.items():
        if i < skip_first:
            continue
        nm = d["name"].split("(")[0].replace("void ", "").replace("omb::", "")[:58]
        a = agg.setdefault(nm, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0)
        a[3] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':58s} {'n':>5s} {'total us':>10s} {'avg us':>9s} {'share':>6s} {'rd MB':>9s} {'wr MB':>9s} {'GB/s':>7s}")
    for k, (c, t, rd, wr) in agg.items():
        gbs = (rd + wr) / t if t else 0.0
        print(f"{k:58s} {c:5d} {t/1e3:10.1f} {t/c/1e3:9.2f} {100*t/tot:5.1f}% {rd/1e6:9.1f} {wr/1e6:9.1f} {gbs:7.0f}")
    print(f"{'TOTAL':58s} {'':5s} {tot/1e3:10.1f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
