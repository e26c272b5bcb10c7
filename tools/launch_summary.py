"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel for ONE step of the bench
(the launches between two consecutive occurrences of an anchor kernel, default xproj)."""
import collections
import csv
import re
import sys


def main(path, anchor="xproj", which=1):
    lines = [l for l in open(path) if not l.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        seq.append((row["Kernel Name"], v / 1000.0 if row["Metric Unit"] == "ns" else v))
    idx = [i for i, (n, _) in enumerate(seq) if anchor in n]
    a, b = (idx[which], idx[which + 1]) if len(idx) > which + 1 else (idx[-1], len(seq))
    step = seq[a:b]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, us in step:
        k = re.sub(r"\(.*", "", n)[:64]
        agg[k][0] += 1
        agg[k][1] += us
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-66s n=%4d total %9.1f us  avg %8.1f  %5.1f%%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
    print("total %.1f us in %d launches" % (tot, len(step)))


if __name__ == "__main__":
    main(sys.argv[1], *(sys.argv[2:3] or ["xproj"]), *([int(sys.argv[3])] if len(sys.argv) > 3 else []))
