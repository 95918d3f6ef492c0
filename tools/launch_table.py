"""Markdown table (kernel, launches, ms, share) from an ncu `--metrics gpu__time_duration.sum --csv` log."""
import collections
import csv
import sys


def table(path):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[start:]))
    agg = collections.OrderedDict()
    for r in rows:
        k = r["Kernel Name"].split("(")[0].replace("void ", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e6
    tot = sum(v[1] for v in agg.values())
    out = ["| kernel | launches | ms | share |", "|---|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| %s | %d | %.2f | %.1f %% |" % (k, v[0], v[1], 100 * v[1] / tot))
    out.append("")
    out.append("total %.1f ms over %d launches" % (tot, len(rows)))
    return "\n".join(out)


if __name__ == "__main__":
    print(table(sys.argv[1]))
