"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time share of
one fused training step (cold-cache, serialised: compare SHARES, not absolutes)."""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        try:
            rows.append((int(r["ID"]), r["Kernel Name"], float(r["Metric Value"].replace(",", ""))))
        except Exception:
            pass
    return rows


def short(name):
    n = name.replace("<unnamed>::", "").replace("(anonymous namespace)::", "").replace("ssnb::", "").replace("void ", "")
    m = re.match(r"([\w:]+)(<[^(]*>)?\(", n)
    if not m:
        return n[:50]
    base = m.group(1)
    tmpl = m.group(2) or ""
    tag = ""
    if "__half" in tmpl:
        tag = "<half>"
    elif "float" in tmpl and "native" not in base:
        tag = "<float>"
    if "(bool)1" in tmpl:
        tag += "[flat]"
    return base.split("::")[-1] + tag


def main():
    rows = load(sys.argv[1])
    k = int(sys.argv[2]) if len(sys.argv) > 2 else -1
    marks = [i for i, r in enumerate(rows) if "gpool_stpp" in r[1]]
    starts = [i for i, r in enumerate(rows) if "nchw_to_nhwc" in r[1] or "nchw_to_s2d" in r[1]]
    m = marks[k]
    s = max(i for i in starts if i < m)
    later = [i for i in starts if i > m]
    e = later[0] if later else len(rows)
    agg, cnt, tot = collections.OrderedDict(), collections.Counter(), 0.0
    for r in rows[s:e]:
        key = short(r[1])
        agg[key] = agg.get(key, 0) + r[2]
        cnt[key] += 1
        tot += r[2]
    print("launch list %s: fused step #%d = launches [%d, %d) : %d launches, %.3f ms summed kernel time" % (sys.argv[1], k, s, e, e - s, tot / 1e6))
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        print("%-44s n=%4d  %9.3f ms  %5.1f%%" % (key[:44], cnt[key], v / 1e6, 100 * v / tot))


if __name__ == "__main__":
    main()
