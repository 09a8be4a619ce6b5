"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: count, total, share."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
agg = collections.OrderedDict()
for r in data:
    name = r[ix["Kernel Name"]]
    short = re.sub(r"\(.*", "", re.sub(r"<.*", "", name))[:70]
    if name.startswith("void rb::") or name.startswith("rb::"):
        short = re.sub(r"\(.*", "", name)[:70]
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += float(r[ix["Metric Value"]].replace(",", ""))
tot = sum(v for _, v in agg.values())
print(f"# {sys.argv[1]}: {len(data)} launches, {tot / 1e6:.3f} ms of kernel time (ncu-serialised, cold cache: compare shares)")
print(f"{'kernel':72s} {'n':>6s} {'total_us':>12s} {'share':>7s}")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{k:72s} {c:6d} {v / 1e3:12.1f} {100 * v / tot:6.1f}%")
