"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list (usage: launch_summary.py in.csv
[header line ...]): launches, mean duration and share of one step (kernels keyed by name + grid)."""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
acc = OrderedDict()
for r in rows:
    name = r[4].replace("void ", "").split("(")[0]
    key = "%s grid %s" % (name, r[8].replace(", ", "x")) if "anchor_hidden_tc2" in name else name
    ns = float(r[14].replace(",", ""))
    a = acc.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += ns
total = sum(v[1] / v[0] for v in acc.values())   # every kernel listed runs once per step
for line in sys.argv[2:]:
    print("# " + line)
print("kernel,launches,mean_us,share_of_step")
for k, (n, ns) in acc.items():
    print("%s,%d,%.1f,%.3f" % (k, n, ns / n / 1e3, ns / n / total))
