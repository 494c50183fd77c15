"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total and share of the last GN step."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 5]
h = rows[0]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
seq = [(r[kn].split("(")[0], float(r[mv].replace(",", "")) * (1e-3 if r[mu] in ("ns", "nsecond") else 1.0)) for r in rows[1:] if r[mv]]
# the last GN step: from the last k_hb_init on
start = max((i for i, (k, _) in enumerate(seq) if "k_hb_init" in k), default=0)
step = seq[start:]
tot = sum(t for _, t in step)
agg = collections.OrderedDict()
for k, t in step:
    c, s = agg.get(k, (0, 0.0))
    agg[k] = (c + 1, s + t)
print("one GN step: %d launches, %.1f us under ncu (cold-cache, serialised)" % (len(step), tot))
for k, (c, s) in agg.items():
    print("%-70s %4d %10.1f us %6.2f%%" % (k[:70], c, s, 100 * s / tot))
