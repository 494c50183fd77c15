"""Top stalled SASS instructions of a kernel in an .ncu-rep (captured with --set full --import-source on)."""
import csv, subprocess, sys, io
rep, topn = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
si = h.index("Source"); ai = h.index("# Samples")
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data = []
for n, r in enumerate(rows[2:]):
    try:
        data.append((float(r[ai] or 0), n, r))
    except Exception:
        pass
tot = sum(d[0] for d in data) or 1
print("total samples", tot, "instructions", len(data))
# aggregate stall reasons
agg = {}
for a, n, r in data:
    for i, c in stall_cols:
        try:
            agg[c] = agg.get(c, 0) + float(r[i] or 0)
        except Exception:
            pass
print("stall mix:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / tot) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for a, n, r in sorted(data, reverse=True)[:topn]:
    reasons = sorted(((float(r[i] or 0), c[6:]) for i, c in stall_cols), reverse=True)[:2]
    print("%5.2f%%  #%-5d %-70s %s" % (100 * a / tot, n, r[si].strip()[:70], " ".join("%s=%d" % (c, v) for v, c in reasons if v > 0)))
