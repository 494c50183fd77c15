"""Turns the raw measurement files of tools/collect_profiles.sh (gpurun_out/) into the committed summaries under profiles/:
launch list (per-kernel totals and share of a step), ncu section summaries and top stalled SASS of the captured kernels,
profiles/traffic.json (DRAM bytes per unit, read by bench.py), and copies of the bench lines."""
import csv, io, json, os, subprocess, sys, collections, shutil

R = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[0]
    return [dict(zip(h, r)) for r in rows[2:]], dict(zip(h, rows[1]))

def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]

# ---- launch list -----------------------------------------------------------------------------------------------------------
f = os.path.join(G, "launches_%s.csv" % R)
if os.path.exists(f):
    rows = list(csv.reader(open(f)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]; kn = h.index("Kernel Name"); mv = h.index("Metric Value")
    seq = [(r[kn], float(r[mv].replace(",", ""))) for r in rows[hi + 1:] if len(r) > mv]
    # one GN step = from one k_landmark_init to the next
    idx = [i for i, (n, _) in enumerate(seq) if "k_landmark_init" in n]
    step = seq[idx[-2]:idx[-1]] if len(idx) > 1 else seq
    tot = sum(v for _, v in step)
    agg = collections.OrderedDict()
    for n, v in step:
        key = n.split("(")[0].replace("void ", "")
        a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += v
    with open(os.path.join(P, "launches_%s.txt" % R), "w") as o:
        o.write("ncu --metrics gpu__time_duration.sum --clock-control none, command: python bench.py --steps 2 --warmup 3 --no-cpu-baseline\n")
        o.write("per-launch times are cold-cache and serialised: the SHARE of a step is the comparable figure\n")
        o.write("one GN step (k_landmark_init .. k_update), %d launches, %.1f us under ncu\n\n" % (len(step), tot / 1e3))
        o.write("%-60s %6s %12s %8s\n" % ("kernel", "count", "total us", "share"))
        for k, (c, v) in agg.items():
            o.write("%-60s %6d %12.1f %7.2f%%\n" % (k[:60], c, v / 1e3, 100 * v / tot))
    shutil.copy(f, os.path.join(P, "launches_%s.csv" % R))

# ---- full captures ---------------------------------------------------------------------------------------------------------
traffic = {}
for tag in ("pcg", "lin", "syrk"):
    rep = os.path.join(G, "prof_%s_%s.ncu-rep" % (tag, R))
    if not os.path.exists(rep):
        continue
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    sass = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_sass_top.py"), rep, "25"], capture_output=True, text=True).stdout
    with open(os.path.join(P, "ncu_%s_%s.txt" % (tag, R)), "w") as o:
        o.write("ncu --set full --import-source on --clock-control none (see tools/collect_profiles.sh), report %s\n\n" % os.path.basename(rep))
        o.write(summ + "\n---- top stalled SASS of the first kernel in the report ----\n" + sass)
    ms, units = raw_metrics(rep)
    for m in ms:
        name = m["Kernel Name"]
        rd = to_bytes(m["dram__bytes_read.sum"], units["dram__bytes_read.sum"]); wr = to_bytes(m["dram__bytes_write.sum"], units["dram__bytes_write.sum"])
        dur = float(m["gpu__time_duration.sum"].replace(",", ""))
        traffic.setdefault("kernels", {})[name.split("(")[0].replace("void ", "")] = {"dram_read": rd, "dram_write": wr, "duration": dur,
                                                                                     "duration_unit": units["gpu__time_duration.sum"]}
b = os.path.join(G, "bench_%s.json" % R)
if os.path.exists(b):
    line = json.load(open(b))
    shutil.copy(b, os.path.join(P, "bench_%s.json" % R))
    ks = traffic.get("kernels", {})
    pcg = [v for k, v in ks.items() if "k_pcg_fused" in k and "prep" not in k]
    if pcg and line.get("pcg_iterations"):
        # the captured launch ran the same solve as the bench step: divide by its CG iterations
        traffic["pcg_dram_bytes_per_cg_iteration"] = (pcg[0]["dram_read"] + pcg[0]["dram_write"]) / line["pcg_iterations"]
    hb = [v for k, v in ks.items() if "k_linearize_bearing" in k or "k_linearize_odometry" in k]   # both kernels of the H, b build
    if hb:
        traffic["hb_build_dram_bytes"] = sum(v["dram_read"] + v["dram_write"] for v in hb)
for extra in ("bench_f32_%s.json", "bench_bj_%s.json", "bench_100k_%s.json", "batch_%s.log", "dense_%s.log", "bench_n2_%s.json", "bench_n4_%s.json", "bench_n8_%s.json"):
    e = os.path.join(G, extra % R)
    if os.path.exists(e) and os.path.getsize(e) > 0:
        shutil.copy(e, os.path.join(P, extra % R))
r = os.path.join(G, "bench_ref_%s.json" % R)
if os.path.exists(r):
    shutil.copy(r, os.path.join(P, "bench_ref_%s.json" % R))
if traffic:
    traffic["source"] = "ncu --set full captures of round %s (profiles/ncu_*_%s.txt)" % (R, R)
    json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(open(os.path.join(P, "launches_%s.txt" % R)).read() if os.path.exists(os.path.join(P, "launches_%s.txt" % R)) else "no launch list")
print(json.dumps(traffic, indent=1)[:1500])
