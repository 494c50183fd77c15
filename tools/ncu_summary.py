"""Prints the key sections of an .ncu-rep (details page) and the top stall lines of the source page."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
h = r[0]; sec = h.index('Section Name'); mn = h.index('Metric Name'); mu = h.index('Metric Unit'); mv = h.index('Metric Value'); kn = h.index('Kernel Name')
want = ('GPU Speed Of Light Throughput', 'Memory Workload Analysis', 'Occupancy', 'Launch Statistics', 'Warp State Statistics',
        'Compute Workload Analysis', 'Scheduler Statistics')
skip = ('Cluster', 'Function Cache', 'Driver Shared', 'Stack Size', 'TPC', 'Green', 'Compression', 'Block Limit Barriers', 'Block Limit SM')
last = None
for row in r[1:]:
    if row[kn] != last:
        print("==", row[kn][:100]); last = row[kn]
    if row[sec] in want and not any(s in row[mn] for s in skip):
        print('  %-30s %-46s %-14s %s' % (row[sec][:30], row[mn], row[mu], row[mv]))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h = r[0]
for row in r[2:]:
    for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'smsp__inst_executed.sum',
              'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum', 'l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum',
              'smsp__inst_executed_op_global_red.sum'):
        if k in h:
            print('  raw %-60s %s %s' % (k, row[h.index(k)], r[1][h.index(k)]))
    for i, c in enumerate(h):
        if 'warp_issue_stalled' in c and c.endswith('per_warp_active.pct') and float(row[i] or 0) > 3:
            print('  stall %-70s %s' % (c, row[i]))
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(src)))
    hi = [i for i, x in enumerate(r) if 'Source' in x and any('Sampl' in y for y in x)]
    if hi:
        h = r[hi[0]]
        si = h.index('Source'); ci = [i for i, y in enumerate(h) if y.startswith('# Samples') or y == 'Warp Stall Sampling (All Samples)' or 'Sampling (All' in y]
        print(h)
        rows = []
        for x in r[hi[0] + 1:]:
            try:
                rows.append((float(x[ci[0]] or 0), x[si][:140], x[0]))
            except Exception:
                pass
        tot = sum(a for a, _, _ in rows) or 1
        for a, s_, ln in sorted(rows, reverse=True)[:int(sys.argv[2])]:
            print('%6.2f%%  L%-5s %s' % (100 * a / tot, ln, s_))
