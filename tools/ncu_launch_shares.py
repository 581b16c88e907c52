"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active... --csv` launch list of bench.py into per-kernel shares of ONE captured training step
(the launches between the last two optimizer kernels).  usage: ncu_launch_shares.py launches.csv out_prefix"""
import csv, collections, re, sys
src, prefix = sys.argv[1], sys.argv[2]
with open(src) as f:
    lines = [l for l in f if not l.startswith('==')]
data = collections.OrderedDict()
for x in csv.DictReader(lines):
    d = data.setdefault(int(x['ID']), {'name': x['Kernel Name']})
    v, u, m = float(x['Metric Value'].replace(',', '')), x['Metric Unit'], x['Metric Name']
    if m.startswith('dram__bytes'):
        v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
    if m == 'gpu__time_duration.sum':
        v *= {'ns': 1e-3, 'us': 1, 'ms': 1e3}[u]
    d[m] = v
ids = sorted(data)
opt = [i for i in ids if 'opt_adam' in data[i]['name']]
a, b = opt[-2], opt[-1]
step = [data[i] for i in ids if a < i <= b]
T, RD, WR, TP = 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'
tot = sum(d[T] for d in step)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
short = lambda n: re.sub(r'\(.*', '', n).replace('void ', '').replace('mpb::', '')
for d in step:
    g = agg[short(d['name'])]
    g[0] += 1; g[1] += d[T]; g[2] += d[RD]; g[3] += d[WR]; g[4] += d[TP] * d[T]
out = ["one captured training step (ncu: serialised, --clock-control none): %d launches, %.1f us total" % (len(step), tot),
       "%-52s %5s %10s %6s %10s %10s %8s" % ("kernel", "n", "time_us", "share", "dram_rd_MB", "dram_wr_MB", "tensor%")]
for n, g in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%-52s %5d %10.1f %5.1f%% %10.1f %10.1f %8.1f" % (n[:52], g[0], g[1], 100 * g[1] / tot, g[2] / 1e6, g[3] / 1e6, g[4] / g[1] if g[1] else 0))
gem = [d for d in step if 'tc_gemm' in d['name']]
gt, rd, wr = sum(d[T] for d in gem), sum(d[RD] for d in gem), sum(d[WR] for d in gem)
out.append("tcgen05 GEMM family: %d launches, %.1f us (%.1f%% of the step), DRAM read %.1f MB + write %.1f MB = %.1f MB per step, %.3f MB per launch" % (
    len(gem), gt, 100 * gt / tot, rd / 1e6, wr / 1e6, (rd + wr) / 1e6, (rd + wr) / 1e6 / len(gem)))
out.append("whole step DRAM traffic: read %.1f MB + write %.1f MB" % (sum(d[RD] for d in step) / 1e6, sum(d[WR] for d in step) / 1e6))
open(prefix + "_launch_shares.txt", "w").write("\n".join(out) + "\n")
with open(prefix + "_launches_step.csv", "w") as f:
    f.write("kernel,time_us,dram_read_bytes,dram_write_bytes,tensor_pipe_pct\n")
    for d in step:
        f.write("%s,%.3f,%d,%d,%.2f\n" % (short(d['name']).replace(',', ';'), d[T], d[RD], d[WR], d[TP]))
print("\n".join(out))
