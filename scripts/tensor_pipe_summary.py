"""Reduce an ncu per-launch list with gpu__time_duration.sum and sm__pipe_tensor_cycles_active...pct_of_peak_sustained_elapsed
(scripts/g_conv_stack.py) to the time-weighted tensor-pipe utilisation of the convolution kernels, per kernel family and overall.
usage: python scripts/tensor_pipe_summary.py gpurun_out/g_stack.csv > profiles/<tag>_g_conv_tensor_pipe.txt"""
import collections
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
per = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) < len(rows[hdr]):
        continue
    d = dict(zip(rows[hdr], r))
    e = per.setdefault(d['ID'], {'name': re.sub(r'\(.*', '', d['Kernel Name']).replace('void ', '').replace('fgc::', '')})
    v = float(d['Metric Value'].replace(',', ''))
    if d['Metric Name'].startswith('gpu__time_duration'):
        e['ms'] = v * {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0}[d['Metric Unit']]
    else:
        e['pct'] = v
fam = collections.defaultdict(lambda: [0, 0.0, 0.0])
for e in per.values():
    if 'ms' not in e or 'pct' not in e:
        continue
    f = fam[e['name']]
    f[0] += 1
    f[1] += e['ms']
    f[2] += e['ms'] * e['pct']
tot = sum(f[1] for f in fam.values())
conv = {k: f for k, f in fam.items() if re.match(r'conv_', k)}
tc = {k: f for k, f in conv.items() if re.match(r'conv_(halo|igemm|wgrad)', k)}
t_conv, w_conv = sum(f[1] for f in conv.values()), sum(f[2] for f in conv.values())
t_tc, w_tc = sum(f[1] for f in tc.values()), sum(f[2] for f in tc.values())
print('# generator forward + backward, bs 64, 192x192, bf16: %d launches, %.2f ms of kernel time (ncu, serialised)' % (
    sum(f[0] for f in fam.values()), tot))
print('# time-weighted sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed')
print('all convolution kernels (tcgen05 + direct narrow): %.2f ms, tensor pipe %.1f %%' % (t_conv, w_conv / max(t_conv, 1e-9)))
print('tcgen05 convolution kernels only:                  %.2f ms, tensor pipe %.1f %%' % (t_tc, w_tc / max(t_tc, 1e-9)))
print('whole generator pass (every kernel):               %.2f ms, tensor pipe %.1f %%' % (tot, sum(f[2] for f in fam.values()) / max(tot, 1e-9)))
print()
for k, f in sorted(conv.items(), key=lambda kv: -kv[1][1]):
    print('%8.3f ms %4d launches  tensor pipe %5.1f %%  %s' % (f[1], f[0], f[2] / f[1], k))
print(json.dumps({"g_conv_tensor_pipe_pct": round(w_conv / max(t_conv, 1e-9), 2), "g_tcgen05_conv_tensor_pipe_pct": round(w_tc / max(t_tc, 1e-9), 2),
                  "g_conv_ms": round(t_conv, 3), "g_pass_ms": round(tot, 3)}))
