"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: total ms, launches, share of the step."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) < len(rows[hdr]):
        continue
    d = dict(zip(rows[hdr], r))
    v = float(d['Metric Value'].replace(',', ''))
    v *= {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0}[d['Metric Unit']]
    name = re.sub(r'\(.*', '', d['Kernel Name']).replace('void ', '').replace('fgc::', '')
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
print('total %.2f ms in %d launches' % (tot, sum(v[0] for v in agg.values())))
fam = collections.defaultdict(float)
for k, v in agg.items():
    f = ('tcgen05 conv' if re.match(r'conv_(halo|igemm|wgrad)', k) else 'direct narrow conv' if k.startswith('conv_small') else
         'patch / weight pack' if re.match(r'im2col|pack_weights', k) else 'caption LSTM' if re.match(r'lstm|embedding|l2norm|rows_group|atanh', k)
         else 'spectral norm' if k.startswith('sn_') else 'streaming / reduction passes')
    fam[f] += v[1]
for f, ms in sorted(fam.items(), key=lambda kv: -kv[1]):
    print('  %-30s %7.2f ms  %5.1f%%' % (f, ms, 100 * ms / tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print('%8.3f ms %5d  %5.1f%%  %s' % (v[1], v[0], 100 * v[1] / tot, k[:100]))
