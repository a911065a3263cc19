"""Summarises an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of
scripts/profile_step.py: kernels of the LAST step, grouped by name."""
import collections
import csv
import sys

path, steps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr, rows = rows[0], rows[1:]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
n = len(rows) // steps
last = rows[-n:]
tot = sum(float(r[vi].replace(',', '')) for r in last) / 1000
agg = collections.OrderedDict()
for r in last:
  k = r[ki].split('(')[0].replace('void ', '')
  k = k if k.startswith('spml::') else 'ATen/cub glue kernels'
  a = agg.setdefault(k, [0, 0.0])
  a[0] += 1
  a[1] += float(r[vi].replace(',', '')) / 1000
print('%s: %d launches per step, %.1f us summed (serialised, cold cache)' % (path, n, tot))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
  print('  %-58s x%-3d %8.1f us  %5.1f %%' % (k[:58], c, t, 100 * t / tot))
