"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one forward pass, per launch.
usage: python scripts/launch_summary.py gpurun_out/launches.csv [which_forward]"""
import csv, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = [r for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
names = [r['Kernel Name'] for r in rows]
idx = [i for i, n in enumerate(names) if 'mel_to_act' in n]
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1
s, e = idx[w], (idx[w + 1] if w + 1 < len(idx) else len(rows))
fw = rows[s:e]
tot = sum(float(r['Metric Value']) for r in fw)
print('forward #%d: %d launches, %.1f us total (serialised, cold-cache: compare shares)' % (w, len(fw), tot / 1e3))
by = {}
for ci, r in enumerate(fw):
    n = r['Kernel Name'].split('(')[0].split('::')[-1][:28]
    t = float(r['Metric Value']) / 1e3
    by[n] = by.get(n, 0) + t
    print('%3d %-28s grid=%-18s %8.1f us' % (ci, n, r['Grid Size'], t))
for n, t in by.items():
    print('%-28s %9.1f us  %5.1f%%' % (n, t, 100 * t * 1e3 / tot))
