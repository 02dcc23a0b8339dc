"""Time-weighted tensor-pipe utilisation of one Euler update from an ncu CSV launch list
(ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --csv ... tools/prof_forward.py):
per kernel class the launches, total time, share of the update and mean tensor-pipe %, and the whole-update figure
sum(t_i * tensor_i) / sum(t_i) -- the north_star's "tensor-pipe utilisation in the transformer forward"."""
import csv, re, sys
from collections import defaultdict

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
iname, imet, ival, iid = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('ID')
per = defaultdict(dict)
for r in rows[1:]:
    try:
        per[(r[iid], r[iname])][r[imet]] = float(r[ival].replace(',', ''))
    except ValueError:
        pass
cls = defaultdict(lambda: [0, 0.0, 0.0])
tot_t = tot_w = 0.0
for (kid, name), m in per.items():
    t = m.get('gpu__time_duration.sum', 0.0)
    u = m.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0.0)
    short = re.sub(r'^void (e2b::)?', '', name).split('(')[0]
    c = cls[short]
    c[0] += 1; c[1] += t; c[2] += t * u
    tot_t += t; tot_w += t * u
unit = 1e6 if tot_t > 1e6 else 1e3          # ncu prints ns (or us)
print(f'{"kernel":58s} {"launches":>8s} {"time ms":>9s} {"share":>7s} {"tensor %":>9s}')
for k, (n, t, w) in sorted(cls.items(), key=lambda kv: -kv[1][1]):
    print(f'{k[:58]:58s} {n:8d} {t / unit:9.2f} {100 * t / tot_t:6.1f}% {w / t if t else 0:8.1f}%')
print(f'{"whole update (cold-cache, serialised launches)":58s} {sum(c[0] for c in cls.values()):8d} {tot_t / unit:9.2f} {100.0:6.1f}% {tot_w / tot_t:8.1f}%')
