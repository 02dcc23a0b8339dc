import json, sys
d = json.load(open(sys.argv[1]))
rows = sorted(d['rows'], key=lambda r: -r['ms'])
tot = sum(r['ms'] for r in rows)
print(f"total {tot:.1f} ms")
kinds = {}
for r in rows:
    kinds.setdefault(r['kind'], 0.0)
    kinds[r['kind']] += r['ms']
print('  '.join(f"{k}={v:.1f}" for k, v in sorted(kinds.items(), key=lambda kv: -kv[1])))
for r in rows[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    per = r['ms'] / r['count']
    tf = r['flops'] / (per * 1e-3) / 1e12 if r['flops'] else 0
    gb = r['bytes'] / (per * 1e-3) / 1e9
    print(f"{r['kind']:12s} M={r['m']:7d} N={r['n']:6d} K={r['k']:5d} x{r['count']:3d} {r['ms']:8.2f} ms ({100*r['ms']/tot:4.1f}%) {per*1e3:8.1f} us {tf:7.1f} TF/s {gb:7.1f} GB/s")
