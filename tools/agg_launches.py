"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv, collections, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
agg = collections.defaultdict(lambda: [0, 0.0]); n = 0
for row in csv.DictReader(lines):
    try: v = float(row['Metric Value'].replace(',', ''))
    except Exception: continue
    u = row['Metric Unit']
    v = v / 1000 if u == 'ns' else (v * 1000 if u == 'ms' else v)
    k = row['Kernel Name'][:70]; agg[k][0] += 1; agg[k][1] += v; n += 1
tot = sum(v[1] for v in agg.values())
print(n, 'launches, total', round(tot), 'us')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{v[1]:10.0f} us {100*v[1]/tot:5.1f}% {v[0]:5d}  {k}")
