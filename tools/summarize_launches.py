"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
tot, cnt = collections.defaultdict(float), collections.Counter()
for row in csv.DictReader(lines):
    try:
        v = float(row['Metric Value'].replace(',', ''))
    except Exception:
        continue
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(row['Metric Unit'], 1.0)
    name = re.sub(r'\(.*', '', row['Kernel Name'])[:80]
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print('# %s: %d launches, %.1f ms summed device time (cold-cache, serialised)' % (path, sum(cnt.values()), T / 1e3))
print('%-82s %10s %6s %7s %9s' % ('kernel', 'us', 'share', 'n', 'avg us'))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print('%-82s %10.0f %5.1f%% %7d %9.1f' % (k, v, 100 * v / T, cnt[k], v / cnt[k]))
