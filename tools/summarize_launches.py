import csv, collections, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
r = csv.DictReader(lines)
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0
for row in r:
    name = row['Kernel Name']; v = float(row['Metric Value'].replace(',', '')); unit = row['Metric Unit']
    if unit == 'ns': v /= 1000
    elif unit == 'ms': v *= 1000
    name = re.sub(r'\(.*', '', name).replace('void ', '').replace('<unnamed>::', '').replace('gg::', '')
    agg[name][0] += 1; agg[name][1] += v; tot += v
print("total %.1f us over %d launches (ncu per-launch times: cold cache, serialised)" % (tot, sum(a[0] for a in agg.values())))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("%8.1f us %5.1f%% n=%3d avg=%7.1f  %s" % (t, 100 * t / tot, n, t / n, k[:80]))
