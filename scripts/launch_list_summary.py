"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', '')); unit = row['Metric Unit']
    v = v / 1e3 if unit == 'ns' else v * 1e3 if unit == 'ms' else v * 1e6 if unit in ('s', 'second') else v
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    name = re.sub(r'rt_foreach_kernel<.*?(rtcore::)?(\w+)\(.*', r'rt_foreach_kernel<\2 lambda>', name)[:80]
    a = agg.setdefault(name, [0, 0.0, []]); a[0] += 1; a[1] += v; a[2].append(round(v, 1))
tot = sum(a[1] for a in agg.values())
print(f"# {sys.argv[1]}: {sum(a[0] for a in agg.values())} launches, {tot:.1f} us total (ncu: cold-cache, serialised; compare shares)")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:80s} n={a[0]:4d} total={a[1]:10.1f} us share={100 * a[1] / tot:5.1f}%  first={a[2][:9]}")
