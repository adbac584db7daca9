"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel, launches,
total and share of the summed device time.  python tools/launch_summary.py launches.csv > summary.csv"""
import csv, collections, re, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors='replace')) if len(r) > 5]
hdr = next(r for r in rows if 'Kernel Name' in r)
ik, im, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot = collections.defaultdict(float); cnt = collections.Counter()
scale = {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 'nsecond': 1e-6, 's': 1e3, 'second': 1e3}
for r in rows:
    if r is hdr or r[im] != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', r[ik]).replace('void <unnamed>::', '')
    tot[name] += float(r[iv].replace(',', '')) * scale.get(r[iu], 1.0)
    cnt[name] += 1
total = sum(tot.values())
print('kernel,launches,total_ms,share')
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print('%s,%d,%.3f,%.4f' % (k, cnt[k], v, v / total))
