"""Summarises an `ncu --csv` launch list (gpu__time_duration.sum [+ dram bytes]) per kernel name.
Usage: python tools/summarize_launches.py file.csv [--md]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[h]
kn, mn, mv, idc = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
per = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) <= mv:
        continue
    per.setdefault(r[idc], {'name': r[kn]})[r[mn]] = float(r[mv].replace(',', ''))
agg = collections.OrderedDict()
for d in per.values():
    name = d['name'].split('(')[0].replace('void ', '').replace('ib::', '')
    name = name[:70]
    a = agg.setdefault(name, {'n': 0, 't': 0.0, 'rd': 0.0, 'wr': 0.0})
    a['n'] += 1
    a['t'] += d.get('gpu__time_duration.sum', 0.0)
    a['rd'] += d.get('dram__bytes_read.sum', 0.0)
    a['wr'] += d.get('dram__bytes_write.sum', 0.0)
tot = sum(a['t'] for a in agg.values())
md = '--md' in sys.argv
print(('| kernel | launches | total us | share | DRAM MB (rd+wr) | DRAM GB/s |\n|---|---|---|---|---|---|') if md else 'kernel  n  total_us  share  dram_MB  GB/s')
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]['t']):
    mb = (a['rd'] + a['wr']) / 1e6
    gbs = (a['rd'] + a['wr']) / a['t'] if a['t'] else 0.0
    fmt = '| `%s` | %d | %.1f | %.1f %% | %.1f | %.0f |' if md else '%-72s %4d %9.1f %5.1f%% %9.1f %7.0f'
    print(fmt % (name, a['n'], a['t'] / 1e3, 100 * a['t'] / tot, mb, gbs))
print(('| **total** | %d | %.1f | | | |' if md else 'TOTAL %d launches %.1f us') % (sum(a['n'] for a in agg.values()), tot / 1e3))
