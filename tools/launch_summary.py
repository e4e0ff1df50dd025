"""Per-kernel averages of a single-pass ncu launch list (csv): duration, DRAM bytes, instructions, registers."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ii = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
per = collections.OrderedDict()
for r in rows[1:]:
    d = per.setdefault(r[ii], {'k': r[ki]})
    d[r[mi]] = float(r[vi].replace(',', ''))
agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault(d['k'].replace('<unnamed>::', '')[:70], collections.Counter())
    a['n'] += 1
    for m in ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum'):
        a[m] += d.get(m, 0.0)
    a['regs'] = d.get('launch__registers_per_thread', 0)
tot = 0.0
for k, a in agg.items():
    n = a['n']
    tot += a['gpu__time_duration.sum'] / n / 1e3
    print(f"{k:70s} n={n:3d} avg {a['gpu__time_duration.sum'] / n / 1e3:8.1f} us  rd {a['dram__bytes_read.sum'] / n / 1e6:7.1f} MB "
          f"wr {a['dram__bytes_write.sum'] / n / 1e6:7.1f} MB  inst {a['smsp__inst_executed.sum'] / n / 1e6:6.2f} M  regs {a['regs']:.0f}")
print(f"sum of averages {tot:.1f} us")
