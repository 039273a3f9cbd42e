"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv, collections, sys
for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split('(')[0][:60]; v = float(r[vi].replace(',', ''))
        if r[ui] == 'ns': v /= 1000
        elif r[ui] == 'ms': v *= 1000
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(path)
    for k, (c, t) in agg.items(): print("  %-60s %5d %10.1f us  avg %7.1f  %5.1f%%" % (k, c, t, t / c, 100 * t / tot))
    print("  total %.1f us" % tot)
