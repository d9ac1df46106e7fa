"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import csv, collections, sys

def main(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split('(')[0].split('::')[-1]
        v = float(r[vi].replace(',', ''))
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}[r[ui]]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print('total ms %.2f' % tot)
    for k, a in agg.items():
        print('  %-28s n=%3d  %9.2f ms  %5.1f%%' % (k, a[0], a[1], 100 * a[1] / tot))

if __name__ == "__main__":
    main(sys.argv[1])
