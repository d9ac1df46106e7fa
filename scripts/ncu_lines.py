"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump by CUDA
source line: share of warp-stall samples and of executed instructions."""
import csv, sys

def main(path, top=45):
    rows = list(csv.reader(open(path)))
    cur = None
    agg = []
    for r in rows:
        if len(r) >= 2 and r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if len(r) > 7 and r[0].isdigit() and r[2] == '-':
            try:
                agg.append((cur, int(r[0]), r[1].strip(), int(r[4]), int(r[7])))
            except ValueError:
                pass
    tot_s = sum(a[3] for a in agg) or 1
    tot_i = sum(a[4] for a in agg) or 1
    print("total samples", tot_s, "total warp inst", tot_i)
    byfile = {}
    for a in agg:
        f = byfile.setdefault(a[0], [0, 0]); f[0] += a[3]; f[1] += a[4]
    for k, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print("  file %-28s samp %5.1f%% inst %5.1f%%" % (k, 100 * v[0] / tot_s, 100 * v[1] / tot_i))
    agg.sort(key=lambda a: -a[3])
    for a in agg[:top]:
        print("%-18s %4d  samp %5.1f%%  inst %5.1f%%  %s" % (a[0], a[1], 100 * a[3] / tot_s, 100 * a[4] / tot_i, a[2][:100]))

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
