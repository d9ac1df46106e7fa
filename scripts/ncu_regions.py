"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump by source line: share of
warp-stall samples, executed instructions and the stall-reason mix; also the kernel's code size.
usage: ncu_regions.py <source.csv> [top N lines] [frames in launch]"""
import csv, sys, collections

def load(path):
    rows = list(csv.reader(open(path)))
    cur = None; hdr = None; out = []; addrs = []
    for r in rows:
        if len(r) >= 2 and r[0] == 'File Path':
            cur = r[1].split('/')[-1]; continue
        if len(r) > 20 and r[0] == 'Line No':
            hdr = r; continue
        if len(r) > 20 and hdr:
            if r[0].isdigit() and r[2] == '-':
                d = {}
                for k, v in zip(hdr[4:], r[4:]):
                    try: d[k] = float(v)
                    except ValueError: d[k] = 0.0
                out.append((cur, int(r[0]), r[1].strip(), d))
            elif r[2].startswith('0x'):
                addrs.append(int(r[2], 16))
    return out, addrs

def main():
    path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    frames = float(sys.argv[3]) if len(sys.argv) > 3 else None
    lines, addrs = load(path)
    tot_s = sum(l[3]['# Samples'] for l in lines) or 1
    tot_i = sum(l[3]['Instructions Executed'] for l in lines) or 1
    if addrs:
        print("code size: %.1f KB (%d distinct instruction addresses seen)" % ((max(addrs) - min(addrs) + 16) / 1024., len(set(addrs))))
    print("total samples %d, warp instructions %d%s" % (tot_s, tot_i, (" = %.0f per frame" % (tot_i / frames)) if frames else ""))
    reasons = [k for k in lines[0][3] if k.startswith('stall_') and 'Not Issued' not in k]
    tot_r = {k: sum(l[3][k] for l in lines) for k in reasons}
    print("stall mix: " + "  ".join("%s %.1f%%" % (k[6:], 100 * v / tot_s) for k, v in sorted(tot_r.items(), key=lambda kv: -kv[1]) if v / tot_s > 0.01))
    byfile = collections.defaultdict(lambda: [0, 0])
    for f, ln, src, d in lines:
        byfile[f][0] += d['# Samples']; byfile[f][1] += d['Instructions Executed']
    for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print("  file %-26s samp %5.1f%% inst %5.1f%%" % (f, 100 * v[0] / tot_s, 100 * v[1] / tot_i))
    lines.sort(key=lambda l: -l[3]['# Samples'])
    for f, ln, src, d in lines[:top]:
        mix = sorted(((k[6:], d[k]) for k in reasons), key=lambda kv: -kv[1])[:3]
        print("%-16s %4d samp %5.2f%% inst %5.2f%% [%s] %s" % (f, ln, 100 * d['# Samples'] / tot_s, 100 * d['Instructions Executed'] / tot_i,
              " ".join("%s %.0f%%" % (k, 100 * v / max(d['# Samples'], 1)) for k, v in mix), src[:70]))

if __name__ == "__main__":
    main()
