"""Stall reasons per issued instruction + headline pipe numbers from an ncu raw csv."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
out = []
for h, u, v in zip(hdr, units, vals):
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
        try:
            out.append((float(v.replace(',', '')), h.split('issue_stalled_')[1].split('_per_issue')[0]))
        except ValueError:
            pass
for x, n in sorted(out, reverse=True)[:10]:
    print('  stall %-28s %6.2f' % (n, x))
want = ['gpu__time_duration.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_elapsed.avg', 'dram__bytes_read.sum', 'dram__bytes_write.sum']
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print('  %-70s %s %s' % (h, v, u))
