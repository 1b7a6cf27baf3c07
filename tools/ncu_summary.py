"""Condense an `ncu --page raw --csv` dump to the columns the roofline needs."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = [('Kernel Name', 'kernel', 44), ('launch__grid_size', 'grid', 7), ('gpu__time_duration.sum', 'ms', 9),
        ('dram__bytes_read.sum', 'rdMB', 9), ('dram__bytes_write.sum', 'wrMB', 9),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%', 6),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%', 6),
        ('l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1%', 6),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%', 6),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%', 6),
        ('launch__registers_per_thread', 'regs', 5), ('smsp__inst_executed.sum', 'winst', 11),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'bankconf', 9)]
idx = [(hdr.index(w), n, wd) if w in hdr else (None, n, wd) for w, n, wd in want]
print(' '.join(n.ljust(wd) for _, n, wd in idx))
for r in rows[2:]:
    out = []
    for i, n, wd in idx:
        v = r[i] if i is not None else 'NA'
        try:
            v = f"{float(v.replace(',', '')):.4g}"
        except ValueError:
            pass
        out.append(v[:wd].ljust(wd))
    print(' '.join(out))
