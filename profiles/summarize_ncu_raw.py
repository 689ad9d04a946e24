#!/usr/bin/env python
"""`ncu -i X.ncu-rep --page raw --csv` -> one line per profiled launch with the columns the roofline needs.
    python profiles/summarize_ncu_raw.py gpurun_out/r2b_kernels_raw.csv > profiles/r02_ncu_kernels_summary.csv"""
import csv
import sys

COLS = [('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram_read'), ('dram__bytes_write.sum', 'dram_write'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct_of_peak'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_pipe_pct_active'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor_pipe_pct_elapsed'),
        ('lts__t_sectors.sum', 'l2_sectors'), ('lts__t_sector_hit_rate.pct', 'l2_hit_pct'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_pct'),
        ('launch__registers_per_thread', 'regs'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block')]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = csv.writer(sys.stdout)
    out.writerow(['id', 'kernel'] + ['%s [%s]' % (n, units[idx[k]]) if units[idx[k]] else n for k, n in COLS if k in idx])
    for r in rows[2:]:
        out.writerow([r[idx['ID']], r[idx['Kernel Name']][:70]] + [r[idx[k]] for k, _ in COLS if k in idx])


if __name__ == '__main__':
    main(sys.argv[1])
