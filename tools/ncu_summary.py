#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full capture) into a small CSV under profiles/.

usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_ncu_summary.csv
Per captured kernel launch: duration, DRAM bytes, pipe utilisations, issue rate, shared-memory
wavefronts / bank conflicts, registers.  (The .ncu-rep itself is too large to commit.)"""
import csv
import subprocess
import sys

KEYS = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(KEYS)
        w.writerow([units[hdr.index(k)] if k in hdr else "" for k in KEYS])
        for r in rows[2:]:
            w.writerow([r[hdr.index(k)] if k in hdr else "" for k in KEYS])
    print(open(out).read())


if __name__ == "__main__":
    main()
