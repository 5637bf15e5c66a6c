#!/usr/bin/env python
"""Condense an `ncu --set full` report into the per-launch columns kept under profiles/.

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > /tmp/x.csv
    python tools/summarize_ncu_full.py /tmp/x.csv > profiles/r1X_ncu_full_*.csv
"""
import csv
import sys

KEEP = [
    "ID", "Kernel Name", "Grid Size", "Block Size",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    cols = [hdr.index(k) for k in KEEP if k in hdr]
    out = csv.writer(sys.stdout)
    for r in rows[start:]:
        if len(r) >= len(hdr):
            out.writerow([r[c] for c in cols])


if __name__ == "__main__":
    main()
