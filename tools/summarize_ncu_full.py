#!/usr/bin/env python
"""Condense an `ncu --set full` report into the per-launch columns kept under profiles/.

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > /tmp/x.csv
    python tools/summarize_ncu_full.py /tmp/x.csv > profiles/r1X_ncu_full_*.csv

    # bench.py's roofline.traffic: DRAM bytes of the dominant launch of one eval forward, keyed by
    # precision mode, stamped with the sha256 of the library the capture was taken from
    python tools/summarize_ncu_full.py /tmp/x.csv --traffic fp16 --lib-sha <sha256> \
        --source profiles/r2X_ncu_full_eval_fp16.csv
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


def traffic(argv):
    """Update profiles/traffic.json from the raw csv: the dominant launch = the longest
    conv_gemm_kernel launch of the capture."""
    import json
    import os
    path = argv[1]
    mode = argv[argv.index("--traffic") + 1]
    sha = argv[argv.index("--lib-sha") + 1] if "--lib-sha" in argv else None
    source = argv[argv.index("--source") + 1] if "--source" in argv else os.path.basename(path)
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units = rows[start], rows[start + 1]
    col = {k: hdr.index(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum",
                                     "dram__bytes_write.sum")}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    best = None
    for r in rows[start + 2:]:
        if len(r) < len(hdr) or "conv_gemm_kernel" not in r[col["Kernel Name"]]:
            continue
        dur = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
        if best is None or dur > best[0]:
            rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * scale[units[col["dram__bytes_read.sum"]]]
            wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * scale[units[col["dram__bytes_write.sum"]]]
            best = (dur, rd, wr, r[col["Kernel Name"]].split("(")[0], r[0])
    out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    try:
        data = json.load(open(out_path))
    except (OSError, ValueError):
        data = {"kernels": {}}
    if sha and data.get("lib_sha256") != sha:
        data = {"kernels": {}}            # captures of another build do not mix
    data["lib_sha256"] = sha
    data["source"] = source
    data["kernels"]["dominant_eval_" + mode] = {
        "kernel": best[3], "ncu_id": best[4], "duration": best[0], "duration_unit": units[col["gpu__time_duration.sum"]],
        "dram_read_bytes": best[1], "dram_write_bytes": best[2], "dram_bytes": best[1] + best[2]}
    json.dump(data, open(out_path, "w"), indent=1)
    print(json.dumps(data["kernels"]["dominant_eval_" + mode]))


def main():
    if "--traffic" in sys.argv:
        return traffic(sys.argv)
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
