"""Pick the roofline-relevant metrics out of `ncu -i X.ncu-rep --page raw --csv` (one block per profiled launch).
usage: ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv; python scripts/ncu_raw_summary.py /tmp/raw.csv"""
import csv
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active",
        "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate", "sm__cycles_active.avg", "smsp__inst_executed.sum",
        "sm__pipe_tensor_subpipe", "sm__pipe_tc"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = [i for i, h in enumerate(hdr) if any(h.startswith(w) for w in WANT)]
for r in rows[2:]:
    for i in idx:
        print("%-70s %s %s" % (hdr[i], r[i], units[i]))
    print("---")
