"""Condense an `ncu --page raw --csv` dump into one row per launch with the counters the roofline arguments use."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
                        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
                        "sm__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__registers_per_thread",
                        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
idx = {h: i for i, h in enumerate(hdr)}
cols = [w for w in want if w in idx]
print("id kernel | " + " | ".join(f"{c} [{units[idx[c]]}]" for c in cols))
for r in data:
    print(r[idx["ID"]], r[idx["Kernel Name"]][:24], "|", " | ".join(r[idx[c]] for c in cols))
