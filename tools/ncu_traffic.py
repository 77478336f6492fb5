"""Per-launch DRAM traffic of one forward from an `ncu --set full ... --page raw --csv` dump of the igemm launches:
writes profiles/<tag>_igemm_dram_traffic.json (bytes per frame, summed over the igemm launches of one forward), which
bench.py reports as roofline.traffic (scaled to its batch).
    python tools/ncu_traffic.py gpurun_out/x_raw.csv <frames in the profiled forward> profiles/r01_igemm_dram_traffic.json [launches per forward]"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
frames = int(sys.argv[2])
idx = {h: i for i, h in enumerate(hdr)}
def val(r, name):
    v, u = float(r[idx[name]].replace(",", "")), units[idx[name]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
launches = []
for r in data:
    launches.append({"id": int(r[idx["ID"]]), "kernel": r[idx["Kernel Name"]][:40], "us": float(r[idx["gpu__time_duration.sum"]]),
                     "dram_read": val(r, "dram__bytes_read.sum"), "dram_write": val(r, "dram__bytes_write.sum"),
                     "tensor_pipe_pct": (float(r[idx["sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"]])
                                         if "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active" in idx else None)})
tot = sum(l["dram_read"] + l["dram_write"] for l in launches)
if len(sys.argv) > 4:                                     # keep the launches of ONE forward: the last N of the capture
    launches = launches[-int(sys.argv[4]):]
tot = sum(l["dram_read"] + l["dram_write"] for l in launches)
out = {"source": sys.argv[1], "frames_in_profiled_forward": frames, "batch": frames, "launches": len(launches),
       "dram_bytes_per_forward": tot, "dram_bytes_per_frame": tot / frames, "per_launch": launches}
json.dump(out, open(sys.argv[3], "w"), indent=1)
print("launches", len(launches), "dram bytes per frame %.1f MB" % (tot / frames / 1e6))
