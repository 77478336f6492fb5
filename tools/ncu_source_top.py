"""Top stall sites of an `ncu --page source --csv --print-source sass` dump (one or more kernels)."""
import csv, sys
path, topn = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
kern, hdr, rows = None, None, []
def flush():
    if not rows: return
    idx = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[idx["# Samples"]] or 0) for r in rows)
    print(f"== {kern}: {len(rows)} instr, {tot} samples, inst executed {sum(int(r[idx['Instructions Executed']] or 0) for r in rows)}")
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in sorted(rows, key=lambda r: -int(r[idx["# Samples"]] or 0))[:topn]:
        n = int(r[idx["# Samples"]] or 0)
        st = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
        print(f"{n:7d} {100.0*n/max(tot,1):5.1f}%  exec {r[idx['Instructions Executed']]:>9}  {r[idx['Address']][-5:]}  {r[idx['Source']][:70]:70s} {st}")
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == "Kernel Name":
        flush(); kern, hdr, rows = r[1][:60], None, []
    elif r[0] == "Address":
        hdr = r
    elif hdr is not None and len(r) >= len(hdr) - 2:
        rows.append(r)
flush()
