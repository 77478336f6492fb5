"""Print one bench JSON line + the per-layer table bench.py wrote (gpurun_out/layers_b<B>_<dtype>.json)."""
import json, sys
d = json.load(open(sys.argv[1]))
print("value %.1f e2e %.1f igemm %.1f TF/s frac %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d["clocks"]))
L = json.load(open(sys.argv[2]))
print("fwd ms", L["fwd_ms"])
for r in L["rows"]:
    print("%-90s %6.3f %7.1fTF %6.0fGB/s" % (r["layer"], r["ms_per_fwd"], r["tflops"], r["gbs"]))
