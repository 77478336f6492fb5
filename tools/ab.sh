#!/bin/bash
# A/B of engine options on ONE box: tools/ab.sh "<opts A>" "<opts B>" [repeats]; prints value + per-layer ms for both.
A="$1"; B="$2"; R="${3:-2}"
for i in $(seq 1 $R); do
  for V in A B; do
    O="$A"; [ $V = B ] && O="$B"
    E="${O%%;*}"; [ "$E" = "$O" ] && E=""; O="${O#*;}"       # "ENV=val ...;bench options" (the env part is optional)
    env $E python bench.py --steps 2 --warmup 2 --lean --batch ${AB_BATCH:-56} $O > gpurun_out/ab_$V$i.json 2> gpurun_out/ab_$V$i.err
    cp gpurun_out/layers_b${AB_BATCH:-56}_${AB_DTYPE:-fp16}.json gpurun_out/ab_layers_$V$i.json
  done
done
python - "$R" <<'PY'
import json, sys
R = int(sys.argv[1])
def load(v):
    vals, lay = [], {}
    for i in range(1, R + 1):
        try:
            vals.append(json.load(open(f"gpurun_out/ab_{v}{i}.json"))["value"])
            for r in json.load(open(f"gpurun_out/ab_layers_{v}{i}.json"))["rows"]:
                lay.setdefault(r["layer"].split(" [")[0], []).append((r["ms_per_fwd"], r["layer"]))
        except Exception as e:
            print("missing", v, i, e)
    return vals, lay
va, la = load("A"); vb, lb = load("B")
print("A fps", ["%.0f" % v for v in va], " B fps", ["%.0f" % v for v in vb])
for k in la:
    a = min(x[0] for x in la[k]); b = min(x[0] for x in lb.get(k, [(0, "")]))
    print("%-28s A %.3f  B %.3f  %+5.1f%%   %s | %s" % (k, a, b, 100 * (b - a) / a if a else 0, la[k][0][1].split(" [")[-1][:40], lb.get(k, [(0, "")])[0][1].split(" [")[-1][:40]))
print("sum A %.3f B %.3f" % (sum(min(x[0] for x in v) for v in la.values()), sum(min(x[0] for x in v) for v in lb.values())))
PY
