"""Per-layer A/B of engine option sets in ONE process (same box, same clocks): for every option set, `reps` profiled forwards of
the 562x744 workload at the bench batch, then the mean CUDA-event time of every launch.
    python tools/layer_ab.py [batch] [reps] "name=v,name=v" "name=v" ...      (an empty string = defaults)"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "att-aspp-unet_b200", ROOT / "oracle", ROOT):
    sys.path.insert(0, str(p))
import torch, bench
from attention_aspp_unet import AttentionASPPUNet
import aau_oracle as O

batch, reps = int(sys.argv[1]), int(sys.argv[2])
sets = sys.argv[3:] or [""]
cfg, sd = bench.make_weights()
x = torch.from_numpy(O.synthetic_sweep(batch, bench.H, bench.W, seed=1, peak=batch // 2)).cuda()
out = torch.empty((batch, 1, bench.H, bench.W), dtype=torch.float32, device="cuda")
results = []
nets = []
for spec in sets:
    net = AttentionASPPUNet(base_c=32)
    net.load_state_dict(sd, strict=True)
    net.eval().prepare("cuda")
    for kv in [t for t in spec.split(",") if t]:
        k, v = kv.split("=")
        net.set_option(k, int(v))
    nets.append(net)
for rnd in range(2):                                   # interleave the sets: two rounds each, so that clock drift hits all alike
    for si, net in enumerate(nets):
        for _ in range(3):
            net(x, out=out)
        net.set_option("profile", 1)
        acc = {}
        for _ in range(reps):
            net(x, out=out)
            for r in net.op_profile():
                a = acc.setdefault(r["layer"].split(" [")[0], [0.0, r["layer"]])
                a[0] += r["ms"]
        net.set_option("profile", 0)
        if rnd == 0:
            results.append(acc)
        else:
            for k, v in acc.items():
                results[si][k][0] = (results[si][k][0] + v[0]) / 2
        net.check_device()
names = []
for acc in results:
    for k in acc:
        if k not in names:
            names.append(k)
print("%-34s" % "layer" + "".join("%14s" % (s[:13] or "default") for s in sets))
for k in names:
    print("%-34s" % k[:34] + "".join("%14.4f" % (acc[k][0] / reps if k in acc else float("nan")) for acc in results))
print("%-34s" % "sum" + "".join("%14.4f" % (sum(v[0] for v in acc.values()) / reps) for acc in results))
for si, acc in enumerate(results):
    print(sets[si] or "default", "|", " ; ".join(v[1].split(" [")[-1][:48] for k, v in acc.items() if "bridge.blocks" in k))
