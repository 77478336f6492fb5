"""N forwards of the 562x744 workload at a given batch, nothing else (the command ncu wraps):
    python tools/one_forward.py [batch] [forwards] [name=value ...]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "att-aspp-unet_b200", ROOT / "oracle", ROOT):
    sys.path.insert(0, str(p))
import torch, bench
from attention_aspp_unet import AttentionASPPUNet
import aau_oracle as O
args = [a for a in sys.argv[1:] if "=" not in a]
batch = int(args[0]) if args else 8
n = int(args[1]) if len(args) > 1 else 3
cfg, sd = bench.make_weights()
net = AttentionASPPUNet(base_c=32)
net.load_state_dict(sd, strict=True)
net.eval().prepare("cuda")
for kv in sys.argv[1:]:
    if "=" in kv:
        k, v = kv.split("=")
        net.set_option(k, int(v))
x = torch.from_numpy(O.synthetic_sweep(batch, bench.H, bench.W, seed=1, peak=batch // 2)).cuda()
for _ in range(n):
    y = net(x)
torch.cuda.synchronize()
print("ok", tuple(y.shape), float(y.float().mean()))
