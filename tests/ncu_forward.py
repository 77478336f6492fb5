"""Profiling target: warm the engine, then run ONE forward of `batch` 562x744 frames (ncu skips the warm-up launches).

    python tests/ncu_forward.py [batch] [dtype] [warm_forwards]
Prints the number of launches per forward so that `ncu -s <warm*launches> -c <launches>` brackets the last forward.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "att-aspp-unet_b200", ROOT / "oracle", ROOT):
    sys.path.insert(0, str(p))
import bench  # noqa: E402
from attention_aspp_unet import AttentionASPPUNet  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dtype = sys.argv[2] if len(sys.argv) > 2 else "bf16"
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg, sd = bench.make_weights()
net = AttentionASPPUNet(base_c=32, act_dtype=dtype)
net.load_state_dict(sd, strict=True)
net.eval()
import aau_oracle as O  # noqa: E402
x = torch.from_numpy(O.synthetic_sweep(batch, bench.H, bench.W, seed=1, peak=batch // 2)).cuda()
for _ in range(warm):
    net(x)
torch.cuda.synchronize()
out = net(x)
torch.cuda.synchronize()
net.check_device()
print("launches_per_forward", net.num_launches(), "batch", batch, "logits std", float(out.std()))
