"""One small forward (2 x 141 x 186 uint8, base_c 32) through every default kernel, checked against the oracle (a quick bring-up aid):
    python tools/small_forward.py [B H W]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "att-aspp-unet_b200", ROOT / "oracle", ROOT):
    sys.path.insert(0, str(p))
import numpy as np, torch
import aau_oracle as O
from attention_aspp_unet import AttentionASPPUNet
B, H, W = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (2, 141, 186)
cfg = O.NetCfg(base_c=32)
sd = O.calibrate_bn(O.make_state_dict(cfg, 2025, "R1"), torch.rand(2, 1, H, W, generator=torch.Generator().manual_seed(3)), cfg)
vol = O.synthetic_sweep(B, H, W, seed=6, peak=B // 2)
ref = O.forward(sd, torch.from_numpy(vol.astype(np.float32) / 255.0).unsqueeze(1), cfg)
net = AttentionASPPUNet(base_c=32)
net.load_state_dict(sd, strict=True)
net.eval()
out = net(torch.from_numpy(vol).cuda())
net.check_device()
torch.cuda.synchronize()
print("ok max|err| %.4f (spread %.3f)" % ((out.cpu() - ref).abs().max().item(), ref.std().item()))
