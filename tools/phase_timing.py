"""Role phase timing of one forward (needs att-aspp-unet_b200/libaau_timing.so built with -DAAU_EPI_TIMING):
    AAU_LIB=att-aspp-unet_b200/libaau_timing.so python tools/phase_timing.py [batch] [name=value ...]
The engine prints, per igemm launch, cycle sums over all CTAs of: MMA warp (wait tmem_empty, wait operands, issue+bookkeeping)
and the first warp of epilogue groups 0 / 1 (bookkeeping, wait store-read, wait accumulator, top barrier, convert, fence+barrier,
pool+fence+barrier, store issue).  This script turns them into cycles per tile."""
import re, subprocess, sys, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "att-aspp-unet_b200", ROOT / "oracle", ROOT):
    sys.path.insert(0, str(p))
import torch, bench
from attention_aspp_unet import AttentionASPPUNet
import aau_oracle as O
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 28
cfg, sd = bench.make_weights()
net = AttentionASPPUNet(base_c=32)
net.load_state_dict(sd, strict=True)
net.eval().prepare("cuda")
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    net.set_option(k, int(v))
x = torch.from_numpy(O.synthetic_sweep(batch, bench.H, bench.W, seed=1, peak=batch // 2)).cuda()
net(x); net(x)
torch.cuda.synchronize()
sys.stderr.flush()
print("=== profiled forward (cycles are sums over CTAs; see tools/phase_timing.py)", file=sys.stderr, flush=True)
net.set_option("profile", 1)
net(x)
torch.cuda.synchronize()
