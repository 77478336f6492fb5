"""BASELINE.json config[4] ("large-batch throughput sweep"): batch 1..256 frames at 512x512 and 562x744, base_c=32.
For every point: frames/s of the forward (CUDA events, inputs resident), tensor-core TFLOP/s of the igemm launches
(profile mode: CUDA events around each launch) and algorithmic GB/s of the three attention-gate launches.
    python tools/batch_sweep.py [out.json]"""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "att-aspp-unet_b200", ROOT / "oracle", ROOT):
    sys.path.insert(0, str(p))
import numpy as np, torch
import bench
from attention_aspp_unet import AttentionASPPUNet

out_path = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "gpurun_out" / "batch_sweep.json")
cfg, sd = bench.make_weights()
net = AttentionASPPUNet(base_c=32)
net.load_state_dict(sd, strict=True)
net.eval().prepare("cuda")
rows = []
for (H, W) in ((512, 512), (562, 744)):
    gflop = 100.839 if H == 512 else 160.319
    for B in (1, 2, 4, 8, 16, 32, 64, 128, 256):
        x = torch.randint(0, 256, (B, H, W), dtype=torch.uint8, generator=torch.Generator().manual_seed(2025)).cuda()
        out = torch.empty((B, 1, H, W), dtype=torch.float32, device="cuda")
        for _ in range(3):
            net(x, out=out)
        torch.cuda.synchronize()
        reps = max(3, min(40, 512 // B))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            net(x, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        net.set_option("profile", 1)
        net(x, out=out)
        prof = net.op_profile()
        net.set_option("profile", 0)
        tc = [r for r in prof if r["kernel"] == "igemm_tc_kernel"]
        gate = [r for r in tc if ".att" in r["layer"]]
        conv = [r for r in tc if ".att" not in r["layer"]]
        row = {"H": H, "W": W, "batch": B, "ms_per_forward": ms, "frames_per_s": 1e3 * B / ms,
               "whole_forward_tflops": B * gflop / ms, "conv_igemm_tflops": sum(r["flops"] for r in conv) / sum(r["ms"] for r in conv) / 1e9,
               "gate_gbs": sum(r["bytes"] for r in gate) / sum(r["ms"] for r in gate) / 1e6, "launches": net.num_launches()}
        rows.append(row)
        print("%dx%d B=%3d  %8.3f ms  %7.1f frames/s  forward %6.1f TF/s  conv igemm %6.1f TF/s  gates %5.0f GB/s" %
              (H, W, B, ms, row["frames_per_s"], row["whole_forward_tflops"], row["conv_igemm_tflops"], row["gate_gbs"]), flush=True)
        del x, out
        torch.cuda.empty_cache()
net.check_device()
Path(out_path).write_text(json.dumps({"config": "BASELINE.json configs[4]: batch sweep, base_c=32, bf16, uint8 input", "rows": rows}, indent=1))
