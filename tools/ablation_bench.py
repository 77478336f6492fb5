"""BASELINE.json configs[3]: the test_ablation variants (full / no attention gates / no ASPP / neither / att_depth 3) on
562x744 frames, base_c=32: frames/s of the forward at batch 28 and parity of logits + psi maps against the oracle on one
frame (both storage types: parity and timing).
    python tools/ablation_bench.py [out.json]"""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "att-aspp-unet_b200", ROOT / "oracle", ROOT):
    sys.path.insert(0, str(p))
import numpy as np, torch
import aau_oracle as O
from attention_aspp_unet import AttentionASPPUNet

H, W, B = 562, 744, 28
out_path = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "gpurun_out" / "ablation_bench.json")
variants = [("full", dict(use_att=True, use_aspp=True, att_depth=4), 158.60), ("no_att", dict(use_att=False, use_aspp=True, att_depth=4), 157.74),
            ("no_aspp", dict(use_att=True, use_aspp=False, att_depth=4), 146.36), ("neither", dict(use_att=False, use_aspp=False, att_depth=4), 145.51),
            ("att_depth3", dict(use_att=True, use_aspp=True, att_depth=3), 158.17)]
vol = O.synthetic_sweep(B, H, W, seed=2025, peak=B // 2)
x1 = torch.from_numpy(vol[:1].astype(np.float32) / 255.0).unsqueeze(1)
calib = torch.from_numpy(O.synthetic_sweep(2, H // 2, W // 2, seed=7, peak=1).astype(np.float32) / 255.0).unsqueeze(1)
rows = []
for name, kw, gflop in variants:
    cfg = O.NetCfg(base_c=32, variant="ablation", **kw)
    sd = O.calibrate_bn(O.make_state_dict(cfg, 2025, "R1"), calib, cfg)
    ref_logits, ref_psis = O.forward(sd, x1, cfg)
    rec = {"variant": name, **kw, "gflop_per_frame": gflop, "state_dict_entries": len(sd)}
    for dt in ("fp16", "bf16"):
        net = AttentionASPPUNet(base_c=32, act_dtype=dt, **kw)
        net.load_state_dict(sd, strict=True)
        net.eval()
        lg, psis = net(x1.cuda())
        err = (lg.cpu() - ref_logits).abs()
        rec[f"{dt}_logits_max_err"], rec[f"{dt}_logits_mean_err"] = float(err.max()), float(err.mean())
        rec[f"{dt}_mask_agreement_0.5"] = float(((lg.cpu() > 0) == (ref_logits > 0)).float().mean())
        rec[f"{dt}_psi_max_err"] = [float((a.cpu() - b).abs().max()) if a.numel() > 1 else 0.0 for a, b in zip(psis, ref_psis)]
        if True:                                                # both storage types are timed (fp16 is the headline since round 2)
            xb = torch.from_numpy(vol).cuda()
            out = torch.empty((B, 1, H, W), dtype=torch.float32, device="cuda")
            for _ in range(3):
                net(xb, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                net(xb, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            rec.update({f"{dt}_ms_per_forward": ms, f"{dt}_frames_per_s": 1e3 * B / ms, f"{dt}_tflops": B * gflop / ms, "launches": net.num_launches()})
        net.check_device()
        del net
    rec["logit_std"] = float(ref_logits.std())
    rows.append(rec)
    print(json.dumps(rec), flush=True)
Path(out_path).write_text(json.dumps({"config": "BASELINE.json configs[3]: ablation variants, 562x744, base_c=32, batch 28", "rows": rows}, indent=1))
