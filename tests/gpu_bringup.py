"""Layer-by-layer bring-up report for the CUDA path (run on the GPU box; not a pytest file).

    python tests/gpu_bringup.py            # runs every case in its own subprocess with a timeout
    python tests/gpu_bringup.py CASE_JSON  # one case, in-process

For each case the engine's named intermediates are compared with the oracle's taps, so the first wrong layer is
visible from one GPU call.  Output: one JSON line per case in gpurun_out/bringup.jsonl.
"""
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "att-aspp-unet_b200", ROOT / "oracle"):
    sys.path.insert(0, str(p))


def run_case(case):
    import numpy as np
    import torch
    import aau_oracle as O
    from attention_aspp_unet import AttentionASPPUNet

    cfg = O.NetCfg(base_c=case["c"], variant=case.get("variant", "pipeline"), use_att=case.get("use_att", True),
                   use_aspp=case.get("use_aspp", True), att_depth=case.get("att_depth", 4))
    B, H, W = case["shape"]
    g = torch.Generator().manual_seed(11)
    sd = O.make_state_dict(cfg, seed=2025, regime="R1")
    sd = O.calibrate_bn(sd, torch.rand(2, 1, H, W, generator=g), cfg)
    x = torch.rand(B, 1, H, W, generator=g)
    taps = {}
    ref = O.forward(sd, x, cfg, taps=taps)
    ref_logits = ref if cfg.variant == "pipeline" else ref[0]
    kw = dict(base_c=cfg.base_c, act_dtype=case.get("dtype", "bf16"))
    if cfg.variant == "ablation":
        kw.update(use_att=cfg.use_att, use_aspp=cfg.use_aspp, att_depth=cfg.att_depth)
    net = AttentionASPPUNet(**kw)
    net.load_state_dict(sd, strict=True)
    net.eval().prepare("cuda")
    if "amode" in case:
        net.set_option("amode", case["amode"])
    t0 = time.time()
    out = net(x.cuda())
    net.check_device()
    logits = (out if cfg.variant == "pipeline" else out[0]).cpu()
    res = {"case": case, "launches": net.num_launches(), "sec": round(time.time() - t0, 3), "layers": {}}
    names = {"d1.0": "d1.0", "x1": "x1", "p4": "p4", "bridge": "bridge", "g4": "g4", "u4a": "u4a", "d4": "u4", "g3": "g3",
             "d3": "u3", "g2": "g2", "d2": "u2", "g1": "g1", "u1a": "u1a"}
    for lvl in (2, 3, 4):
        names[f"x{lvl}"] = f"xatt{lvl}" if lvl in cfg.gate_levels() else f"x{lvl}"
    for eng, orc in names.items():
        try:
            t = net.debug_tensor(eng).cpu()
        except Exception as e:   # noqa
            res["layers"][eng] = "n/a: " + str(e)[:60]
            continue
        r = taps[orc]
        d = (t - r).abs()
        res["layers"][eng] = {"max_err": round(d.max().item(), 5), "mean_err": round(d.mean().item(), 6),
                              "ref_absmax": round(r.abs().max().item(), 4), "nan": int(torch.isnan(t).sum().item())}
    d = (logits - ref_logits).abs()
    res["logits"] = {"max_err": round(d.max().item(), 5), "mean_err": round(d.mean().item(), 6), "p99": round(d.flatten().kthvalue(int(0.99 * d.numel())).values.item(), 5),
                     "ref_std": round(ref_logits.std().item(), 4), "nan": int(torch.isnan(logits).sum().item())}
    for thr in (0.05, 0.48, 0.5):
        a = torch.sigmoid(logits) > thr
        b = torch.sigmoid(ref_logits) > thr
        res["logits"][f"agree@{thr}"] = round((a == b).float().mean().item(), 6)
    if cfg.variant == "ablation":
        for i, nm in enumerate(("psi3", "psi2")):
            if out[1][i].numel() > 1:
                res[nm] = round((out[1][i].cpu() - ref[1][i]).abs().max().item(), 5)
    return res


CASES = [
    {"c": 32, "shape": [2, 64, 64], "amode": 0, "dtype": "fp16"},
    {"c": 32, "shape": [2, 64, 64], "dtype": "fp16"},
    {"c": 32, "shape": [1, 141, 93], "dtype": "fp16"},
    {"c": 16, "shape": [2, 80, 72], "dtype": "fp16"},
    {"c": 48, "shape": [1, 64, 80], "dtype": "fp16"},
    {"c": 16, "shape": [2, 80, 72], "variant": "ablation", "dtype": "fp16"},
    {"c": 32, "shape": [2, 562, 744], "dtype": "fp16"},
    {"c": 32, "shape": [2, 562, 744]},
]

if __name__ == "__main__":
    if len(sys.argv) > 1:
        print("RESULT " + json.dumps(run_case(json.loads(sys.argv[1]))))
        sys.exit(0)
    out_dir = ROOT / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    with open(out_dir / "bringup.jsonl", "w") as f:
        for case in CASES:
            try:
                p = subprocess.run([sys.executable, __file__, json.dumps(case)], capture_output=True, text=True, timeout=240)
                line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
                rec = json.loads(line[0][7:]) if line else {"case": case, "error": (p.stderr or p.stdout)[-1500:], "rc": p.returncode}
            except subprocess.TimeoutExpired:
                rec = {"case": case, "error": "timeout"}
            f.write(json.dumps(rec) + "\n")
            f.flush()
            print(json.dumps(rec)[:3000], flush=True)
