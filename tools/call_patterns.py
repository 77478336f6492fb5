"""The reference's own call patterns on one B200, engine vs the stock torch-eager / cuDNN path:
  (a) ROI wrapper: `torch.sigmoid(net(x[i:i+8]))` over 128 patches of 224x224, base 16   (model_attention_aspp.py:54)
  (b) CLI slice loop: flip-TTA of ONE 512x512 frame, base_c 32                            (attention_aspp_unet_pipeline_stage.py:336-338,495)
Both with the engine's CUDA-graph replay on (default for small batches) and off.  Writes a JSON to argv[1].
    python tools/call_patterns.py gpurun_out/r02_call_patterns.json"""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "att-aspp-unet_b200", ROOT / "oracle", ROOT):
    sys.path.insert(0, str(p))
import numpy as np, torch
import aau_oracle as O
from attention_aspp_unet import AttentionASPPUNet
from fetal_abdomen import FetalAbdomenSegmentation
import pipeline_predict as PP

dev = torch.device("cuda")
torch.backends.cudnn.benchmark = True                  # as the reference sets (attention_aspp_unet_pipeline_stage.py:553)
out = {"torch": torch.__version__, "rows": []}


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def weights(c, hw):
    cfg = O.NetCfg(base_c=c)
    sd = O.calibrate_bn(O.make_state_dict(cfg, 2025, "R1"), torch.rand(2, 1, hw, hw, generator=torch.Generator().manual_seed(3)), cfg)
    return cfg, sd


# ---- (a) 128 ROI patches, batches of 8, c = 16
cfg, sd = weights(16, 224)
x = torch.rand(128, 1, 224, 224, generator=torch.Generator().manual_seed(1)).to(dev)
sdc = {k: v.to(dev) for k, v in sd.items()}
net = AttentionASPPUNet(in_ch=1, num_classes=1, base=16)
net.load_state_dict(sd, strict=True)
net.eval().prepare(dev)
seg = FetalAbdomenSegmentation(net=net, batch=8)


def lib_a():
    with torch.no_grad():
        return torch.cat([torch.sigmoid(O.forward(sdc, x[i:i + 8], cfg)).squeeze(1) for i in range(0, 128, 8)])


def eng_a():
    outs = []
    for i in range(0, 128, 8):
        lg = net(x[i:i + 8])
        outs.append(torch.sigmoid(lg).squeeze(1))
    return torch.cat(outs)


ref = lib_a()
for graph in (-1, 0):
    net.set_option("graph", graph)
    got = eng_a()
    t = timed(eng_a, 10)
    out["rows"].append({"pattern": "wrapper: 16 x (8x224x224, c=16) + sigmoid", "impl": "engine", "graph_replay": bool(net.last_forward_was_graph()),
                        "ms_per_sweep_of_128": 1e3 * t, "frames_per_s": 128 / t, "max_abs_prob_diff_vs_cudnn_fp32": float((got - ref).abs().max())})
t = timed(lib_a, 10)
out["rows"].append({"pattern": "wrapper: 16 x (8x224x224, c=16) + sigmoid", "impl": "torch eager fp32 / cuDNN (benchmark=True)", "ms_per_sweep_of_128": 1e3 * t, "frames_per_s": 128 / t})

# ---- (b) flip-TTA of one 512x512 frame, c = 32
cfg, sd = weights(32, 256)
sdc = {k: v.to(dev) for k, v in sd.items()}
net = AttentionASPPUNet(base_c=32)
net.load_state_dict(sd, strict=True)
net.eval().prepare(dev)
pp = PP.PipelinePredictor(net, batch=1)
x1 = torch.rand(1, 1, 512, 512, generator=torch.Generator().manual_seed(2)).to(dev)


def lib_b():
    with torch.no_grad():
        l = O.forward(sdc, x1, cfg)
        lf = torch.flip(O.forward(sdc, torch.flip(x1, [-1]), cfg), [-1])
        return torch.sigmoid((l + lf) / 2)[0, 0].cpu().numpy()


def eng_b():
    return pp.predict_prob_tta(x1)[0].cpu().numpy()


ref = lib_b()
for graph in (-1, 0):
    net.set_option("graph", graph)
    got = eng_b()
    t = timed(eng_b, 30)
    out["rows"].append({"pattern": "CLI: flip-TTA of one 512x512 frame, c=32, prob back on the host", "impl": "engine", "graph_replay": bool(net.last_forward_was_graph()),
                        "ms_per_frame": 1e3 * t, "frames_per_s": 1 / t, "max_abs_prob_diff_vs_cudnn_fp32": float(np.abs(got - ref).max())})
t = timed(lib_b, 30)
out["rows"].append({"pattern": "CLI: flip-TTA of one 512x512 frame, c=32, prob back on the host", "impl": "torch eager fp32 / cuDNN (benchmark=True)", "ms_per_frame": 1e3 * t, "frames_per_s": 1 / t})
net.check_device()
for r in out["rows"]:
    print(r)
Path(sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "gpurun_out" / "r02_call_patterns.json")).write_text(json.dumps(out, indent=1))
