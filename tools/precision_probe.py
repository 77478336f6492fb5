"""Where does 16-bit storage lose accuracy?  CPU experiment on the fp32 oracle (no GPU): round activations and / or BN-folded
weights to fp16 (or bf16) at the same places the engine does, all layers or ONE layer at a time, and report the logit error
and the mask agreement at 0.5 against the unrounded fp32 forward on one 562x744 frame of the bench's weights.

    python tools/precision_probe.py [fp16|bf16] [H W]
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "oracle", ROOT):
    sys.path.insert(0, str(p))
import numpy as np
import torch
import torch.nn.functional as F
import aau_oracle as O
import bench

dt = torch.float16 if (len(sys.argv) < 2 or sys.argv[1] == "fp16") else torch.bfloat16
Hh, Ww = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (562, 744)
cfg, sd = bench.make_weights()
vol = O.synthetic_sweep(1, Hh, Ww, seed=31, peak=0)
x = torch.from_numpy(vol.astype(np.float32) / 255.0).unsqueeze(1)
if "--rand" in sys.argv:                                    # white-noise float frame + weights calibrated on noise (tests' r1_case)
    g = torch.Generator().manual_seed(11)
    sd = O.calibrate_bn(O.make_state_dict(cfg, 2025, "R1"), torch.rand(2, 1, Hh, Ww, generator=g), cfg)
    x = torch.rand(1, 1, Hh, Ww, generator=g)
ref = O.forward(sd, x, cfg)


def q(t):
    return t.to(dt).float()


ACT = {"on": set(), "all": False}          # layer names whose OUTPUT is rounded
orig_cbr, orig_gate, orig_ct, orig_aspp = O._cbr, O._gate_pipeline, F.conv_transpose2d, O._aspp


def cbr(ctx, x, p, padding=1, dilation=1):
    y = orig_cbr(ctx, x, p, padding, dilation)
    return q(y) if (ACT["all"] or p in ACT["on"]) else y


def gate(ctx, g, x, p):
    y, psi = orig_gate(ctx, g, x, p)
    return (q(y) if (ACT["all"] or p in ACT["on"]) else y), psi


def aspp(ctx, x, p="bridge", rates=(6, 12, 18)):
    y = orig_aspp(ctx, x, p, rates)          # (branch outputs are not rounded separately in this probe)
    return q(y) if (ACT["all"] or "bridge" in ACT["on"]) else y


CT_COUNT = {"i": 0}


def conv_t(inp, w, b, stride=2):
    y = orig_ct(inp, w, b, stride=stride)
    lvl = 4 - CT_COUNT["i"] % 4
    CT_COUNT["i"] += 1
    return q(y) if (ACT["all"] or f"u{lvl}.up" in ACT["on"]) else y


O._cbr, O._gate_pipeline, O._aspp = cbr, gate, aspp
O.F.conv_transpose2d = conv_t


def fold_and_round(sd, only=None):
    """BN folded into the conv weights (as aau_commit_weights does), folded weights rounded; BN left as a pure shift."""
    out = dict(sd)
    for k in list(sd.keys()):
        if k.endswith(".block.0.weight") or (k.startswith("bridge.blocks.") and k.endswith(".0.weight")) or k == "bridge.project.0.weight" \
                or k.endswith(".Wg.0.weight") or k.endswith(".Wx.0.weight"):
            bn = k[:-len("0.weight")] + "1"
            layer = k.split(".block.0")[0] if ".block.0" in k else k[:-len(".0.weight")]
            if only is not None and layer != only:
                continue
            s = sd[bn + ".weight"].double() / torch.sqrt(sd[bn + ".running_var"].double() + 1e-5)
            t = sd[bn + ".bias"].double() - sd[bn + ".running_mean"].double() * s
            out[k] = q((sd[k].double() * s.view(-1, 1, 1, 1)).float())
            out[bn + ".weight"] = torch.ones_like(sd[bn + ".weight"])
            out[bn + ".bias"] = t.float()
            out[bn + ".running_mean"] = torch.zeros_like(sd[bn + ".running_mean"])
            out[bn + ".running_var"] = torch.full_like(sd[bn + ".running_var"], 1.0 - 1e-5)
        elif k.endswith(".up.weight"):
            if only is not None and k[:-len(".weight")] != only:
                continue
            out[k] = q(sd[k])
    return out


def report(tag, out):
    e = (out - ref).abs()
    ag = ((out > 0) == (ref > 0)).float().mean().item()
    print(f"{tag:34s} max {e.max():.5f}  mean {e.mean():.6f}  rms {e.pow(2).mean().sqrt():.6f}  agree@0.5 {ag:.5f}", flush=True)
    return e.pow(2).mean().item()


print(f"dtype {dt}, frame {Hh}x{Ww}, logit std {ref.std():.3f}")
ACT["all"] = True
report("activations + weights (all layers)", O.forward(fold_and_round(sd), x, cfg))
report("activations only (all layers)", O.forward(sd, x, cfg))
ACT["all"] = False
report("weights only (all layers)", O.forward(fold_and_round(sd), x, cfg))
layers = ["d1.0", "d1.1", "d2.0", "d2.1", "d3.0", "d3.1", "d4.0", "d4.1", "bridge", "u4.up", "u4.att", "u4.conv.0", "u4.conv.1", "u3.up", "u3.att",
          "u3.conv.0", "u3.conv.1", "u2.up", "u2.att", "u2.conv.0", "u2.conv.1", "u1.up", "u1.conv.0", "u1.conv.1"]
tot_a = tot_w = 0.0
for L in (layers if "--layers" in sys.argv else []):
    ACT["on"] = {L}
    ms_a = report(f"  act  {L}", O.forward(sd, x, cfg)) if L != "u1.conv.1" else 0.0     # (u1.conv.1's output never leaves fp32)
    ACT["on"] = set()
    wl = {"bridge": None}.get(L, L)
    ms_w = 0.0
    if wl is not None and not L.endswith(".att"):
        ms_w = report(f"  wgt  {L}", O.forward(fold_and_round(sd, only=wl), x, cfg))
    tot_a += ms_a
    tot_w += ms_w
print(f"sum of single-layer mean-square errors: activations {tot_a:.3e}  weights {tot_w:.3e}")


# ---- what-if experiments (round 2): exact stem weights (hi + lo fp16 split on the tensor cores), and 3x3 weight rounding
# that preserves the sum over the 3x3 window of every (out, in) channel pair (smooth inputs see the SUM of the nine taps)
def round_keep_tap_sum(w):
    """w: [O, I, 3, 3] fp32 (BN folded).  Nearest rounding, then move the taps with the largest rounding error by one ulp
    in the direction that brings sum(rounded) back to round(sum(w)) -- each weight stays within one ulp of its value."""
    O_, I_ = w.shape[:2]
    flat = w.reshape(O_ * I_, 9).double()
    r = flat.float().to(dt).double()
    for _ in range(4):
        resid = flat.sum(1) - r.sum(1)                       # what the window sum lost
        up = torch.nextafter(r.float().to(dt), torch.full_like(r, float("inf")).to(dt)).double()
        dn = torch.nextafter(r.float().to(dt), torch.full_like(r, float("-inf")).to(dt)).double()
        step = torch.where(resid.unsqueeze(1) > 0, up - r, dn - r)          # signed one-ulp moves
        err = flat - r                                                       # individual rounding errors
        score = err * torch.sign(resid).unsqueeze(1)                         # prefer taps already rounded the "wrong" way
        k = score.argmax(1)
        idx = torch.arange(flat.shape[0])
        mv = step[idx, k]
        do = resid.abs() > mv.abs() * 0.5
        r[idx[do], k[do]] += mv[do]
    return r.float().reshape(w.shape)


def fold_variant(sd, exact_stem=True, keep_sum=True):
    out = fold_and_round(sd)
    for k in list(sd.keys()):
        if k.endswith(".block.0.weight"):
            bn = k[:-len("0.weight")] + "1"
            s = sd[bn + ".weight"].double() / torch.sqrt(sd[bn + ".running_var"].double() + 1e-5)
            wf = (sd[k].double() * s.view(-1, 1, 1, 1)).float()
            if k.startswith("d1.0.") and exact_stem:
                out[k] = wf
            elif keep_sum and not k.startswith("d1.0."):
                out[k] = round_keep_tap_sum(wf)
    return out


ACT["all"] = True
report("all, exact stem weights", O.forward(fold_variant(sd, True, False), x, cfg))
report("all, exact stem + tap-sum rounding", O.forward(fold_variant(sd, True, True), x, cfg))
ACT["all"] = False
report("weights only, exact stem", O.forward(fold_variant(sd, True, False), x, cfg))
report("weights only, exact stem + tap-sum", O.forward(fold_variant(sd, True, True), x, cfg))
