"""CPU oracle for the AttentionASPPUNet hot path.  TEST INFRASTRUCTURE ONLY.

This file is a restatement, in plain fp32 ``torch.nn.functional`` calls and numpy/scipy, of the
reference's algorithm for the hot path.  It exists so the CUDA path can be checked on boxes where
``/root/reference`` is not mounted.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product package
(``att-aspp-unet_b200/``) must never import it.

Parity status: PINNED.  ``oracle/gen_golden.py`` imports the real reference modules from
``/root/reference`` (with inert stubs for its missing third-party imports), loads the state dicts this
file generates with ``strict=True`` (pins key names + shapes), runs the reference forward and commits
the outputs under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks this restatement against
those vectors on every run.  The reference itself ships no tests or golden tensors.

The arithmetic of the reference lives in PyTorch (pinned ``torch==2.2.0``, requirements.txt:6); the
functional ops used here are the same ATen kernels the reference's ``nn.Module`` objects dispatch to.

Reference lines restated (all paths relative to /root/reference):
  ConvBNReLU           attention_aspp_unet_pipeline_stage.py:59-65   / test_ablation.py:73-83
  ASPP                 attention_aspp_unet_pipeline_stage.py:67-83   / test_ablation.py:85-126
  AttentionGate        attention_aspp_unet_pipeline_stage.py:85-92   (flavour "pipeline")
  AttentionGate        test_ablation.py:128-143                      (flavour "ablation")
  DummyAttention       attention_aspp_unet_pipeline_stage.py:95-96   / test_ablation.py:145-147
  UpBlock              attention_aspp_unet_pipeline_stage.py:98-109  / test_ablation.py:149-166
  AttentionASPPUNet    attention_aspp_unet_pipeline_stage.py:111-127 / test_ablation.py:168-218
  postprocess          model_attention_aspp.py:69-89
  select_fetal_abdomen_mask_and_frame   model_attention_aspp.py:91-97
  predict_prob_tta     attention_aspp_unet_pipeline_stage.py:336-338
  convert_2d_mask_to_3d inference.py:257-273
  load_image_file_as_array (frame conditioning)   model_attention_aspp.py:11-17
  crop_roi_224         model_attention_aspp.py:20-30
  FetalAbdomenSegmentation.predict (sample 128, ROI, batch 8, paste back)   model_attention_aspp.py:41-65
  write_array_as_image_file (volume semantics)    inference.py:208-254
  refine_mask / _circularity_score / select_best  test_ablation.py:373-403 (pipeline twin :340-353; its select_best is broken)
  measure_ac_mm / _ellipse_circum                 attention_aspp_unet_pipeline_stage.py:355-374
  pipeline slice loop (resize 512, TTA, resize back, blur, threshold, refine)   attention_aspp_unet_pipeline_stage.py:486-501
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm2d default; the reference never overrides it


@dataclass(frozen=True)
class NetCfg:
    """Constructor arguments of the two reference classes."""
    base_c: int = 32
    in_channels: int = 1
    num_classes: int = 1
    variant: str = "pipeline"      # "pipeline" (attention_aspp_unet_pipeline_stage.py) | "ablation" (test_ablation.py)
    use_att: bool = True           # ablation only
    use_aspp: bool = True          # ablation only
    att_depth: int = 4             # ablation only

    def gate_levels(self) -> Tuple[int, ...]:
        """Decoder levels (4 = u4 … 1 = u1) that carry a real attention gate."""
        if self.variant == "pipeline":
            return (4, 3, 2)                       # :120-121, u1 has use_att=False
        lv = []
        if self.use_att and self.att_depth >= 4:
            lv.append(4)                           # test_ablation.py:199
        if self.use_att and self.att_depth >= 3:
            lv.append(3)                           # test_ablation.py:200
        return tuple(lv)

    def f_int(self, level: int) -> int:
        out_c = self.base_c * (1 << (level - 1))   # u4: 8c, u3: 4c, u2: 2c
        if self.variant == "pipeline":
            return out_c // 2                      # :102
        return max(8, out_c // 4)                  # test_ablation.py:131-132


# --------------------------------------------------------------------------------------------
# state_dict layout
# --------------------------------------------------------------------------------------------
def _bn_entries(prefix: str, n: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    return [(prefix + ".weight", (n,), "bn_w"), (prefix + ".bias", (n,), "bn_b"),
            (prefix + ".running_mean", (n,), "bn_m"), (prefix + ".running_var", (n,), "bn_v"),
            (prefix + ".num_batches_tracked", (), "bn_n")]


def _cbr_entries(prefix: str, cin: int, cout: int, k: int = 3):
    return [(prefix + ".block.0.weight", (cout, cin, k, k), "conv_w")] + _bn_entries(prefix + ".block.1", cout)


def state_dict_spec(cfg: NetCfg) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, kind) in the reference's ``state_dict()`` order."""
    c = cfg.base_c
    spec: List[Tuple[str, Tuple[int, ...], str]] = []
    chans = [cfg.in_channels, c, 2 * c, 4 * c, 8 * c]
    for lvl in range(1, 5):
        spec += _cbr_entries(f"d{lvl}.0", chans[lvl - 1], chans[lvl])
        spec += _cbr_entries(f"d{lvl}.1", chans[lvl], chans[lvl])
    if cfg.variant == "pipeline" or cfg.use_aspp:
        ic, oc = 8 * c, 16 * c
        spec += [("bridge.blocks.0.0.weight", (oc, ic, 1, 1), "conv_w")] + _bn_entries("bridge.blocks.0.1", oc)
        for i in (1, 2, 3):
            spec += [(f"bridge.blocks.{i}.0.weight", (oc, ic, 3, 3), "conv_w")] + _bn_entries(f"bridge.blocks.{i}.1", oc)
        spec += [("bridge.pool.1.weight", (oc, ic, 1, 1), "conv_w")] + _bn_entries("bridge.pool.2", oc)
        spec += [("bridge.project.0.weight", (oc, 5 * oc, 1, 1), "conv_w")] + _bn_entries("bridge.project.1", oc)
    else:
        spec += _cbr_entries("bridge.0", 8 * c, 16 * c)
    gates = cfg.gate_levels()
    for lvl in (4, 3, 2, 1):
        out_c = c * (1 << (lvl - 1))
        in_c = 2 * out_c
        p = f"u{lvl}"
        spec += [(p + ".up.weight", (in_c, out_c, 2, 2), "convT_w"), (p + ".up.bias", (out_c,), "bias")]
        if lvl in gates:
            fi = cfg.f_int(lvl)
            if cfg.variant == "pipeline":
                spec += [(p + ".att.Wg.0.weight", (fi, out_c, 1, 1), "conv_w")] + _bn_entries(p + ".att.Wg.1", fi)
                spec += [(p + ".att.Wx.0.weight", (fi, out_c, 1, 1), "conv_w")] + _bn_entries(p + ".att.Wx.1", fi)
                spec += [(p + ".att.psi.0.weight", (1, fi, 1, 1), "conv_w")] + _bn_entries(p + ".att.psi.1", 1)
            else:
                spec += [(p + ".att.Wg.weight", (fi, out_c, 1, 1), "conv_w"),
                         (p + ".att.Wx.weight", (fi, out_c, 1, 1), "conv_w"),
                         (p + ".att.psi.1.weight", (1, fi, 1, 1), "conv_w"),
                         (p + ".att.psi.1.bias", (1,), "bias")]
        spec += _cbr_entries(p + ".conv.0", in_c, out_c)
        spec += _cbr_entries(p + ".conv.1", out_c, out_c)
    spec += [("out_conv.weight", (cfg.num_classes, c, 1, 1), "conv_w"), ("out_conv.bias", (cfg.num_classes,), "bias")]
    return spec


def make_state_dict(cfg: NetCfg, seed: int = 2025, regime: str = "R0") -> Dict[str, torch.Tensor]:
    """Deterministic synthetic weights (no checkpoints ship with the reference).

    R0: the distributions of PyTorch's default init (conv: U(±1/sqrt(fan_in)); BN: γ=1, β=0, μ=0, σ²=1).
    R1: same conv weights, γ~U(0.5,1.5), β~N(0,0.2²); running stats must then be filled by
        :func:`calibrate_bn` (SURVEY.md §7 hard part 1).
    Generated from a CPU ``torch.Generator`` so the same tensors appear on every box.
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for key, shape, kind in state_dict_spec(cfg):
        if kind in ("conv_w", "convT_w"):
            # nn.Conv2d / nn.ConvTranspose2d: kaiming_uniform_(a=sqrt(5)) -> bound = 1/sqrt(fan_in),
            # fan_in = shape[1] * kh * kw for both (torch's _calculate_fan_in_and_fan_out)
            fan_in = shape[1] * shape[2] * shape[3]
            bound = 1.0 / np.sqrt(fan_in)
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "bias":
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        elif kind == "bn_w":
            sd[key] = torch.ones(shape) if regime == "R0" else torch.rand(shape, generator=g) + 0.5
        elif kind == "bn_b":
            sd[key] = torch.zeros(shape) if regime == "R0" else torch.randn(shape, generator=g) * 0.2
        elif kind == "bn_m":
            sd[key] = torch.zeros(shape)
        elif kind == "bn_v":
            sd[key] = torch.ones(shape)
        elif kind == "bn_n":
            sd[key] = torch.zeros(shape, dtype=torch.int64)
        else:  # pragma: no cover
            raise ValueError(kind)
    return sd


# --------------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------------
class _Ctx:
    def __init__(self, sd, calibrate: bool):
        self.sd = sd
        self.calibrate = calibrate


def _bn(ctx: _Ctx, x: torch.Tensor, p: str) -> torch.Tensor:
    sd = ctx.sd
    if ctx.calibrate and x.shape[2] * x.shape[3] == 1:
        # BN on the globally pooled 1x1 map (ASPP image branch): a calibration batch of a few frames gives a
        # near-zero variance and scale factors in the hundreds, which no trained network has.  Keep the module
        # defaults (mean 0, var 1) there; gamma / beta stay random.
        return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                            False, 0.0, BN_EPS)
    if ctx.calibrate:
        # train-mode BatchNorm2d with momentum=None after exactly one batch: running stats become the
        # batch mean and the *unbiased* batch variance; the batch itself is normalised with the biased one.
        n = x.numel() // x.shape[1]
        mean = x.mean((0, 2, 3))
        var_b = x.var((0, 2, 3), unbiased=False)
        sd[p + ".running_mean"] = mean.clone()
        sd[p + ".running_var"] = var_b * (n / max(n - 1, 1))
        sd[p + ".num_batches_tracked"] = torch.ones((), dtype=torch.int64)
        return F.batch_norm(x, None, None, sd[p + ".weight"], sd[p + ".bias"], True, 0.0, BN_EPS)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, BN_EPS)


def _cbr(ctx, x, p, padding=1, dilation=1):
    """ConvBNReLU: conv(no bias) -> BN -> ReLU  (attention_aspp_unet_pipeline_stage.py:59-65)."""
    x = F.conv2d(x, ctx.sd[p + ".block.0.weight"], None, 1, padding, dilation)
    return F.relu(_bn(ctx, x, p + ".block.1"))


def _aspp(ctx, x, p="bridge", rates=(6, 12, 18)):
    """ASPP (attention_aspp_unet_pipeline_stage.py:67-83): cat order [1x1, d6, d12, d18, pool]."""
    sd = ctx.sd
    h, w = x.shape[2:]
    feats = [F.relu(_bn(ctx, F.conv2d(x, sd[f"{p}.blocks.0.0.weight"]), f"{p}.blocks.0.1"))]
    for i, r in enumerate(rates, start=1):
        y = F.conv2d(x, sd[f"{p}.blocks.{i}.0.weight"], None, 1, r, r)
        feats.append(F.relu(_bn(ctx, y, f"{p}.blocks.{i}.1")))
    g = F.adaptive_avg_pool2d(x, 1)
    g = F.relu(_bn(ctx, F.conv2d(g, sd[f"{p}.pool.1.weight"]), f"{p}.pool.2"))
    feats.append(F.interpolate(g, (h, w), mode="bilinear", align_corners=False))
    y = F.conv2d(torch.cat(feats, 1), sd[f"{p}.project.0.weight"])
    return F.relu(_bn(ctx, y, f"{p}.project.1"))      # Dropout(0.1) is the identity in eval


def _gate_pipeline(ctx, g, x, p):
    """x * sigmoid(BN(psi(relu(BN(Wg g) + BN(Wx x)))))  (attention_aspp_unet_pipeline_stage.py:85-92)."""
    sd = ctx.sd
    a = _bn(ctx, F.conv2d(g, sd[p + ".Wg.0.weight"]), p + ".Wg.1")
    b = _bn(ctx, F.conv2d(x, sd[p + ".Wx.0.weight"]), p + ".Wx.1")
    s = _bn(ctx, F.conv2d(F.relu(a + b), sd[p + ".psi.0.weight"]), p + ".psi.1")
    psi = torch.sigmoid(s)
    return x * psi, psi


def _gate_ablation(ctx, g, x, p):
    """a = sigmoid(conv1x1_bias(relu(Wg g + Wx x))); returns (x*a + x, a)  (test_ablation.py:128-143)."""
    sd = ctx.sd
    s = F.conv2d(g, sd[p + ".Wg.weight"]) + F.conv2d(x, sd[p + ".Wx.weight"])
    a = torch.sigmoid(F.conv2d(F.relu(s), sd[p + ".psi.1.weight"], sd[p + ".psi.1.bias"]))
    return x * a + x, a


def _up(ctx, cfg: NetCfg, g, x, lvl: int, taps=None):
    """UpBlock (attention_aspp_unet_pipeline_stage.py:98-109): convT 2x2 s2 (+bias), bilinear fix-up when the
    floor-pooled size differs, gate, cat([x_att, g]) -- skip first -- then two ConvBNReLU."""
    sd = ctx.sd
    p = f"u{lvl}"
    g = F.conv_transpose2d(g, sd[p + ".up.weight"], sd[p + ".up.bias"], stride=2)
    if g.shape[-2:] != x.shape[-2:]:
        g = F.interpolate(g, size=x.shape[-2:], mode="bilinear", align_corners=False)
    psi = None
    if lvl in cfg.gate_levels():
        if cfg.variant == "pipeline":
            x, psi = _gate_pipeline(ctx, g, x, p + ".att")
        else:
            x, psi = _gate_ablation(ctx, g, x, p + ".att")
    if taps is not None:
        taps[f"g{lvl}"] = g
        taps[f"xatt{lvl}"] = x
    y = _cbr(ctx, torch.cat([x, g], 1), p + ".conv.0")
    if taps is not None:
        taps[f"u{lvl}a"] = y
    return _cbr(ctx, y, p + ".conv.1"), psi


def forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, cfg: NetCfg = NetCfg(), calibrate: bool = False,
            taps: Optional[dict] = None):
    """AttentionASPPUNet.forward (attention_aspp_unet_pipeline_stage.py:123-127 / test_ablation.py:205-218).

    ``x`` is fp32 ``[B, in_channels, H, W]``.  Returns logits ``[B, num_classes, H, W]`` for the pipeline
    variant and ``(logits, [psi3, psi2])`` for the ablation variant (disabled gates give ``zeros(1,1,1,1)``).
    ``taps`` (optional dict) receives the intermediate tensors by name for layer-by-layer checks.
    """
    ctx = _Ctx(sd, calibrate)
    with torch.no_grad():
        x = x.float()
        skips = []
        h = x
        for lvl in range(1, 5):
            if lvl > 1:
                h = F.max_pool2d(h, 2)
            h = _cbr(ctx, h, f"d{lvl}.0")
            if taps is not None:
                taps[f"d{lvl}.0"] = h
            h = _cbr(ctx, h, f"d{lvl}.1")
            skips.append(h)
            if taps is not None:
                taps[f"x{lvl}"] = h
        h = F.max_pool2d(h, 2)
        if taps is not None:
            taps["p4"] = h
        if cfg.variant == "pipeline" or cfg.use_aspp:
            h = _aspp(ctx, h)
        else:
            h = _cbr(ctx, h, "bridge.0")              # + Dropout(0.1): identity in eval (test_ablation.py:194-197)
        if taps is not None:
            taps["bridge"] = h
        psis = {}
        for lvl in (4, 3, 2, 1):
            h, psi = _up(ctx, cfg, h, skips[lvl - 1], lvl, taps)
            psis[lvl] = psi
            if taps is not None:
                taps[f"u{lvl}"] = h
        logits = F.conv2d(h, sd["out_conv.weight"], sd["out_conv.bias"])
    if cfg.variant == "pipeline":
        return logits
    z = torch.zeros(1, 1, 1, 1)
    return logits, [psis[4] if psis[4] is not None else z, psis[3] if psis[3] is not None else z]


def calibrate_bn(sd: Dict[str, torch.Tensor], x: torch.Tensor, cfg: NetCfg = NetCfg()) -> Dict[str, torch.Tensor]:
    """Fill every BN's running stats from one batch (regime R1).  Equivalent to running the reference module
    once on ``x`` with every ``BatchNorm2d`` in ``train()`` mode and ``momentum=None`` (Dropout left in eval) and
    then calling ``eval()``; ``oracle/gen_golden.py`` asserts that equivalence against the real module.  The one
    exception is the BN after the ASPP global pooling (``bridge.pool.2``), which keeps identity statistics."""
    sd = dict(sd)
    forward(sd, x, cfg, calibrate=True)
    return sd


def predict_prob_tta(sd, x, cfg: NetCfg = NetCfg()) -> torch.Tensor:
    """sigmoid((net(x) + flip(net(flip(x)))) / 2)  (attention_aspp_unet_pipeline_stage.py:336-338)."""
    def net(t):
        out = forward(sd, t, cfg)
        return out if cfg.variant == "pipeline" else out[0]
    return torch.sigmoid((net(x) + torch.flip(net(torch.flip(x, [-1])), [-1])) / 2)


# --------------------------------------------------------------------------------------------
# selection head (integer work; numpy / scipy exactly as the reference)
# --------------------------------------------------------------------------------------------
def frame_areas(prob: np.ndarray, thr: float = 0.05) -> np.ndarray:
    """Per-frame count of pixels with prob > thr (model_attention_aspp.py:71,74)."""
    return (prob > thr).astype(np.uint8).sum((1, 2))


def postprocess(prob: np.ndarray, thr: float = 0.05) -> np.ndarray:
    """FetalAbdomenSegmentation.postprocess (model_attention_aspp.py:69-89)."""
    import scipy.ndimage as ndi
    bin_ = (prob > thr).astype(np.uint8)
    frame_idx = int(bin_.sum((1, 2)).argmax())
    if bin_[frame_idx].sum() == 0:
        return np.zeros_like(bin_, np.uint8)
    frame = bin_[frame_idx]
    structure = np.ones((3, 3), dtype=np.uint8)
    frame = ndi.binary_dilation(frame, structure=structure, iterations=1)
    labeled, n = ndi.label(frame, structure=structure)
    if n:
        sizes = ndi.sum(frame, labeled, index=range(1, n + 1))
        frame = (labeled == (np.argmax(sizes) + 1)).astype(np.uint8)
    mask = np.zeros_like(bin_, np.uint8)
    mask[frame_idx] = frame
    return mask


def select_fetal_abdomen_mask_and_frame(mask_3d: np.ndarray):
    """model_attention_aspp.py:91-97."""
    if mask_3d.ndim == 2:
        return (mask_3d > 0).astype(np.uint8), 0
    areas = mask_3d.sum((1, 2))
    idx = int(areas.argmax())
    if areas[idx] == 0:
        return np.zeros(mask_3d.shape[1:], np.uint8), -1
    return (mask_3d[idx] > 0).astype(np.uint8), idx


def convert_2d_mask_to_3d(mask_2d: np.ndarray, frame_number: int, number_of_frames: int) -> np.ndarray:
    """inference.py:257-273: 1->2 relabel, all-zero volume when frame_number == -1, ValueError out of range."""
    mask_2d = np.where(mask_2d == 1, 2, 0).astype(np.uint8)
    vol = np.zeros((number_of_frames, mask_2d.shape[0], mask_2d.shape[1]), dtype=np.uint8)
    if frame_number == -1:
        return vol
    if frame_number is not None and 0 <= frame_number < number_of_frames:
        vol[frame_number] = mask_2d
        return vol
    raise ValueError("frame_number out of range")


# --------------------------------------------------------------------------------------------
# wrapper path around the network: conditioning, ROI-224 predict, output volume
# --------------------------------------------------------------------------------------------
def condition_frames(frames: np.ndarray) -> np.ndarray:
    """model_attention_aspp.py:11-17: min-max -> uint8, CLAHE(1.0, 8x8), median 3, /255 (float32 [N,H,W])."""
    import cv2
    clahe = cv2.createCLAHE(clipLimit=1.0, tileGridSize=(8, 8))
    stack = [cv2.medianBlur(clahe.apply(cv2.normalize(sl, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)), 3) for sl in frames]
    return np.stack(stack).astype(np.float32) / 255.0


def crop_roi_224(img: np.ndarray):
    """model_attention_aspp.py:20-30."""
    import cv2
    h, w = img.shape
    thr = img.mean() * 1.2
    ys, xs = np.where(img > thr)
    cx, cy = (w // 2, h // 2) if len(xs) == 0 else (int(xs.mean()), int(ys.mean()))
    x0, y0 = max(0, cx - 112), max(0, cy - 112)
    x0, y0 = min(x0, w - 224), min(y0, h - 224)
    patch = img[y0:y0 + 224, x0:x0 + 224]
    if patch.shape != (224, 224):
        patch = cv2.copyMakeBorder(patch, 0, 224 - patch.shape[0], 0, 224 - patch.shape[1], cv2.BORDER_CONSTANT, value=0)
    return patch, (x0, y0)


def predict_roi224(sd, frames01: np.ndarray, cfg: NetCfg = NetCfg(base_c=16)) -> np.ndarray:
    """FetalAbdomenSegmentation.predict after the file read (model_attention_aspp.py:44-60): float32 [128,H,W]."""
    import cv2
    vol = frames01[None]
    idxs = np.linspace(0, vol.shape[1] - 1, 128).astype(int)
    vol = vol[:, idxs]
    N, H, W = vol.shape[1:]
    patches, coords = [], []
    for sl in vol[0]:
        p, xy = crop_roi_224(sl)
        patches.append(p)
        coords.append(xy)
    tensor = torch.from_numpy(np.stack(patches)).unsqueeze(1)
    with torch.no_grad():
        outs = [torch.sigmoid(forward(sd, tensor[i:i + 8], cfg)).squeeze(1) for i in range(0, N, 8)]
    prob_roi = torch.cat(outs).numpy()
    prob_full = np.zeros((N, H, W), np.float32)
    for i, (x0, y0) in enumerate(coords):
        h_roi, w_roi = min(224, H - y0), min(224, W - x0)
        prob_full[i, y0:y0 + h_roi, x0:x0 + w_roi] = cv2.resize(prob_roi[i], (w_roi, h_roi))
    return prob_full


def output_volume(mask_2d: np.ndarray, frame_number: int, number_of_frames: int) -> np.ndarray:
    """The uint8 volume write_array_as_image_file hands to SimpleITK (inference.py:221-230): relabel 1->2, place
    the frame, then `> 0.5 -> 1`."""
    vol = convert_2d_mask_to_3d(np.squeeze(mask_2d).astype(np.float32), frame_number, number_of_frames)
    return np.where(vol > 0.5, 1, 0).astype(np.uint8)


# --------------------------------------------------------------------------------------------
# pipeline CLI recipe
# --------------------------------------------------------------------------------------------
def label8(m: np.ndarray) -> np.ndarray:
    """skimage.measure.label(m) for a 2-D array: full (8-) connectivity.  skimage is absent; scipy's labelling with a 3x3
    structure yields the same partition (label numbering may differ, the callers only use component membership / sizes)."""
    import scipy.ndimage as ndi
    return ndi.label(m, structure=np.ones((3, 3), np.uint8))[0]


def refine_mask(m: np.ndarray) -> np.ndarray:
    """test_ablation.py:373-387."""
    import cv2
    from scipy.ndimage import binary_fill_holes
    if m.sum() == 0:
        return m
    lab = label8(m)
    cnt = np.bincount(lab.ravel())
    cnt[0] = 0
    keep = [i for i, c in enumerate(cnt) if c >= max(20, int(0.0015 * m.size))]
    if not keep:
        return np.zeros_like(m)
    m = (np.isin(lab, keep)).astype(np.uint8)
    lab2 = label8(m)
    bc = np.bincount(lab2.ravel())
    bc[0] = 0
    m = (lab2 == np.argmax(bc)).astype(np.uint8)
    k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (7, 7))
    return binary_fill_holes(cv2.morphologyEx(m, cv2.MORPH_CLOSE, k)).astype(np.uint8)


def circularity_score(mask: np.ndarray) -> float:
    """test_ablation.py:389-396."""
    import cv2
    cnts, _ = cv2.findContours(mask.astype(np.uint8), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    if not cnts:
        return 0.0
    c = max(cnts, key=cv2.contourArea)
    area = cv2.contourArea(c)
    peri = cv2.arcLength(c, True)
    return 0.0 if peri <= 1e-6 else 4 * np.pi * area / (peri ** 2)


def select_best(stack, topk: int = 5) -> int:
    """test_ablation.py:398-403."""
    if len(stack) == 0:
        return 0
    areas = np.array([(m > 0).sum() for m in stack])
    idx = areas.argsort()[::-1][: max(1, min(topk, len(areas)))]
    return int(max(idx, key=lambda i: circularity_score(stack[i])))


def measure_ac_mm(mask01: np.ndarray, spacing) -> float:
    """attention_aspp_unet_pipeline_stage.py:355-374."""
    import cv2
    import math
    cnts, _ = cv2.findContours(mask01.astype(np.uint8), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
    if not cnts:
        return 0.0
    c = max(cnts, key=cv2.contourArea)
    if len(c) >= 5:
        (_, _), (MA, ma), _ = cv2.fitEllipse(c)
        a, b = MA / 2 * spacing[0], ma / 2 * spacing[1]
        h = ((a - b) ** 2) / ((a + b) ** 2)
        return math.pi * (a + b) * (1 + 3 * h / (10 + math.sqrt(4 - 3 * h)))
    return cv2.arcLength(c, True) * float(sum(spacing) / 2)


def pipeline_slice_prob(sd, frame_u8: np.ndarray, cfg: NetCfg = NetCfg(), img_size: int = 512) -> np.ndarray:
    """One slice of the CLI loop up to the blurred probability map (attention_aspp_unet_pipeline_stage.py:487-497):
    condition, Resize(512) (albumentations -> cv2.resize, INTER_LINEAR), ToFloat(255), flip TTA, resize back, blur."""
    import cv2
    e = np.rint(condition_frames(frame_u8[None])[0] * 255).astype(np.uint8)
    x = torch.from_numpy(cv2.resize(e, (img_size, img_size), interpolation=cv2.INTER_LINEAR).astype(np.float32) / 255.0)[None, None]
    with torch.no_grad():
        prob = predict_prob_tta(sd, x, cfg)[0, 0].numpy()
    prob = cv2.resize(prob, frame_u8.shape[::-1])
    return cv2.GaussianBlur(prob, (5, 5), 0)


# --------------------------------------------------------------------------------------------
# synthetic ACOUSLIC-style sweep
# --------------------------------------------------------------------------------------------
def synthetic_sweep(n_frames: int = 840, h: int = 562, w: int = 744, seed: int = 2025, peak: int = 304) -> np.ndarray:
    """uint8 ``[n_frames, h, w]``: fan-shaped field of view with Rayleigh speckle, exact 0 outside, and a bright
    ellipse whose size varies smoothly with the frame index and peaks at ``peak``."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    ax, ay = (w - 1) / 2.0, -0.12 * h
    r = np.hypot(xx - ax, yy - ay)
    ang = np.arctan2(xx - ax, yy - ay)
    fan = (r > 0.18 * h) & (r < 1.08 * h) & (np.abs(ang) < 0.66)
    out = np.empty((n_frames, h, w), np.uint8)
    pk = min(peak, n_frames - 1)
    for i in range(n_frames):
        speckle = rng.rayleigh(24.0, size=(h, w)).astype(np.float32)
        t = (i - pk) / max(n_frames, 1)
        s = float(np.exp(-(t * 3.0) ** 2))
        a, b = (0.10 + 0.10 * s) * w, (0.09 + 0.11 * s) * h
        cx, cy = 0.5 * w + 0.05 * w * np.sin(i * 0.05), 0.55 * h + 0.04 * h * np.cos(i * 0.031)
        e = ((xx - cx) / a) ** 2 + ((yy - cy) / b) ** 2
        img = speckle + 90.0 * np.exp(-np.clip(e, 0, 30) ** 2) * (0.6 + 0.4 * s)
        out[i] = np.where(fan, np.clip(img, 0, 255), 0).astype(np.uint8)
    return out
