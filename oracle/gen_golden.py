"""Generate tests/golden/*.npz by running the REAL reference modules from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python oracle/gen_golden.py

For every case: the state dict comes from ``aau_oracle.make_state_dict`` (seeded, reproducible anywhere), is
loaded into the unmodified reference ``AttentionASPPUNet`` with ``strict=True`` (this pins the state_dict
layout of SURVEY.md §8 a9), the reference forward is run in fp32 on CPU and its outputs are stored.  For
regime R1 the reference module itself is put in ``train()`` with ``momentum=None`` for one batch and the
resulting running statistics are compared with ``aau_oracle.calibrate_bn``.
The selection head goldens call the reference's ``postprocess`` / ``select_fetal_abdomen_mask_and_frame`` /
``convert_2d_mask_to_3d`` on seeded random probability volumes.
"""
import json
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("AAU_REFERENCE", "/root/reference"))
sys.path.insert(0, str(ROOT / "oracle"))
import aau_oracle as O  # noqa: E402


def import_reference():
    """Import the reference with inert stand-ins for the third-party modules this image lacks."""
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    dummy = lambda *a, **k: None  # noqa: E731
    names = ["CLAHE", "Compose", "HorizontalFlip", "MedianBlur", "RandomBrightnessContrast", "RandomGamma", "Resize",
             "ToFloat", "ShiftScaleRotate", "GaussNoise", "ElasticTransform", "GridDistortion", "OneOf", "Normalize",
             "VerticalFlip", "Rotate", "RandomResizedCrop", "GaussianBlur", "Affine", "CoarseDropout", "PadIfNeeded"]
    for n in ("albumentations", "skimage", "SimpleITK"):
        if n not in sys.modules:
            try:
                __import__(n)
            except ImportError:
                stub(n, **({k: dummy for k in names} if n == "albumentations" else {}))
    if "albumentations.pytorch" not in sys.modules:
        try:
            __import__("albumentations.pytorch")
        except ImportError:
            stub("albumentations.pytorch", ToTensorV2=dummy)
    if "skimage.measure" not in sys.modules:
        try:
            __import__("skimage.measure")
        except ImportError:
            stub("skimage.measure", label=dummy, regionprops=dummy)
    sys.path.insert(0, str(REF))
    import attention_aspp_unet_pipeline_stage as pipe
    try:
        import test_ablation as abl
    except Exception as e:  # pragma: no cover
        # a missing stub name: add it on the fly and retry once
        print("retry import test_ablation:", e)
        raise
    # the wrapper imports a module the reference never ships; give it the real class under that name
    stub("attention_aspp_unet", AttentionASPPUNet=pipe.AttentionASPPUNet)
    os.environ.setdefault("MODEL_TAG", "att_aspp")
    import model_attention_aspp as wrap
    return pipe, abl, wrap


def ref_model(pipe, abl, cfg: O.NetCfg):
    if cfg.variant == "pipeline":
        return pipe.AttentionASPPUNet(in_channels=cfg.in_channels, num_classes=cfg.num_classes, base_c=cfg.base_c)
    return abl.AttentionASPPUNet(in_channels=cfg.in_channels, num_classes=cfg.num_classes, base_c=cfg.base_c,
                                 use_att=cfg.use_att, use_aspp=cfg.use_aspp, att_depth=cfg.att_depth)


def case_input(seed, b, h, w, kind="rand"):
    if kind == "rand":
        return torch.rand(b, 1, h, w, generator=torch.Generator().manual_seed(seed))
    vol = O.synthetic_sweep(b, h, w, seed=seed, peak=b // 2)
    return torch.from_numpy(vol.astype(np.float32) / 255.0).unsqueeze(1)


CASES = [
    # name, cfg, regime, (B,H,W), input kind, store stride
    ("pipe_c16_R0_64x80", O.NetCfg(base_c=16), "R0", (2, 64, 80), "rand", 1),
    ("pipe_c16_R1_141x93", O.NetCfg(base_c=16), "R1", (2, 141, 93), "rand", 1),
    ("pipe_c32_R1_64x64", O.NetCfg(base_c=32), "R1", (1, 64, 64), "rand", 1),
    ("pipe_c32_R1_562x744", O.NetCfg(base_c=32), "R1", (1, 562, 744), "sweep", 7),
    ("abl_full_c16_R1_80x72", O.NetCfg(base_c=16, variant="ablation"), "R1", (2, 80, 72), "rand", 1),
    ("abl_noatt_c16_R1_80x72", O.NetCfg(base_c=16, variant="ablation", use_att=False), "R1", (1, 80, 72), "rand", 1),
    ("abl_noaspp_c16_R1_80x72", O.NetCfg(base_c=16, variant="ablation", use_aspp=False), "R1", (1, 80, 72), "rand", 1),
    ("abl_neither_c16_R1_80x72", O.NetCfg(base_c=16, variant="ablation", use_att=False, use_aspp=False), "R1",
     (1, 80, 72), "rand", 1),
    ("abl_depth3_c16_R1_81x73", O.NetCfg(base_c=16, variant="ablation", att_depth=3), "R1", (1, 81, 73), "rand", 1),
]


def main():
    pipe, abl, wrap = import_reference()
    out_dir = ROOT / "tests" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    manifest = {}
    for name, cfg, regime, (b, h, w), kind, stride in CASES:
        seed = 2025
        sd = O.make_state_dict(cfg, seed=seed, regime=regime)
        model = ref_model(pipe, abl, cfg)
        ref_keys = [(k, list(v.shape)) for k, v in model.state_dict().items()]
        assert [k for k, _ in ref_keys] == list(sd.keys()), f"{name}: key order/layout differs from the reference"
        model.load_state_dict(sd, strict=True)
        x = case_input(seed + 1, b, h, w, kind)
        if regime == "R1":
            xc = case_input(seed + 2, max(b, 2), h, w, kind)
            model.eval()                      # Dropout stays off; only the BN layers collect statistics
            for mn, m in model.named_modules():
                if isinstance(m, torch.nn.BatchNorm2d) and mn != "bridge.pool.2":   # 1x1 map: see aau_oracle._bn
                    m.momentum = None
                    m.train()
            with torch.no_grad():
                model(xc)
            model.eval()
            sd_cal = O.calibrate_bn(sd, xc, cfg)
            worst = 0.0
            for k, v in model.state_dict().items():
                if v.dtype.is_floating_point:
                    d = (v - sd_cal[k]).abs().max().item() / (v.abs().max().item() + 1e-6)
                    worst = max(worst, d)
            assert worst < 2e-4, f"{name}: oracle BN calibration differs from the reference (rel {worst})"
            # use the oracle's own calibrated stats everywhere (reproducible without the reference)
            model.load_state_dict(sd_cal, strict=True)
            sd = sd_cal
        model.eval()
        with torch.no_grad():
            out = model(x)
        rec = {}
        if cfg.variant == "pipeline":
            logits = out
        else:
            logits, psis = out
            rec["psi3"] = psis[0].numpy()
            rec["psi2"] = psis[1].numpy()
        lg = logits.numpy()
        rec["logits"] = lg[:, :, ::stride, ::stride].copy()
        rec["logits_stats"] = np.array([lg.mean(), lg.std(), lg.min(), lg.max()], np.float64)
        np.savez_compressed(out_dir / f"{name}.npz", **rec)
        manifest[name] = {"cfg": cfg.__dict__, "regime": regime, "shape": [b, h, w], "input": kind, "seed": seed,
                          "stride": stride, "n_state_dict_entries": len(ref_keys),
                          "logits_mean_std_min_max": [float(v) for v in rec["logits_stats"]]}
        print(name, "entries", len(ref_keys), "logits mean/std/min/max", rec["logits_stats"])
        if name in ("pipe_c16_R0_64x80", "abl_full_c16_R1_80x72"):
            manifest[name]["state_dict_layout"] = ref_keys

    # ---- selection head (integer work, bit exact) ----
    rng = np.random.default_rng(7)
    sel = {}
    vols = {
        "random": rng.random((12, 40, 52), dtype=np.float32) * 0.06,
        "blobs": np.zeros((9, 48, 64), np.float32),
        "empty": np.zeros((5, 16, 16), np.float32),
        "ties": np.zeros((6, 16, 16), np.float32),
    }
    yy, xx = np.mgrid[0:48, 0:64]
    for i in range(9):
        r = 4 + 2 * min(i, 8 - i)
        vols["blobs"][i][(yy - 20) ** 2 + (xx - 30) ** 2 < r * r] = 0.9
        vols["blobs"][i][(yy - 40) ** 2 + (xx - 8) ** 2 < 9] = 0.5       # a second, smaller component
    vols["ties"][2, :4, :4] = 1.0
    vols["ties"][4, 8:12, 8:12] = 1.0                                       # same area: first index must win
    algo = object.__new__(wrap.FetalAbdomenSegmentation)                    # postprocess() uses no instance state
    for k, v in vols.items():
        m3 = wrap.FetalAbdomenSegmentation.postprocess(algo, v)
        m2, idx = wrap.select_fetal_abdomen_mask_and_frame(m3)
        sel[k + "_prob"] = v
        sel[k + "_mask3d"] = m3
        sel[k + "_mask2d"] = m2
        sel[k + "_idx"] = np.array(idx)
    np.savez_compressed(out_dir / "selection.npz", **sel)
    (out_dir / "manifest.json").write_text(json.dumps(manifest, indent=1, default=str))
    print("wrote", out_dir)


if __name__ == "__main__":
    main()
