"""Golden vectors for the pipeline CLI recipe, produced by the REAL reference helpers
(build container only):  python oracle/gen_golden_pipeline.py

The reference functions run unmodified: `predict_prob_tta`, `refine_mask`, `_circularity_score`, `select_best`,
`measure_ac_mm` of test_ablation.py (the pipeline script's own `select_best` is broken) and the slice loop
of `predict` is replayed with the reference's own calls.  Two stand-ins: `skimage.measure.label` (absent) is given a
scipy implementation with skimage's default full connectivity -- that one function is therefore NOT pinned -- and
albumentations' Resize / ToFloat are the `cv2.resize(INTER_LINEAR)` / `/255` they wrap.
Outputs -> tests/golden/pipeline_recipe.npz.
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import aau_oracle as O  # noqa: E402
import gen_golden as G  # noqa: E402

CASE = dict(n_frames=12, h=281, w=372, seed=77, peak=6, base_c=16, thr=0.48)
BIAS_SHIFT = 0.08           # puts the 0.48 threshold inside the spread of the TTA probabilities of the seeded weights


def recipe_state_dict():
    cfg = O.NetCfg(base_c=CASE["base_c"], variant="ablation")
    sd = O.make_state_dict(cfg, 2025, "R1")
    sd = O.calibrate_bn(sd, torch.rand(2, 1, 256, 256, generator=torch.Generator().manual_seed(4)), cfg)
    sd["out_conv.bias"] = sd["out_conv.bias"] + BIAS_SHIFT
    return cfg, sd


def main():
    import cv2
    import scipy.ndimage as ndi
    pipe, abl, wrap = G.import_reference()
    abl.label = lambda m: ndi.label(m, structure=np.ones((3, 3), np.uint8))[0]      # skimage.measure.label stand-in (8-connectivity)
    cfg, sd = recipe_state_dict()
    net = abl.AttentionASPPUNet(in_channels=1, num_classes=1, base_c=CASE["base_c"])
    net.load_state_dict(sd, strict=True)
    net.eval()
    sweep = O.synthetic_sweep(CASE["n_frames"], CASE["h"], CASE["w"], seed=CASE["seed"], peak=CASE["peak"])
    clahe = cv2.createCLAHE(1.0, (8, 8))
    probs, preds = [], []
    for sl in sweep:                                              # attention_aspp_unet_pipeline_stage.py:487-501 / test_ablation.py:823-835
        sl_u8 = cv2.normalize(sl, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
        e = cv2.medianBlur(clahe.apply(sl_u8), 3)
        x = torch.from_numpy(cv2.resize(e, (512, 512), interpolation=cv2.INTER_LINEAR).astype(np.float32) / 255.0)[None, None]
        prob = abl.predict_prob_tta(net, x)
        prob = cv2.resize(prob, sl.shape[::-1])
        prob = cv2.GaussianBlur(prob, (5, 5), 0)
        probs.append(prob)
        preds.append(abl.refine_mask((prob > CASE["thr"]).astype(np.uint8)))
    preds = np.stack(preds)
    probs = np.stack(probs)
    bf = abl.select_best(preds, 5)
    ac = round(abl.measure_ac_mm(preds[bf], (0.28, 0.28)), 1)
    circ = np.array([abl._circularity_score(m) for m in preds])
    raw = (probs > CASE["thr"]).astype(np.uint8)
    out = ROOT / "tests" / "golden" / "pipeline_recipe.npz"
    np.savez_compressed(out, case=json.dumps(CASE), bias_shift=np.array(BIAS_SHIFT), prob_sub=probs[:, ::3, ::3].astype(np.float16),
                        raw_masks=np.packbits(raw), refined=np.packbits(preds), areas=preds.reshape(len(preds), -1).sum(1).astype(np.int64),
                        circularity=circ, best_frame=np.array(bf), ac_mm=np.array(ac))
    print("best frame", bf, "AC", ac, "areas", preds.reshape(len(preds), -1).sum(1), "circ", np.round(circ, 3), "prob mean/std", probs.mean(), probs.std(),
          "->", out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
