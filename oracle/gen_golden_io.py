"""Golden vectors for the wrapper path AROUND the network, produced by the REAL reference
code in /root/reference (build container only):

    python oracle/gen_golden_io.py

`SimpleITK` is absent here, so a stand-in whose ReadImage / GetArrayFromImage hand back a seeded synthetic sweep is
installed before the reference modules are imported; everything downstream of the file read is the unmodified
reference: `load_image_file_as_array` (cv2 normalize / CLAHE / median), `crop_roi_224`,
`FetalAbdomenSegmentation.predict` (128 sampled frames, batch 8, sigmoid, cv2 paste back), `postprocess`,
`select_fetal_abdomen_mask_and_frame`, `inference.convert_2d_mask_to_3d` and `inference.write_json_file`.
The network inside is the reference `AttentionASPPUNet(base_c=16)` holding the oracle's seeded R1 state dict whose
out_conv bias is moved to logit(0.05), so the wrapper's 0.05 threshold cuts through the probability range.
Outputs -> tests/golden/wrapper_io.npz (sub-sampled probabilities, ROI corners, masks, frame number, JSON text).
"""
import json
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import aau_oracle as O  # noqa: E402
import gen_golden as G  # noqa: E402

CASE = dict(n_frames=140, h=250, w=330, seed=31, peak=75)      # small sweep, larger than the 224 ROI in both axes
BIAS_SHIFT = float(np.log(0.05 / 0.95))


def wrapper_state_dict():
    cfg = O.NetCfg(base_c=16)
    sd = O.make_state_dict(cfg, 2025, "R1")
    calib = torch.rand(2, 1, 224, 224, generator=torch.Generator().manual_seed(3))
    sd = O.calibrate_bn(sd, calib, cfg)
    sd["out_conv.bias"] = sd["out_conv.bias"] + BIAS_SHIFT
    return cfg, sd


def main():
    sweep = O.synthetic_sweep(CASE["n_frames"], CASE["h"], CASE["w"], seed=CASE["seed"], peak=CASE["peak"])
    sitk = types.ModuleType("SimpleITK")
    sitk.ReadImage = lambda path: ("image", str(path))
    sitk.GetArrayFromImage = lambda img: sweep
    sys.modules["SimpleITK"] = sitk
    pipe, abl, wrap = G.import_reference()
    import inference as ref_inf                                   # MODEL_TAG=att_aspp set by import_reference
    cfg, sd = wrapper_state_dict()
    net = pipe.AttentionASPPUNet(in_channels=1, num_classes=1, base_c=16)
    net.load_state_dict(sd, strict=True)
    net.eval()
    algo = object.__new__(wrap.FetalAbdomenSegmentation)           # __init__ wants a checkpoint file and kwargs the class rejects
    algo.device = torch.device("cpu")
    algo.net = net
    cond = wrap.load_image_file_as_array(location=Path("synthetic.mha"))      # (1, N, H, W)
    prob = algo.predict(["synthetic.mha"], save_probabilities=False)
    idxs = np.linspace(0, CASE["n_frames"] - 1, 128).astype(int)
    coords = np.array([wrap.crop_roi_224(sl)[1] for sl in cond[0][idxs]], np.int32)
    post = algo.postprocess(prob)
    mask2d, frame = wrap.select_fetal_abdomen_mask_and_frame(post)
    vol3d = ref_inf.convert_2d_mask_to_3d(mask_2d=mask2d.astype(np.float32), frame_number=frame, number_of_frames=CASE["n_frames"])
    final = np.where(vol3d > 0.5, 1, 0).astype(np.uint8)
    with tempfile.TemporaryDirectory() as d:
        ref_inf.write_json_file(location=Path(d) / "f.json", content=frame)
        json_text = (Path(d) / "f.json").read_text()
    areas = (prob > 0.05).astype(np.uint8).sum((1, 2))
    out = ROOT / "tests" / "golden" / "wrapper_io.npz"
    assert frame >= 0 and np.array_equal(final[frame], mask2d) and np.array_equal(post[frame], mask2d)
    np.savez_compressed(out, case=json.dumps(CASE), cond_frames_u8=np.rint(cond[0][[0, 70, 139]] * 255).astype(np.uint8), coords=coords,
                        prob_sub=prob[::16, ::3, ::3].astype(np.float16), prob_stats=np.array([prob.mean(), prob.std(), prob.max()]),
                        areas=areas.astype(np.int64), bin_best=np.packbits(prob[frame] > 0.05), mask2d=np.packbits(mask2d),
                        frame=np.array(frame), final_nonzero_frames=np.flatnonzero(final.reshape(final.shape[0], -1).any(1)),
                        json_text=np.array(json_text))
    print("frame", frame, "area", int(mask2d.sum()), "areas range", areas.min(), areas.max(), "prob mean/std/max", prob.mean(), prob.std(), prob.max(),
          "json", repr(json_text), "->", out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
