"""The pipeline CLI recipe: flip-TTA probabilities, refine_mask, select_best, AC in mm.
Golden vectors: tests/golden/pipeline_recipe.npz from oracle/gen_golden_pipeline.py, i.e. the reference's own
predict_prob_tta / refine_mask / select_best / measure_ac_mm (test_ablation.py) on a seeded synthetic sweep."""
import json

import numpy as np
import pytest
import torch

import aau_oracle as O
from conftest import GOLDEN


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN / "pipeline_recipe.npz")


@pytest.fixture(scope="module")
def case(gold):
    c = json.loads(str(gold["case"]))
    sweep = O.synthetic_sweep(c["n_frames"], c["h"], c["w"], seed=c["seed"], peak=c["peak"])
    cfg = O.NetCfg(base_c=c["base_c"], variant="ablation")
    sd = O.calibrate_bn(O.make_state_dict(cfg, 2025, "R1"), torch.rand(2, 1, 256, 256, generator=torch.Generator().manual_seed(4)), cfg)
    sd["out_conv.bias"] = sd["out_conv.bias"] + float(gold["bias_shift"])
    return c, sweep, cfg, sd


def unpack(bits, c):
    n, h, w = c["n_frames"], c["h"], c["w"]
    return np.unpackbits(bits)[: n * h * w].reshape(n, h, w)


# ------------------------------------------------------------------------------------------------ CPU
def test_oracle_slice_probability_matches_reference(gold, case):
    c, sweep, cfg, sd = case
    for i in (0, 7):
        prob = O.pipeline_slice_prob(sd, sweep[i], cfg)
        np.testing.assert_allclose(prob[::3, ::3], gold["prob_sub"][i].astype(np.float32), atol=3e-4, rtol=2e-3)


def test_host_tail_is_bit_exact_on_reference_masks(gold, case):
    """refine_mask / select_best / measure_ac_mm of the oracle AND of the product package on the reference's raw masks."""
    c = case[0]
    import pipeline_predict as PP
    raw, refined = unpack(gold["raw_masks"], c), unpack(gold["refined"], c)
    for impl in (O, PP):
        got = np.stack([impl.refine_mask(m.copy()) for m in raw])
        assert np.array_equal(got, refined)
        assert np.array_equal(got.reshape(len(got), -1).sum(1), gold["areas"])
        assert impl.select_best(got, 5) == int(gold["best_frame"])
        assert round(impl.measure_ac_mm(got[int(gold["best_frame"])], (0.28, 0.28)), 1) == float(gold["ac_mm"])
    np.testing.assert_allclose([PP._circularity_score(m) for m in refined], gold["circularity"], rtol=1e-12)
    assert PP.select_best([], 5) == 0 and PP.measure_ac_mm(np.zeros((8, 8), np.uint8), (1, 1)) == 0.0
    assert not PP.refine_mask(np.zeros((16, 16), np.uint8)).any()
    speck = np.zeros((200, 200), np.uint8)
    speck[5:8, 5:8] = 1                                              # 9 pixels < max(20, 0.15 %): removed
    assert not PP.refine_mask(speck).any()
    vol = PP.convert_mask_2d_to_3d(refined[0], 3, 5)
    assert set(np.unique(vol)) == {0, 2} and vol[3].sum() == 2 * refined[0].sum() and not PP.convert_mask_2d_to_3d(refined[0], 9, 5).any()


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_engine_tta_and_case_prediction(gold, case, tmp_path):
    c, sweep, cfg, sd = case
    import metaimage
    import pipeline_predict as PP
    from attention_aspp_unet import AttentionASPPUNet
    net = AttentionASPPUNet(base_c=c["base_c"], use_att=True, use_aspp=True, att_depth=4, act_dtype="fp16")
    net.load_state_dict(sd, strict=True)
    net.eval()
    pp = PP.PipelinePredictor(net, batch=8)
    # flip TTA against the oracle on one batch (float input, as the reference feeds it)
    x = torch.rand(3, 1, 96, 128, generator=torch.Generator().manual_seed(1))
    ref = O.predict_prob_tta(sd, x, cfg)[:, 0]
    got = pp.predict_prob_tta(x.cuda()).cpu()
    assert (got - ref).abs().max().item() < 2e-3
    flipped = pp.predict_prob_tta(torch.flip(x, [-1]).cuda()).cpu()
    assert (torch.flip(flipped, [-1]) - got).abs().max().item() < 1e-3          # TTA output is flip-equivariant
    # device head / tail of the slice loop against the OpenCV calls they replace
    import cv2
    rng = np.random.default_rng(3)
    for (sh, sw, dh, dw) in ((281, 372, 512, 512), (562, 744, 512, 512), (97, 131, 512, 512), (600, 800, 512, 512), (512, 512, 300, 200)):
        img = rng.integers(0, 256, (2, sh, sw), dtype=np.uint8)
        got = pp.resize_on_device(torch.from_numpy(img).cuda(), (dh, dw)).cpu().numpy()
        want = np.stack([cv2.resize(f, (dw, dh), interpolation=cv2.INTER_LINEAR) for f in img])
        assert np.array_equal(got, want), (sh, sw, dh, dw)           # cv2.resize(uint8, INTER_LINEAR): bit exact
    for (H, W) in ((281, 372), (562, 744), (97, 131), (33, 31)):
        prob = torch.sigmoid(torch.from_numpy(rng.normal(0, 2, (3, 64, 64)).astype(np.float32)))
        prob = torch.nn.functional.interpolate(prob[None], size=(512, 512), mode="bicubic", align_corners=False)[0].clamp(0, 1).contiguous()
        m, a = pp.tail_on_device(prob.cuda(), (H, W), c["thr"])
        blurred = np.stack([cv2.GaussianBlur(cv2.resize(p_, (W, H)), (5, 5), 0) for p_ in prob.numpy()])
        want = (blurred > c["thr"]).astype(np.uint8)
        near = np.abs(blurred - c["thr"]) < 2e-6                     # OpenCV's own last-bit summation order decides these
        assert np.array_equal(m.cpu().numpy()[~near], want[~near]) and near.mean() < 1e-4, (H, W)
        assert np.array_equal(a.cpu().numpy(), m.cpu().numpy().reshape(3, -1).sum(1))
    # whole recipe on the sweep: blurred probabilities, refined masks, best frame, AC
    raw, raw_areas = pp.raw_masks(sweep, c["thr"])
    gold_raw = unpack(gold["raw_masks"], c)
    assert (raw == gold_raw).mean() >= 0.998 and np.array_equal(raw_areas, raw.reshape(len(raw), -1).sum(1))
    masks = pp.predict_masks(sweep, c["thr"])
    refined = unpack(gold["refined"], c)
    assert masks.shape == refined.shape and masks.dtype == np.uint8
    assert (masks == refined).mean() >= 0.995
    areas = masks.reshape(len(masks), -1).sum(1)
    assert np.abs(areas - gold["areas"]).max() <= 0.01 * gold["areas"].max()
    metaimage.write_mha(tmp_path / "caseA.mha", sweep, spacing=(0.28, 0.28, 0.28))
    cv2.imwrite(str(tmp_path / "caseB_s0003.png"), sweep[3])
    (tmp_path / "sp.json").write_text(json.dumps({"caseB": {"spacing": [0.3, 0.3]}}))
    rows = PP.predict(tmp_path, tmp_path / "out", net=net, spacing_json=str(tmp_path / "sp.json"), thr=c["thr"], batch=8)
    assert [r[0] for r in rows] == ["caseA", "caseB"] and rows[1][1] == 3
    vol, hdr = metaimage.read_mha(tmp_path / "out/caseA/images/fetal-abdomen-segmentation/output.mha")
    bf = json.loads((tmp_path / "out/caseA/fetal-abdomen-frame-number.json").read_text())
    assert vol.shape == sweep.shape and set(np.unique(vol)) <= {0, 2} and np.flatnonzero(vol.reshape(len(vol), -1).any(1)).tolist() == [bf]
    assert abs(rows[0][2] - float(gold["ac_mm"])) <= 0.02 * float(gold["ac_mm"])          # AC within 2 % of the reference's
    assert (tmp_path / "out/ac_results.csv").read_text().splitlines()[0] == "case_id,frame_idx,ac_mm"
    assert (tmp_path / "out/caseB_s0003_mask.png").exists()
