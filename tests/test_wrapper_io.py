"""The wrapper path around the network: frame conditioning, ROI-224 predict, mask
volume + frame-number JSON.  Golden vectors: tests/golden/wrapper_io.npz, produced by oracle/gen_golden_io.py from
the REAL reference wrapper (model_attention_aspp.py, inference.py) on a seeded synthetic sweep."""
import json
import sys
import zlib

import numpy as np
import pytest
import torch

import aau_oracle as O
from conftest import GOLDEN

BIAS_SHIFT = float(np.log(0.05 / 0.95))


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN / "wrapper_io.npz")


@pytest.fixture(scope="module")
def case(gold):
    c = json.loads(str(gold["case"]))
    sweep = O.synthetic_sweep(c["n_frames"], c["h"], c["w"], seed=c["seed"], peak=c["peak"])
    return c, sweep


def wrapper_state_dict():
    cfg = O.NetCfg(base_c=16)
    sd = O.make_state_dict(cfg, 2025, "R1")
    sd = O.calibrate_bn(sd, torch.rand(2, 1, 224, 224, generator=torch.Generator().manual_seed(3)), cfg)
    sd["out_conv.bias"] = sd["out_conv.bias"] + BIAS_SHIFT
    return cfg, sd


# ------------------------------------------------------------------------------------------------ CPU: oracle
def test_oracle_conditioning_and_roi_match_reference(gold, case):
    c, sweep = case
    cond = O.condition_frames(sweep)
    assert np.array_equal(np.rint(cond[[0, 70, 139]] * 255).astype(np.uint8), gold["cond_frames_u8"])   # cv2: bit exact
    idxs = np.linspace(0, c["n_frames"] - 1, 128).astype(int)
    coords = np.array([O.crop_roi_224(sl)[1] for sl in cond[idxs]], np.int32)
    assert np.array_equal(coords, gold["coords"])


def test_oracle_predict_matches_reference(gold, case):
    c, sweep = case
    cfg, sd = wrapper_state_dict()
    prob = O.predict_roi224(sd, O.condition_frames(sweep), cfg)
    assert prob.shape == (128, c["h"], c["w"]) and prob.dtype == np.float32
    np.testing.assert_allclose(prob[::16, ::3, ::3], gold["prob_sub"].astype(np.float32), atol=2e-4, rtol=2e-3)
    areas = O.frame_areas(prob, 0.05)
    assert np.abs(areas - gold["areas"]).max() <= 4                  # fp32 reduction-order noise at the threshold only
    post = O.postprocess(prob)
    mask2d, frame = O.select_fetal_abdomen_mask_and_frame(post)
    assert frame == int(gold["frame"])
    g2d = np.unpackbits(gold["mask2d"])[: c["h"] * c["w"]].reshape(c["h"], c["w"])
    assert (mask2d != g2d).mean() < 1e-4


def test_oracle_integer_tail_is_bit_exact(gold, case):
    c, _ = case
    import scipy.ndimage as ndi
    bin_best = np.unpackbits(gold["bin_best"])[: c["h"] * c["w"]].reshape(1, c["h"], c["w"]).astype(np.float32)
    post = O.postprocess(bin_best, thr=0.5)                          # the reference's integer tail on the reference's own mask
    g2d = np.unpackbits(gold["mask2d"])[: c["h"] * c["w"]].reshape(c["h"], c["w"])
    assert np.array_equal(post[0], g2d)
    vol = O.output_volume(g2d, int(gold["frame"]), c["n_frames"])
    assert vol.dtype == np.uint8 and set(np.unique(vol)) <= {0, 1}
    assert np.array_equal(np.flatnonzero(vol.reshape(vol.shape[0], -1).any(1)), gold["final_nonzero_frames"])
    assert np.array_equal(vol[int(gold["frame"])], g2d)
    assert not O.output_volume(g2d, -1, 5).any()
    with pytest.raises(ValueError):
        O.convert_2d_mask_to_3d(g2d, c["n_frames"], c["n_frames"])


# ------------------------------------------------------------------------------------------------ CPU: host modules
def test_metaimage_round_trip_and_header(tmp_path):
    import metaimage
    rng = np.random.default_rng(5)
    vol = (rng.random((7, 19, 23)) > 0.8).astype(np.uint8)
    path = tmp_path / "m.mha"
    n = metaimage.write_mha(path, vol, spacing=(0.28, 0.28, 0.28), compress=True)
    raw = path.read_bytes()
    assert len(raw) == n
    head, _, payload = raw.partition(b"ElementDataFile = LOCAL\n")
    fields = dict(l.split(" = ", 1) for l in head.decode().strip().splitlines())
    assert fields["ObjectType"] == "Image" and fields["NDims"] == "3" and fields["DimSize"] == "23 19 7"
    assert fields["ElementType"] == "MET_UCHAR" and fields["ElementSpacing"] == "0.28 0.28 0.28"
    assert fields["CompressedData"] == "True" and int(fields["CompressedDataSize"]) == len(payload)
    assert zlib.decompress(payload) == vol.tobytes()                 # x fastest, then y, then frames
    back, hdr = metaimage.read_mha(path)
    assert back.dtype == np.uint8 and np.array_equal(back, vol)
    for dt in (np.int16, np.uint16, np.float32):
        a = (rng.random((3, 5, 6)) * 100).astype(dt)
        metaimage.write_mha(tmp_path / "a.mha", a, compress=False)
        b, _ = metaimage.read_mha(tmp_path / "a.mha")
        assert b.dtype == dt and np.array_equal(a, b)
    (tmp_path / "bad.mha").write_bytes(b"ObjectType = Image\nNDims = 3\n")
    with pytest.raises(metaimage.MetaImageError):
        metaimage.read_mha(tmp_path / "bad.mha")


def test_host_glue_matches_reference_semantics(gold, case, tmp_path):
    """convert_2d_mask_to_3d / write_array_as_image_file / write_json_file of the product package (no GPU needed)."""
    c, sweep = case
    import fetal_abdomen as FA
    import metaimage
    import inference as INF                                           # the package's inference.py (att-aspp-unet_b200 is first on sys.path)
    assert INF.__file__.endswith("att-aspp-unet_b200/inference.py")
    g2d = np.unpackbits(gold["mask2d"])[: c["h"] * c["w"]].reshape(c["h"], c["w"])
    frame = int(gold["frame"])
    v = INF.convert_2d_mask_to_3d(mask_2d=g2d.astype(np.float32), frame_number=frame, number_of_frames=c["n_frames"])
    assert np.array_equal(v, O.convert_2d_mask_to_3d(g2d.astype(np.float32), frame, c["n_frames"]))
    assert not INF.convert_2d_mask_to_3d(mask_2d=g2d, frame_number=-1, number_of_frames=4).any()
    for bad in (None, c["n_frames"], -2):
        with pytest.raises(ValueError):
            INF.convert_2d_mask_to_3d(mask_2d=g2d, frame_number=bad, number_of_frames=c["n_frames"])
    out = INF.write_array_as_image_file(location=tmp_path / "images/fetal-abdomen-segmentation", array=g2d, frame_number=frame,
                                        number_of_frames=c["n_frames"], filename="case7.mha")
    vol, hdr = metaimage.read_mha(out)
    assert np.array_equal(vol, O.output_volume(g2d, frame, c["n_frames"])) and hdr["ElementSpacing"] == "0.28 0.28 0.28"
    INF.write_json_file(location=tmp_path / "fetal-abdomen-frame-number.json", content=frame)
    assert (tmp_path / "fetal-abdomen-frame-number.json").read_text() == str(gold["json_text"])
    # conditioning + ROI of the product package against the reference's
    cond = FA.preprocess_sweep(sweep[[0, 70, 139]])
    assert np.array_equal(np.rint(cond * 255).astype(np.uint8), gold["cond_frames_u8"])
    full = FA.preprocess_sweep(sweep)
    idxs = np.linspace(0, c["n_frames"] - 1, 128).astype(int)
    assert np.array_equal(np.array([FA.crop_roi_224(sl)[1] for sl in full[idxs]], np.int32), gold["coords"])
    small = np.zeros((100, 150), np.float32)                          # smaller than the ROI: zero padded to 224x224
    small[40:60, 50:80] = 1.0
    p, _ = FA.crop_roi_224(np.pad(small, ((0, 130), (0, 80))))
    assert p.shape == (224, 224)


# ------------------------------------------------------------------------------------------------ GPU: engine
@pytest.mark.gpu
def test_engine_predict_and_run_match_reference(gold, case, tmp_path):
    c, sweep = case
    from attention_aspp_unet import AttentionASPPUNet
    from fetal_abdomen import FetalAbdomenSegmentation, select_fetal_abdomen_mask_and_frame
    import inference as INF
    import metaimage
    cfg, sd = wrapper_state_dict()
    g2d = np.unpackbits(gold["mask2d"])[: c["h"] * c["w"]].reshape(c["h"], c["w"])
    in_dir = tmp_path / "in/images/stacked-fetal-ultrasound"
    in_dir.mkdir(parents=True)
    metaimage.write_mha(in_dir / "sweep.mha", sweep, spacing=(0.28, 0.28, 0.28))
    for dtype, tol, agree in (("fp16", 2e-3, 0.999), ("bf16", 1.5e-2, 0.99)):
        net = AttentionASPPUNet(in_ch=1, num_classes=1, base=16, act_dtype=dtype)
        miss, unexp = net.load_state_dict(sd, strict=False)
        assert not miss and not unexp
        algo = FetalAbdomenSegmentation(net=net, batch=32)
        prob = algo.predict([str(in_dir / "sweep.mha")])
        assert prob.shape == (128, c["h"], c["w"]) and prob.dtype == np.float32
        err = np.abs(prob[::16, ::3, ::3] - gold["prob_sub"].astype(np.float32))
        assert err.max() < tol, f"{dtype}: probability volume differs from the reference by {err.max()}"
        post = algo.postprocess(prob)
        mask2d, frame = select_fetal_abdomen_mask_and_frame(post, _engine=algo)
        if dtype == "fp16":
            assert frame == int(gold["frame"])
            assert (mask2d == g2d).mean() >= agree
        out_dir = tmp_path / f"out_{dtype}"
        assert INF.run("caseX", algorithm=algo, input_path=tmp_path / "in", output_path=out_dir) == 0
        vol, hdr = metaimage.read_mha(out_dir / "images/fetal-abdomen-segmentation/caseX.mha")
        fr = json.loads((out_dir / "fetal-abdomen-frame-number.json").read_text())
        assert vol.shape == sweep.shape and vol.dtype == np.uint8 and set(np.unique(vol)) <= {0, 1}
        assert fr == frame and np.array_equal(np.flatnonzero(vol.reshape(vol.shape[0], -1).any(1)), [frame])
        assert np.array_equal(vol[frame], mask2d)
    # engine-native mode: every frame at full resolution
    algo.batch = 20
    assert INF.run("caseF", algorithm=algo, input_path=tmp_path / "in", output_path=tmp_path / "out_full", mode="full") == 0
    vol, _ = metaimage.read_mha(tmp_path / "out_full/images/fetal-abdomen-segmentation/caseF.mha")
    fr = json.loads((tmp_path / "out_full/fetal-abdomen-frame-number.json").read_text())
    assert vol.shape == sweep.shape and -1 <= fr < sweep.shape[0]


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(5, 562, 744), (3, 250, 330), (4, 301, 333), (2, 64, 64), (2, 120, 97)])
def test_device_conditioning_is_bit_exact_with_opencv(shape):
    """aau_condition_frames (min-max, CLAHE 1.0 / 8x8, median 3) against the oracle's cv2 calls -- the same calls the
    reference makes (pinned by the reference-run golden above) -- on speckle, random, constant and narrow-range frames."""
    from attention_aspp_unet import AttentionASPPUNet
    from fetal_abdomen import FetalAbdomenSegmentation
    n, h, w = shape
    rng = np.random.default_rng(n * h + w)
    sweep = O.synthetic_sweep(n, h, w, seed=h, peak=n // 2)
    sweep[1] = rng.integers(0, 256, (h, w), dtype=np.uint8)
    if n > 2:
        sweep[2] = (rng.random((h, w)) ** 3 * 180 + 17).astype(np.uint8)          # narrow range: the min-max stretch matters
    if n > 3:
        sweep[3] = 77                                                              # constant frame: scale 0
    want = np.rint(O.condition_frames(sweep) * 255).astype(np.uint8)
    algo = FetalAbdomenSegmentation(net=AttentionASPPUNet(base_c=16), batch=4)
    got = algo.condition_on_device(torch.from_numpy(sweep).cuda()).cpu().numpy()
    assert got.shape == want.shape
    bad = int((got != want).sum())
    assert bad == 0, f"{bad} of {got.size} conditioned pixels differ from OpenCV (max |d| {np.abs(got.astype(int) - want.astype(int)).max()})"
