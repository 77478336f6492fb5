"""Pins the CPU oracle (oracle/aau_oracle.py) to outputs of the REAL reference modules.

The vectors under tests/golden/ were produced by oracle/gen_golden.py, which imports /root/reference, loads the
oracle-generated state dicts with strict=True and runs the reference forward / selection functions.  These tests
re-create the same weights and inputs from seeds and require the restatement to reproduce the reference.
"""
import numpy as np
import pytest
import torch

import aau_oracle as O
from conftest import GOLDEN, golden_case

SMALL = ["pipe_c16_R0_64x80", "pipe_c16_R1_141x93", "pipe_c32_R1_64x64", "abl_full_c16_R1_80x72", "abl_noatt_c16_R1_80x72",
         "abl_noaspp_c16_R1_80x72", "abl_neither_c16_R1_80x72", "abl_depth3_c16_R1_81x73"]


@pytest.mark.parametrize("name", SMALL)
def test_forward_matches_reference(name, manifest):
    cfg, sd, x, gold = golden_case(name, manifest)
    assert len(sd) == manifest[name]["n_state_dict_entries"]
    out = O.forward(sd, x, cfg)
    if cfg.variant == "pipeline":
        logits = out
    else:
        logits, (psi3, psi2) = out
        np.testing.assert_allclose(psi3.numpy(), gold["psi3"], atol=2e-5, rtol=0)
        np.testing.assert_allclose(psi2.numpy(), gold["psi2"], atol=2e-5, rtol=0)
    # fp32 CPU reruns of the same ATen kernels: only reduction-order noise is allowed
    np.testing.assert_allclose(logits.numpy(), gold["logits"], atol=5e-5, rtol=0)


def test_forward_matches_reference_full_frame(manifest):
    """One 562x744 frame, base_c=32, BN-calibrated weights: the bench geometry (stored subsampled by 7)."""
    name = "pipe_c32_R1_562x744"
    cfg, sd, x, gold = golden_case(name, manifest)
    logits = O.forward(sd, x, cfg).numpy()
    s = manifest[name]["stride"]
    np.testing.assert_allclose(logits[:, :, ::s, ::s], gold["logits"], atol=1e-4, rtol=0)
    stats = np.array([logits.mean(), logits.std(), logits.min(), logits.max()])
    np.testing.assert_allclose(stats, gold["logits_stats"], atol=1e-4, rtol=0)
    assert logits.std() > 0.1, "R1 weights must give non-degenerate logits (SURVEY.md section 7, hard part 1)"


def test_state_dict_layout_matches_reference(manifest):
    for name, cfg in (("pipe_c16_R0_64x80", O.NetCfg(base_c=16)), ("abl_full_c16_R1_80x72", O.NetCfg(base_c=16, variant="ablation"))):
        lay = manifest[name]["state_dict_layout"]
        spec = O.state_dict_spec(cfg)
        assert [(k, list(s)) for k, s, _ in spec] == [(k, list(s)) for k, s in lay]
    assert len(O.state_dict_spec(O.NetCfg(base_c=32))) == 196


@pytest.mark.parametrize("vol", ["random", "blobs", "empty", "ties"])
def test_selection_matches_reference(vol):
    g = np.load(GOLDEN / "selection.npz")
    prob = g[vol + "_prob"]
    m3 = O.postprocess(prob)
    assert m3.dtype == np.uint8 and np.array_equal(m3, g[vol + "_mask3d"])
    m2, idx = O.select_fetal_abdomen_mask_and_frame(m3)
    assert idx == int(g[vol + "_idx"]) and np.array_equal(m2, g[vol + "_mask2d"])


def test_selection_2d_and_volume_helpers():
    m2, idx = O.select_fetal_abdomen_mask_and_frame(np.array([[0, 3], [0, 0]], np.uint8))
    assert idx == 0 and m2.tolist() == [[0, 1], [0, 0]]
    vol = O.convert_2d_mask_to_3d(np.array([[1, 0]], np.uint8), 1, 3)
    assert vol.shape == (3, 1, 2) and vol[1, 0, 0] == 2 and vol.sum() == 2
    assert O.convert_2d_mask_to_3d(np.array([[1, 0]], np.uint8), -1, 3).sum() == 0
    with pytest.raises(ValueError):
        O.convert_2d_mask_to_3d(np.array([[1, 0]], np.uint8), 3, 3)


def test_tta_is_flip_symmetric(manifest):
    cfg, sd, x, _ = golden_case("pipe_c16_R0_64x80", manifest)
    p = O.predict_prob_tta(sd, x[:1], cfg)
    q = O.predict_prob_tta(sd, torch.flip(x[:1], [-1]), cfg)
    np.testing.assert_allclose(p.numpy(), torch.flip(q, [-1]).numpy(), atol=1e-6)


def test_synthetic_sweep_properties():
    v = O.synthetic_sweep(6, 96, 128, seed=3, peak=3)
    assert v.dtype == np.uint8 and v.shape == (6, 96, 128)
    nz = (v[0] > 0).mean()
    assert 0.4 < nz < 0.8                                              # fan-shaped field of view
    assert v[:, 0, 0].max() == 0                                       # exact zeros outside the fan
    assert len({v[i].tobytes() for i in range(6)}) == 6                # every frame distinct
