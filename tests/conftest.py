"""Shared test plumbing: import paths, the `gpu` marker and a few helpers.

`-m "not gpu"` covers the oracle against the committed golden vectors, the host logic and the C-ABI surface;
`-m gpu` are the parity tests proper (CUDA path through libaau.so vs the oracle).
Nothing here reads /root/reference at run time (it does not exist on the GPU box).
"""
import json
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "att-aspp-unet_b200", ROOT / "oracle", ROOT):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def manifest():
    return json.loads((GOLDEN / "manifest.json").read_text())


def golden_case(name, manifest):
    """(cfg, state_dict, x, golden npz) for a committed case; weights/inputs are regenerated from seeds."""
    import aau_oracle as O
    m = manifest[name]
    c = m["cfg"]
    cfg = O.NetCfg(base_c=c["base_c"], variant=c["variant"], use_att=c["use_att"], use_aspp=c["use_aspp"], att_depth=c["att_depth"])
    b, h, w = m["shape"]
    seed = m["seed"]

    def inp(s, bb):
        if m["input"] == "rand":
            return torch.rand(bb, 1, h, w, generator=torch.Generator().manual_seed(s))
        vol = O.synthetic_sweep(bb, h, w, seed=s, peak=bb // 2)
        return torch.from_numpy(vol.astype(np.float32) / 255.0).unsqueeze(1)

    sd = O.make_state_dict(cfg, seed=seed, regime=m["regime"])
    if m["regime"] == "R1":
        sd = O.calibrate_bn(sd, inp(seed + 2, max(b, 2)), cfg)
    return cfg, sd, inp(seed + 1, b), np.load(GOLDEN / f"{name}.npz")
