"""CPU-side tests (-m "not gpu"): module surface / state_dict contract, C-ABI exports, host selection helpers and the
world_size-2 gather (gloo).  No compute call into libaau happens here (there is no GPU in this container)."""
import ctypes
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import aau_oracle as O
from conftest import GOLDEN, ROOT


def test_cabi_exports_every_declared_symbol():
    import _capi
    header = (ROOT / "include" / "aau.h").read_text()
    declared = set(re.findall(r"\b(aau_[a-z_0-9]+)\s*\(", header))
    declared -= {"aau_status", "aau_config", "aau_handle"}
    assert len(declared) >= 15
    if not _capi.LIB_PATH.exists():
        import __graft_entry__ as g
        g.build()
    L = ctypes.CDLL(str(_capi.LIB_PATH))
    for name in sorted(declared):
        assert hasattr(L, name), f"libaau.so does not export {name}"
    bound = {n for n, _, _ in _capi.SYMBOLS}
    assert declared == bound, f"ctypes table and header differ: {declared ^ bound}"


def test_every_planner_option_is_documented_in_the_header():
    """include/aau.h lists the names aau_set_option accepts; the table in the engine is the source of truth."""
    eng = (ROOT / "att-aspp-unet_b200" / "csrc" / "aau_engine.cu").read_text()
    table = eng[eng.index("plan_options[] = {"):]
    table = table[:table.index("};")]
    names = set(re.findall(r'\{"(\w+)", &e\.opt_', table))
    assert len(names) >= 20, names
    doc = (ROOT / "include" / "aau.h").read_text()
    doc = doc[doc.index("Planner options"):doc.index("int aau_set_option")]
    missing = sorted(n for n in names if f'"{n}"' not in doc)
    assert not missing, f"options accepted by aau_set_option but not described in include/aau.h: {missing}"


def test_product_sources_never_reach_for_the_oracle_or_a_cpu_path():
    """The oracle is test infrastructure: nothing under the package (or the C ABI sources) may import / execute it."""
    pkg = ROOT / "att-aspp-unet_b200"
    for f in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*")):
        text = f.read_text()
        assert "aau_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f.name


def test_tile_maps_enumerate_every_tile_once():
    """decode_tile / TileIter of the kernel header are __host__ __device__: tests/decode_check.cu walks them on the CPU for
    the planner's launch geometries (stacked M-blocks, several N tiles, several problems, two-frame tiles, CTA-pair order)."""
    import shutil, subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = ROOT / "tests" / "_build"
    out.mkdir(exist_ok=True)
    exe = out / "decode_check"
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-o", str(exe), str(ROOT / "tests" / "decode_check.cu")],
                   check=True, cwd=str(ROOT))
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "ok (0 failures)" in r.stdout, r.stdout[-2000:]


def test_create_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import _capi
    L = _capi.lib()
    cfg = _capi.AauConfig(1, 1, 32, 0, 1, 1, 4, 0)
    h = ctypes.c_void_p()
    st = L.aau_create(ctypes.byref(cfg), 0, ctypes.byref(h))
    assert st != 0 and not h.value
    assert b"no CUDA device" in L.aau_last_error(None) or b"CPU fallback" in L.aau_last_error(None)


def test_module_state_dict_contract(manifest):
    from attention_aspp_unet import AttentionASPPUNet
    m = AttentionASPPUNet(in_ch=1, num_classes=1, base=16)                      # model_attention_aspp.py:36 spelling
    lay = manifest["pipe_c16_R0_64x80"]["state_dict_layout"]
    assert [(k, list(v.shape)) for k, v in m.state_dict().items()] == [(k, list(s)) for k, s in lay]
    m32 = AttentionASPPUNet()                                                   # canonical defaults :112
    assert len(m32.state_dict()) == 196 and sum(p.numel() for p in m32.parameters()) == 9_426_567 - 0
    a = AttentionASPPUNet(base_c=16, use_att=True, use_aspp=True, att_depth=4)
    lay = manifest["abl_full_c16_R1_80x72"]["state_dict_layout"]
    assert [(k, list(v.shape)) for k, v in a.state_dict().items()] == [(k, list(s)) for k, s in lay]
    counts = {(): 150, (("use_att", False),): 142, (("use_aspp", False),): 120, (("use_att", False), ("use_aspp", False)): 112,
              (("att_depth", 3),): 146}
    for kw, n in counts.items():
        kw = dict(kw)
        if not kw:
            kw = {"use_att": True}
        assert len(AttentionASPPUNet(base_c=32, **kw).state_dict()) == n


def test_module_matches_oracle_spec_for_every_config():
    from attention_aspp_unet import AttentionASPPUNet
    cases = [(O.NetCfg(base_c=48), dict(base_c=48)),
             (O.NetCfg(base_c=16, variant="ablation", att_depth=3), dict(base_c=16, att_depth=3)),
             (O.NetCfg(base_c=16, variant="ablation", use_aspp=False, use_att=False), dict(base_c=16, use_aspp=False, use_att=False))]
    for cfg, kw in cases:
        sd = AttentionASPPUNet(**kw).state_dict()
        assert [(k, tuple(v.shape)) for k, v in sd.items()] == [(k, tuple(s)) for k, s, _ in O.state_dict_spec(cfg)]


def test_load_state_dict_semantics():
    from attention_aspp_unet import AttentionASPPUNet
    cfg = O.NetCfg(base_c=16)
    sd = O.make_state_dict(cfg, 1, "R1")
    m = AttentionASPPUNet(base_c=16)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(m.state_dict()["u3.att.Wg.0.weight"], sd["u3.att.Wg.0.weight"])
    # legacy spelling + extra + missing keys with strict=False (attention_aspp_unet_pipeline_stage.py:134-141)
    legacy = {k.replace(".Wg.", ".W_g.").replace(".Wx.", ".W_x."): v for k, v in sd.items()}
    legacy["bogus.weight"] = torch.zeros(1)
    del legacy["out_conv.bias"]
    miss, unexp = AttentionASPPUNet(base_c=16).load_state_dict(legacy, strict=False)
    assert miss == ["out_conv.bias"] and unexp == ["bogus.weight"]
    with pytest.raises(RuntimeError):
        AttentionASPPUNet(base_c=16).load_state_dict(legacy, strict=True)
    wrapped = AttentionASPPUNet(base_c=16, use_att=True)
    sda = O.make_state_dict(O.NetCfg(base_c=16, variant="ablation"), 1, "R0")
    assert not wrapped.load_state_dict({"state_dict": sda}, strict=True).missing_keys      # test_ablation.py:224-225


def test_module_refuses_cpu_and_training_forward():
    from attention_aspp_unet import AttentionASPPUNet
    m = AttentionASPPUNet(base_c=16)
    with pytest.raises(RuntimeError, match="eval"):
        m(torch.zeros(1, 1, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.eval()(torch.zeros(1, 1, 32, 32))
    for bad in (dict(base_c=20), dict(in_channels=3), dict(num_classes=2), dict(act_dtype="fp8"), dict(variant="pipeline", use_att=False)):
        with pytest.raises(ValueError):
            AttentionASPPUNet(**bad)


def test_largest_component_matches_oracle_postprocess():
    from fetal_abdomen import largest_component, merge_shard_scores
    g = np.load(GOLDEN / "selection.npz")
    for vol in ("random", "blobs", "ties"):
        prob = g[vol + "_prob"]
        bin_ = (prob > 0.05).astype(np.uint8)
        idx = int(bin_.sum((1, 2)).argmax())
        assert np.array_equal(largest_component(bin_[idx]), g[vol + "_mask3d"][idx])
    areas, idx = merge_shard_scores([np.array([0, 3, 5]), np.array([5, 1]), np.array([], np.int32)])
    assert idx == 2 and areas.tolist() == [0, 3, 5, 5, 1]
    assert merge_shard_scores([np.zeros(4, np.int32)])[1] == -1 and merge_shard_scores([])[1] == -1


def test_shard_ranges_cover_and_balance():
    from sharding import owner_of, select_global, shard_range
    for n, w in ((840, 8), (840, 4), (7, 4), (3, 8), (0, 2), (105, 1)):
        blocks = [shard_range(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1
    assert shard_range(840, 8, 3) == (315, 420) and owner_of(419, 840, 8) == 3 and owner_of(420, 840, 8) == 4
    assert select_global(np.array([1, 7, 7, 2])) == (1, 7) and select_global(np.zeros(3, np.int32)) == (-1, 0)


def _gather_worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(ROOT / "att-aspp-unet_b200"))
    from sharding import gather_areas, select_global, shard_range
    full = (np.random.default_rng(3).integers(0, 50, n)).astype(np.int32)
    full[[5, 11]] = 99                                            # a tie across ranks: the lower frame must win
    lo, hi = shard_range(n, world, rank)
    got = gather_areas(full[lo:hi], n)
    q.put((rank, got.tolist(), select_global(got)))
    dist.destroy_process_group()


def test_two_rank_gather_gloo():
    n, world = 17, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, n, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(60) for p in procs]
    full = (np.random.default_rng(3).integers(0, 50, n)).astype(np.int32)
    full[[5, 11]] = 99
    for _, got, sel in res:
        assert got == full.tolist() and tuple(sel) == (5, 99)


# ------------------------------------------------------------------------------------------------ host-only helpers of libaau
def _lib():
    import _capi
    if not _capi.LIB_PATH.exists():
        import __graft_entry__ as g
        g.build()
    return _capi.lib()


def test_logit_cutoff_is_the_threshold_in_logit_space():
    """sigmoid(x) > t  <=>  x > cutoff(t) (SURVEY.md identity i7): the library's cutoff (correctly rounded fp32 sigmoid, found
    by bisection in C++) and the Python layer's (bisection on THIS host's torch.sigmoid) separate the fp32 line exactly."""
    from fetal_abdomen import logit_cutoff
    L = _lib()
    for t in (0.05, 0.48, 0.5, 0.9, 1e-3, 0.999):
        c = float(L.aau_logit_cutoff(ctypes.c_float(t)))
        up = float(np.nextafter(np.float32(c), np.float32(np.inf)))
        sig = lambda x: np.float32(1.0 / (1.0 + np.exp(-np.float64(x))))   # noqa: E731  correctly rounded fp32 sigmoid
        assert not (sig(c) > np.float32(t)) and sig(up) > np.float32(t), t
        assert abs(c - np.log(t / (1 - t))) < 1e-5 * max(1.0, abs(c))
        ct = logit_cutoff(t)                                              # torch's own sigmoid: the same cutoff up to a few ulps
        # the two sigmoids may differ by an ulp of their OUTPUT: 2^-24 / sigmoid'(c) in logit space
        assert abs(ct - c) <= 4 * 2.0 ** -24 / (t * (1 - t)) + 8 * abs(float(np.spacing(np.float32(c)))), (t, c, ct)
        x = torch.tensor([ct, float(np.nextafter(np.float32(ct), np.float32(np.inf)))] * 32, dtype=torch.float32)
        assert (torch.sigmoid(x) > np.float32(t)).tolist()[:2] == [False, True]
    assert float(L.aau_logit_cutoff(ctypes.c_float(1.0))) == float("inf") and float(L.aau_logit_cutoff(ctypes.c_float(-0.1))) == float("-inf")
    assert logit_cutoff(1.0) == float("inf") and logit_cutoff(-0.1) == float("-inf")


@pytest.mark.parametrize("fp16", [1, 0])
def test_window_sum_preserving_rounding(fp16):
    """aau_round_window_keep_sum (what aau_commit_weights applies to every 3x3 window): each tap within one ulp of its value,
    the rounded taps sum to within half an ulp (of the largest tap) of the true sum, never worse than plain nearest rounding."""
    L = _lib()
    rng = np.random.default_rng(5)
    to_f = (lambda b: b.view(np.float16).astype(np.float64)) if fp16 else (lambda b: (b.astype(np.uint32) << 16).view(np.float32).astype(np.float64))
    ulp = (lambda v: np.spacing(np.abs(v).astype(np.float16)).astype(np.float64)) if fp16 else (lambda v: np.abs(v) * 2.0 ** -7)
    worse = better = 0
    for _ in range(400):
        w = rng.uniform(-1, 1, 9) * rng.choice([1e-3, 3e-2, 0.2, 1.0])
        out = np.zeros(9, np.uint16)
        assert L.aau_round_window_keep_sum(w.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), fp16, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16))) == 0
        r = to_f(out)
        assert np.all(np.abs(r - w) <= 1.001 * ulp(w) + 1e-12)            # every weight stays within one ulp
        nearest = w.astype(np.float16).astype(np.float64) if fp16 else to_f(((w.astype(np.float32).view(np.uint32) + 0x7FFF + ((w.astype(np.float32).view(np.uint32) >> 16) & 1)) >> 16).astype(np.uint16))
        e_keep, e_near = abs(r.sum() - w.sum()), abs(nearest.sum() - w.sum())
        assert e_keep <= 0.5001 * ulp(w).max() + 1e-12                    # the window sum is kept to half an ulp of the largest tap
        assert e_keep <= e_near + 1e-12
        worse += e_keep > e_near
        better += e_keep < e_near
    assert better > 100 and worse == 0                                  # (equal whenever nearest rounding already keeps the sum)
