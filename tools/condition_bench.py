"""Frame conditioning throughput: aau_condition_frames on an 840-frame 562x744 uint8 sweep
resident in HBM (CUDA events) against the reference's own OpenCV calls on the host (bounded sample).
    python tools/condition_bench.py [out.json]"""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "att-aspp-unet_b200", ROOT / "oracle", ROOT):
    sys.path.insert(0, str(p))
import numpy as np, torch
import bench, aau_oracle as O
from attention_aspp_unet import AttentionASPPUNet
from fetal_abdomen import FetalAbdomenSegmentation

out_path = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "gpurun_out" / "condition_bench.json")
vol = bench.make_sweep()
dev = torch.from_numpy(vol).cuda()
algo = FetalAbdomenSegmentation(net=AttentionASPPUNet(base_c=16), batch=56)
B = 56
outb = torch.empty((B,) + dev.shape[1:], dtype=torch.uint8, device="cuda")
for _ in range(2):
    for s in range(0, dev.shape[0], B):
        algo.condition_on_device(dev[s:s + B], out=outb)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    for s in range(0, dev.shape[0], B):
        algo.condition_on_device(dev[s:s + B], out=outb)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
n = vol.shape[0]
gpu_fps = 1e3 * n / ms
alg_bytes = 2.0 * vol.size                                         # one read + one write of the sweep (algorithmic)
sample = 24
t0 = time.perf_counter()
ref = O.condition_frames(vol[:sample])
cpu_fps = sample / (time.perf_counter() - t0)
got = algo.condition_on_device(dev[:sample].contiguous()).cpu().numpy()
exact = bool(np.array_equal(got, np.rint(ref * 255).astype(np.uint8)))
peak = bench.peaks()["hbm"]
res = {"workload": "840 x 562x744 uint8 frames, batches of 56, min-max + CLAHE(1.0, 8x8) + median 3", "ms_per_sweep": ms, "frames_per_s": gpu_fps,
       "algorithmic_GBs": alg_bytes / ms / 1e6, "hbm_peak_GBs": peak, "frac_of_hbm_peak": alg_bytes / ms / 1e6 / peak,
       "kernels_per_batch": 4, "bit_exact_with_opencv_on_sample": exact,
       "cpu_baseline": {"frames_per_s": cpu_fps, "kind": "reference's OpenCV calls, one host thread", "sample": f"{sample} frames"}}
print(json.dumps(res, indent=1))
Path(out_path).write_text(json.dumps(res, indent=1))
