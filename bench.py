#!/usr/bin/env python
"""Headline benchmark: frames/sec of the AttentionASPPUNet forward + per-frame score + best-frame selection on a
synthetic ACOUSLIC-style sweep (840 frames of 744x562, base_c=32) -- BASELINE.json `metric`, config[1] at N=1 and
config[2] (one sweep per GPU, sharded by case, host gather of scores) at N>1.

    python bench.py --gpus N --steps K --warmup W            # the CUDA engine (libaau.so)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores (CPU oracle)

A step = one pass of the hot path over one sweep.  `value` = frames of all ranks / max-over-ranks device time with
the uint8 sweep already resident in HBM; `e2e` = the same through the public host API
(FetalAbdomenSegmentation.segment_sweep) from pinned host memory, H2D copies and the D2H of areas / index / mask
inside the timed region.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
for p in (ROOT / "att-aspp-unet_b200", ROOT / "oracle"):
    sys.path.insert(0, str(p))

H, W, N_FRAMES, BASE_C = 562, 744, 840, 32
GFLOP_PER_FRAME = 160.319          # SURVEY.md section 8d / BASELINE.md section 3 (2*MAC, dense tap count)
METRIC = "frames/sec Att-ASPP-UNet fwd (744x562 US)"


_REAL_STDOUT = None


def quiet_stdout():
    """Route fd 1 to stderr for the whole run (NCCL and friends print banners there) and keep the real stdout for the
    ONE JSON line the driver parses."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm": d["hbm_gbs"], "tc_burst": d["bf16_tflops"], "tc_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm": 6650.0, "tc_burst": 1590.0, "tc_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [t.strip() for t in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


WORKLOAD = ("config[1]: one 840-frame 744x562 uint8 sweep per GPU, AttentionASPPUNet base_c=32, BN-calibrated random weights, "
            "forward + sigmoid/threshold/area + first-max argmax")


def make_weights():
    import aau_oracle as O
    cfg = O.NetCfg(base_c=BASE_C)
    g = torch.Generator().manual_seed(2025)
    calib = torch.from_numpy(O.synthetic_sweep(2, H // 2, W // 2, seed=7, peak=1).astype(np.float32) / 255.0).unsqueeze(1)
    sd = O.calibrate_bn(O.make_state_dict(cfg, 2025, "R1"), calib, cfg)
    return cfg, sd


def make_sweep(n_unique=24, seed=2025):
    """uint8 [840,562,744]: `n_unique` distinct synthetic frames tiled along the sweep with a per-frame intensity
    ramp (generating 840 Rayleigh-speckle frames on the host would dominate the run; the network does the same work)."""
    import aau_oracle as O
    base = O.synthetic_sweep(n_unique, H, W, seed=seed, peak=n_unique // 2)
    reps = (N_FRAMES + n_unique - 1) // n_unique
    return np.ascontiguousarray(np.concatenate([base] * reps)[:N_FRAMES])


# ----------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU implementation of the path (its torch fp32 forward + numpy selection, restated in
    oracle/aau_oracle.py because the reference itself cannot travel to the GPU box), on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import aau_oracle as O
    cfg, sd = make_weights()
    frames = args.ref_frames
    vol = O.synthetic_sweep(frames, H, W, seed=2025, peak=frames // 2)
    x = torch.from_numpy(vol.astype(np.float32) / 255.0).unsqueeze(1)
    threads = torch.get_num_threads()

    def step():
        logits = O.forward(sd, x, cfg)
        prob = torch.sigmoid(logits)[:, 0].numpy()
        m3 = O.postprocess(prob, thr=args.prob_thr)
        return O.select_fetal_abdomen_mask_and_frame(m3)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    fps = frames * args.steps / dt
    sample = f"{frames} synthetic 562x744 frames per step (one batch-{frames} fp32 forward + sigmoid + postprocess + select), of the 840-frame sweep"
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": frames, "note": "bounded CPU sample of that workload"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample,
                             "host_cpus": os.cpu_count(), "torch": torch.__version__},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------------------------------------------
def cpu_baseline_sample(cfg, sd, frames=8, reps=2):
    import aau_oracle as O
    vol = O.synthetic_sweep(frames, H, W, seed=2025, peak=frames // 2)
    x = torch.from_numpy(vol.astype(np.float32) / 255.0).unsqueeze(1)
    O.forward(sd, x[:1], cfg)                        # warm-up (thread pool, oneDNN primitives)
    t0 = time.perf_counter()
    for _ in range(reps):
        logits = O.forward(sd, x, cfg)
        O.select_fetal_abdomen_mask_and_frame(O.postprocess(torch.sigmoid(logits)[:, 0].numpy()))
    dt = time.perf_counter() - t0
    return {"value": frames * reps / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{reps} x batch-{frames} of the sweep's 562x744 frames, fp32 torch CPU forward + numpy selection (oracle)",
            "host_cpus": os.cpu_count()}


def run_engine(args):
    import torch.distributed as dist
    from attention_aspp_unet import AttentionASPPUNet
    from fetal_abdomen import FetalAbdomenSegmentation
    from sharding import gather_areas

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg, sd = make_weights()
    net = AttentionASPPUNet(base_c=BASE_C, act_dtype=args.dtype)
    net.load_state_dict(sd, strict=True)
    net.eval().prepare(dev)
    for kv in args.opt:
        k, v = kv.split("=")
        net.set_option(k, int(v))
    seg = FetalAbdomenSegmentation(net=net, batch=args.batch, device=dev)
    vol_np = make_sweep(seed=2025 + rank)                       # N>1: one case per rank (config[2], sharded by case)
    vol_pinned = torch.from_numpy(vol_np).pin_memory()
    vol_dev = torch.from_numpy(vol_np).to(dev)
    B = args.batch
    starts = list(range(0, N_FRAMES, B))
    logits_all = torch.empty((N_FRAMES, 1, H, W), dtype=torch.float32, device=dev)
    areas = torch.zeros(N_FRAMES, dtype=torch.int32, device=dev)
    best = torch.zeros(2, dtype=torch.int32, device=dev)
    thr = args.prob_thr

    def device_step():
        for s in starts:
            b = min(B, N_FRAMES - s)
            lg = net(vol_dev[s: s + b], out=logits_all[s: s + b])
            seg._scores.run(lg[:, 0], 0, thr, areas[s: s + b], None, None)
        seg._scores.best(areas, best)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- kernel-only arm (inputs resident in HBM)
    for _ in range(args.warmup):
        device_step()
    net.check_device()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        device_step()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    launches_per_step = len(starts) * (net.num_launches() + 1) + 1
    best_dev = [int(v) for v in best.cpu().numpy()]

    # ---------------- end-to-end arm (public API, pinned host input, results back on the host)
    res = None
    for _ in range(max(1, min(args.warmup, 2))):
        res = seg.segment_sweep(vol_pinned, prob_thr=thr)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = seg.segment_sweep(vol_pinned, prob_thr=thr)
        if world > 1:
            gather_areas(np.array([res["best_area"], res["best_idx"]], np.int32), 2 * world, dev)   # host gather of per-case scores
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))   # host work after the last kernel counts too

    # ---------------- per-launch timing for the roofline (profile mode, one extra untimed sweep)
    net.set_option("profile", 1)
    tc_ms = tc_fl = 0.0
    layer_rows = {}
    nb = 0
    for s in starts[: max(1, min(len(starts), 6))]:
        b = min(B, N_FRAMES - s)
        net(vol_dev[s: s + b], out=logits_all[s: s + b])
        for r in net.op_profile():
            acc = layer_rows.setdefault(r["layer"], {"kernel": r["kernel"], "ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
            acc["ms"] += r["ms"]; acc["flops"] += r["flops"]; acc["bytes"] += r["bytes"]; acc["n"] += 1
            if r["kernel"] == "igemm_tc_kernel":
                tc_ms += r["ms"]; tc_fl += r["flops"]
        nb += 1
    net.set_option("profile", 0)
    n_tc = sum(1 for v in layer_rows.values() if v["kernel"] == "igemm_tc_kernel")

    # max over ranks
    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    frames_total = N_FRAMES * args.steps * world
    fps = frames_total / (ms / 1e3)
    fps_e2e = frames_total / (ms_e2e / 1e3)
    tc_tflops = tc_fl / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else 0.0
    fwd_ms = sum(v["ms"] for v in layer_rows.values()) / max(nb, 1)
    traffic, traffic_src = None, None
    tp = ROOT / "profiles" / "r01_igemm_dram_traffic.json"          # ncu --set full capture of the igemm launches of one forward
    if tp.exists():
        try:
            tj = json.loads(tp.read_text())
            traffic = tj["dram_bytes_per_frame"] * B                 # per forward of this bench's batch, like `achieved`
            traffic_src = f"ncu dram__bytes_read+write summed over the {tj['launches']} igemm launches of one forward: {tj['dram_bytes_per_frame'] / 1e6:.0f} MB per frame (profiles/r01_igemm_dram_traffic.json) x batch"
        except Exception:
            pass
    roof = {"bound": "tensor", "achieved": tc_tflops, "peak": pk["tc_sustained"], "unit": "TFLOP/s", "frac": tc_tflops / pk["tc_sustained"],
            "traffic": traffic, "traffic_note": traffic_src, "kernel": "igemm_tc_kernel", "peak_source": pk["source"] + " (sustained bf16: kernel timed inside a long step)",
            "note": f"algorithmic FLOPs of the {n_tc} igemm_tc_kernel launches of one forward / their summed CUDA-event time; "
                    f"they are {tc_ms / max(sum(v['ms'] for v in layer_rows.values()), 1e-9):.0%} of the forward",
            "whole_forward_tflops": fps / world * GFLOP_PER_FRAME / 1e3, "whole_forward_frac": fps / world * GFLOP_PER_FRAME / 1e3 / pk["tc_sustained"]}
    out_dir = ROOT / "gpurun_out"
    try:
        out_dir.mkdir(exist_ok=True)
        rows = [{"layer": k, **v, "ms_per_fwd": v["ms"] / nb, "tflops": v["flops"] / max(v["ms"], 1e-9) / 1e9,
                 "gbs": v["bytes"] / max(v["ms"], 1e-9) / 1e6} for k, v in layer_rows.items()]
        (out_dir / f"layers_b{B}_{args.dtype}.json").write_text(json.dumps({"batch": B, "fwd_ms": fwd_ms, "rows": rows}, indent=1))
    except Exception:
        pass
    line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic",
            "config": {"workload": WORKLOAD + (" (config[2]: one case per GPU, host gather)" if world > 1 else ""),
                       "frames_per_step_per_gpu": N_FRAMES, "batch": B, "act_dtype": args.dtype, "l2": "inputs larger than L2 (351 MB sweep; >6 GB of activations per batch)",
                       "parallelism": f"dp{world} by case, no collective on the forward path"},
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": int(res["h2d_bytes"]), "d2h_bytes_per_step": int(res["d2h_bytes"]),
                    "api": "FetalAbdomenSegmentation.segment_sweep(pinned uint8 sweep) -> areas, best index, post-processed mask"},
            "gpu_launches": launches_per_step * args.steps, "clocks": clocks, "roofline": roof,
            "selected_frame": {"device_arm": best_dev[0], "e2e_arm": res["best_idx"], "area": res["best_area"], "prob_thr": thr,
                               "distinct_areas": int(len(np.unique(areas.cpu().numpy())))}}
    if world == 1:
        line["cpu_baseline"] = cpu_baseline_sample(cfg, sd, frames=args.cpu_frames, reps=5 if args.cpu_frames >= 8 else 2)   # ~10 s of CPU work by default
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--batch", type=int, default=56, help="frames per forward (a divisor of 840)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (amode, resident, ctas)")
    ap.add_argument("--ref-frames", type=int, default=4, help="frames per step of the CPU reference arm (bounded sample)")
    ap.add_argument("--cpu-frames", type=int, default=8, help="frames in the cpu_baseline sample of the engine arm")
    ap.add_argument("--prob-thr", type=float, default=0.5,
                    help="probability threshold of the per-frame area score (0.5 = the pipeline CLI's binarisation; the wrapper's "
                         "0.05 marks every pixel of a random-weight network, so every frame would tie at the full-frame area)")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    if not torch.cuda.is_available():
        emit({"error": "no CUDA device: the engine has no CPU fallback"})
        sys.exit(2)
    run_engine(args)


if __name__ == "__main__":
    main()
