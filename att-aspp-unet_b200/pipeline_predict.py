"""The thesis' full-frame inference recipe on the B200 engine.

Mirror of ``predict`` and its helpers in the reference's pipeline script (attention_aspp_unet_pipeline_stage.py:336-374,
399-523; the working ``select_best`` is the ablation twin's, test_ablation.py:389-403):

    per slice:  min-max -> CLAHE(1.0, 8x8) -> median 3 -> resize to 512x512 -> /255
                prob = sigmoid((net(x) + flip(net(flip(x)))) / 2)            (flip test-time augmentation)
                resize back, GaussianBlur 5x5, > THR (0.48), refine_mask
    per case :  best frame = most circular of the 5 largest masks, output.mha + frame-number json, AC in mm (ellipse fit)

What runs where: conditioning (uint8 sweeps), both network passes, the flip and the sigmoid-of-mean run on the GPU
through libaau (``aau_condition_frames``, ``aau_forward``, ``aau_flip_w``, ``aau_tta_prob``), batched over the sweep;
the per-slice OpenCV / SciPy steps the reference performs on the host (resize, blur, connected components, morphology,
contours, ellipse fit) stay host calls to the same libraries, spread over a thread pool.  ``skimage.measure.label`` (absent
here) is replaced by ``scipy.ndimage.label`` with the same full (8-) connectivity.
"""
from __future__ import annotations

import csv
import json
import math
import os
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

try:
    import _capi
    from attention_aspp_unet import AttentionASPPUNet
    from fetal_abdomen import FetalAbdomenSegmentation, preprocess_sweep
    from metaimage import read_mha, write_mha
except ImportError:                                     # package-style import
    from . import _capi                                 # type: ignore
    from .attention_aspp_unet import AttentionASPPUNet  # type: ignore
    from .fetal_abdomen import FetalAbdomenSegmentation, preprocess_sweep  # type: ignore
    from .metaimage import read_mha, write_mha          # type: ignore

IMG_SIZE = 512                                          # attention_aspp_unet_pipeline_stage.py:29
DEFAULT_THR = 0.48                                      # :405


# ------------------------------------------------------------------------------------------------ host helpers
def _label8(m: np.ndarray) -> np.ndarray:
    import scipy.ndimage as ndi
    return ndi.label(m, structure=np.ones((3, 3), np.uint8))[0]


def refine_mask(m: np.ndarray) -> np.ndarray:
    """Drop components under max(20, 0.15 % of the frame), keep the largest, close with a 7x7 ellipse, fill holes."""
    import cv2
    from scipy.ndimage import binary_fill_holes
    if m.sum() == 0:
        return m
    lab = _label8(m)
    cnt = np.bincount(lab.ravel())
    cnt[0] = 0
    keep = [i for i, c in enumerate(cnt) if c >= max(20, int(0.0015 * m.size))]
    if not keep:
        return np.zeros_like(m)
    m = np.isin(lab, keep).astype(np.uint8)
    lab2 = _label8(m)
    bc = np.bincount(lab2.ravel())
    bc[0] = 0
    m = (lab2 == np.argmax(bc)).astype(np.uint8)
    k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (7, 7))
    return binary_fill_holes(cv2.morphologyEx(m, cv2.MORPH_CLOSE, k)).astype(np.uint8)


def _circularity_score(mask: np.ndarray) -> float:
    import cv2
    cnts, _ = cv2.findContours(mask.astype(np.uint8), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    if not cnts:
        return 0.0
    c = max(cnts, key=cv2.contourArea)
    area, peri = cv2.contourArea(c), cv2.arcLength(c, True)
    return 0.0 if peri <= 1e-6 else 4 * np.pi * area / (peri ** 2)


def select_best(stack: Sequence[np.ndarray], topk: int = 5) -> int:
    if len(stack) == 0:
        return 0
    areas = np.array([(m > 0).sum() for m in stack])
    idx = areas.argsort()[::-1][: max(1, min(topk, len(areas)))]
    return int(max(idx, key=lambda i: _circularity_score(stack[i])))


def _ellipse_circum(a: float, b: float) -> float:
    h = ((a - b) ** 2) / ((a + b) ** 2)
    return math.pi * (a + b) * (1 + 3 * h / (10 + math.sqrt(4 - 3 * h)))


def measure_ac_mm(mask01: np.ndarray, spacing: Tuple[float, float]) -> float:
    """Abdominal circumference in mm: Ramanujan circumference of the ellipse fitted to the largest contour."""
    import cv2
    cnts, _ = cv2.findContours(mask01.astype(np.uint8), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
    if not cnts:
        return 0.0
    c = max(cnts, key=cv2.contourArea)
    if len(c) >= 5:
        (_, _), (MA, ma), _ = cv2.fitEllipse(c)
        return _ellipse_circum(MA / 2 * spacing[0], ma / 2 * spacing[1])
    return cv2.arcLength(c, True) * float(sum(spacing) / 2)


def convert_mask_2d_to_3d(mask: np.ndarray, frame: int, nf: int) -> np.ndarray:
    vol = np.zeros((nf,) + mask.shape, np.uint8)
    if 0 <= frame < nf:
        vol[frame] = (mask > 0).astype(np.uint8) * 2
    return vol


def write_output_mha_and_json(mask: np.ndarray, frame: int, ref: Path, od: Path, header: Optional[dict] = None) -> Path:
    """``<od>/<case>/images/fetal-abdomen-segmentation/output.mha`` (mask value 2 at ``frame``, geometry of the input) and
    ``<od>/<case>/fetal-abdomen-frame-number.json``."""
    ref = Path(ref)
    if header is None:
        _, header = read_mha(ref)
    nf = int(header["DimSize"].split()[2])
    spacing = tuple(float(t) for t in header.get("ElementSpacing", "1 1 1").split())
    cd = Path(od) / ref.stem
    (cd / "images/fetal-abdomen-segmentation").mkdir(parents=True, exist_ok=True)
    write_mha(cd / "images/fetal-abdomen-segmentation/output.mha", convert_mask_2d_to_3d(mask, frame, nf), spacing=spacing, compress=False)
    with open(cd / "fetal-abdomen-frame-number.json", "w") as f:
        json.dump(frame, f, indent=2)
    return cd


# ------------------------------------------------------------------------------------------------ device part
class PipelinePredictor:
    """Batched flip-TTA inference of conditioned frames at ``IMG_SIZE`` x ``IMG_SIZE``."""

    def __init__(self, net: AttentionASPPUNet, device: str | torch.device = "cuda", batch: int = 60, threads: Optional[int] = None):
        self.seg = FetalAbdomenSegmentation(net=net, device=device, batch=batch)
        self.net, self.device, self.batch = self.seg.net, self.seg.device, int(batch)
        self.threads = threads or min(32, os.cpu_count() or 1)

    def _logits(self, x: torch.Tensor) -> torch.Tensor:
        out = self.net(x)
        return out if isinstance(out, torch.Tensor) else out[0]

    @torch.no_grad()
    def predict_prob_tta(self, x: torch.Tensor) -> torch.Tensor:
        """``sigmoid((net(x) + flip(net(flip(x, [-1])), [-1])) / 2)`` for a device batch ``x``: uint8 ``[B,H,W]`` (values are
        divided by 255 by the network's first kernel) or float32 ``[B,1,H,W]``.  Returns float32 ``[B,H,W]`` on the device."""
        L, hnd = _capi.lib(), self.net.engine_handle()              # created by FetalAbdomenSegmentation (net.prepare)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        B, H, W = x.shape[0], x.shape[-2], x.shape[-1]
        xf = torch.empty_like(x)
        kind = _capi.AAU_X_U8 if x.dtype == torch.uint8 else _capi.AAU_X_F32
        _capi.check(hnd, L.aau_flip_w(hnd, x.data_ptr(), kind, B * H, W, xf.data_ptr(), stream), "aau_flip_w")
        l = self._logits(x)
        lf = self._logits(xf)
        prob = torch.empty((B, H, W), dtype=torch.float32, device=self.device)
        _capi.check(hnd, L.aau_tta_prob(hnd, l.data_ptr(), lf.data_ptr(), B * H, W, prob.data_ptr(), stream), "aau_tta_prob")
        return prob

    def condition(self, frames: np.ndarray) -> np.ndarray:
        """uint8 ``[N,H,W]`` conditioned frames (device kernels for uint8 input, the reference's host calls otherwise)."""
        if frames.dtype == np.uint8:
            out = np.empty_like(frames)
            for s in range(0, frames.shape[0], self.batch):
                out[s:s + self.batch] = self.seg.condition_on_device(torch.from_numpy(np.ascontiguousarray(frames[s:s + self.batch])).to(self.device)).cpu().numpy()
            return out
        return np.rint(preprocess_sweep(frames) * 255.0).astype(np.uint8)

    def predict_masks(self, frames: np.ndarray, thr: float = DEFAULT_THR) -> np.ndarray:
        """Refined uint8 {0,1} masks ``[N,H,W]`` of raw frames, following the reference slice loop (:492-501)."""
        import cv2
        n, H, W = frames.shape
        cond = self.condition(frames)
        with ThreadPoolExecutor(self.threads) as pool:
            small = np.stack(list(pool.map(lambda e: cv2.resize(e, (IMG_SIZE, IMG_SIZE), interpolation=cv2.INTER_LINEAR), cond)))
            prob = np.empty((n, IMG_SIZE, IMG_SIZE), np.float32)
            for s in range(0, n, self.batch):
                prob[s:s + self.batch] = self.predict_prob_tta(torch.from_numpy(small[s:s + self.batch]).to(self.device)).cpu().numpy()

            def finish(p):
                p = cv2.GaussianBlur(cv2.resize(p, (W, H)), (5, 5), 0)
                return refine_mask((p > thr).astype(np.uint8))
            return np.stack(list(pool.map(finish, prob)))

    def predict_case(self, sweep: np.ndarray, spacing_xy: Tuple[float, float], thr: float = DEFAULT_THR) -> dict:
        preds = self.predict_masks(sweep, thr)
        bf = select_best(preds, 5)
        return {"best_frame": bf, "mask": preds[bf], "ac_mm": round(measure_ac_mm(preds[bf], spacing_xy), 1), "masks": preds}


def predict(input_dir, out_dir, *, net: AttentionASPPUNet, spacing_json: Optional[str] = None, thr: Optional[float] = None,
            thr_config: str = "./checkpoints/thr.json", batch: int = 60) -> List[Tuple[str, int, float]]:
    """The reference CLI's ``predict(args)`` over a directory of ``.mha`` sweeps and ``.png/.jpg`` frames: masks, per-case
    output.mha + json and ``ac_results.csv`` (``case_id, frame_idx, ac_mm``) in ``out_dir``."""
    import cv2
    THR = DEFAULT_THR if thr is None else thr
    if thr is None and Path(thr_config).exists():
        try:
            THR = float(json.load(open(thr_config))["best_thr"])
        except Exception:
            pass
    spacing_map = {}
    if spacing_json:
        try:
            spacing_map = json.load(open(spacing_json))
        except Exception as e:
            print(f"cannot load spacing_json: {e}")

    def spacing_of(case_id):
        v = spacing_map.get(case_id)
        if isinstance(v, dict) and "spacing" in v:
            v = v["spacing"]
        return (float(v[0]), float(v[1])) if isinstance(v, (list, tuple)) and len(v) >= 2 else None

    pp = PipelinePredictor(net, batch=batch)
    od = Path(out_dir)
    od.mkdir(parents=True, exist_ok=True)
    rows: List[Tuple[str, int, float]] = []
    for p in sorted(Path(input_dir).iterdir()):
        ext = p.suffix.lower()
        if ext in {".png", ".jpg", ".jpeg"}:
            sl = cv2.imread(str(p), cv2.IMREAD_GRAYSCALE)
            mask = pp.predict_masks(sl[None], THR)[0]
            cv2.imwrite(str(od / f"{p.stem}_mask.png"), mask * 255)
            stem = p.stem
            case_id, frame_idx = stem, -1
            if "_s" in stem:
                case_id = stem.split("_s")[0]
                try:
                    frame_idx = int(stem.split("_s")[1])
                except Exception:
                    frame_idx = -1
            sp = spacing_of(case_id)
            if sp is None:
                print(f"no spacing for {case_id}, skip AC")
            else:
                rows.append((case_id, frame_idx, round(measure_ac_mm(mask, sp), 1)))
        elif ext == ".mha":
            vol, header = read_mha(p)
            sp = tuple(float(t) for t in header.get("ElementSpacing", "1 1 1").split())
            res = pp.predict_case(vol, (sp[0], sp[1]), THR)
            write_output_mha_and_json(res["mask"], res["best_frame"], p, od, header)
            rows.append((p.stem, int(res["best_frame"]), res["ac_mm"]))
            print(f"{p.stem}: best_frame={res['best_frame']}, AC={res['ac_mm']:.1f} mm")
    if rows:
        with open(od / "ac_results.csv", "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["case_id", "frame_idx", "ac_mm"])
            w.writerows(rows)
    return rows
