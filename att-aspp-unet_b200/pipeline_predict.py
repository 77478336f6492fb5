"""The thesis' full-frame inference recipe on the B200 engine.

Same observable behaviour as ``predict`` and its helpers in the reference's pipeline script
(attention_aspp_unet_pipeline_stage.py:336-374, 399-523; the working frame selection is the ablation twin's,
test_ablation.py:389-403), organised around what runs where:

    device, batched over the sweep (libaau):
        min-max -> CLAHE(1.0, 8x8) -> median 3           aau_condition_frames      (bit exact with OpenCV)
        resize to 512x512                                 aau_resize_u8             (bit exact with cv2.resize INTER_LINEAR)
        sigmoid((net(x) + flip(net(flip(x)))) / 2)        aau_forward x2, aau_flip_w, aau_tta_prob
        resize back, GaussianBlur 5x5, > THR, area        aau_tail_masks
    host, per slice on a thread pool (integer morphology on uint8 masks -- the only thing that crosses PCIe):
        component filter + closing + hole filling          clean_mask
    host, per case:
        frame = most circular of the 5 largest masks, AC = Ramanujan perimeter of the fitted ellipse, outputs

``skimage.measure.label`` (absent here) is replaced by ``scipy.ndimage.label`` with full (8-) connectivity.
The reference's function names (``refine_mask``, ``select_best``, ``measure_ac_mm`` ...) are kept as aliases so code written
against it keeps working.
"""
from __future__ import annotations

import csv
import json
import math
import os
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

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
MIN_COMPONENT_PIXELS = 20                               # refine_mask: components below max(20, 0.15 % of the frame) are noise
MIN_COMPONENT_FRACTION = 0.0015
_EIGHT = np.ones((3, 3), np.uint8)


# ------------------------------------------------------------------------------------------------ host: mask clean-up
def clean_mask(mask: np.ndarray) -> np.ndarray:
    """Binary clean-up of one slice (reference ``refine_mask``): discard components smaller than
    ``max(20, 0.15 % of the frame)``, keep the largest survivor, close it with a 7x7 ellipse and fill its holes.

    One labelling pass is enough: removing whole components never merges the others, so the survivor the reference
    finds after re-labelling is simply the largest label that passed the size floor (first one on ties -- both
    labellings number components in raster order)."""
    import cv2
    import scipy.ndimage as ndi
    if not mask.any():
        return mask
    labels, count = ndi.label(mask, structure=_EIGHT)
    sizes = np.bincount(labels.ravel(), minlength=count + 1)
    sizes[0] = 0                                                          # background is not a component
    floor = max(MIN_COMPONENT_PIXELS, int(MIN_COMPONENT_FRACTION * mask.size))
    survivors = np.flatnonzero(sizes >= floor)
    if survivors.size == 0:
        return np.zeros_like(mask)
    winner = survivors[np.argmax(sizes[survivors])]
    blob = (labels == winner).astype(np.uint8)
    closed = cv2.morphologyEx(blob, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (7, 7)))
    return ndi.binary_fill_holes(closed).astype(np.uint8)


def _main_contour(mask: np.ndarray, approx: int):
    import cv2
    contours, _ = cv2.findContours(np.ascontiguousarray(mask, dtype=np.uint8), cv2.RETR_EXTERNAL, approx)
    return max(contours, key=cv2.contourArea) if contours else None


def circularity(mask: np.ndarray) -> float:
    """Isoperimetric quotient ``4 pi A / P^2`` of the largest outer contour (1 for a disc, 0 for an empty mask)."""
    import cv2
    outline = _main_contour(mask, cv2.CHAIN_APPROX_SIMPLE)
    if outline is None:
        return 0.0
    perimeter = cv2.arcLength(outline, True)
    if perimeter <= 1e-6:
        return 0.0
    return 4 * np.pi * cv2.contourArea(outline) / (perimeter ** 2)


def pick_frame(masks: Sequence[np.ndarray], topk: int = 5, areas: Optional[np.ndarray] = None) -> int:
    """Index of the most circular mask among the ``topk`` largest (reference ``select_best``); 0 for an empty stack."""
    if len(masks) == 0:
        return 0
    if areas is None:
        areas = np.array([np.count_nonzero(m) for m in masks])
    shortlist = areas.argsort()[::-1][: max(1, min(topk, len(areas)))]    # same ordering call as the reference: ties resolve alike
    scores = [circularity(masks[i]) for i in shortlist]
    return int(shortlist[int(np.argmax(scores))])                         # first best, as Python's max() over the shortlist


def ramanujan_perimeter(a: float, b: float) -> float:
    """Ramanujan's second approximation of the perimeter of an ellipse with semi-axes ``a``, ``b``."""
    h = (a - b) ** 2 / (a + b) ** 2
    return math.pi * (a + b) * (1 + 3 * h / (10 + math.sqrt(4 - 3 * h)))


def abdominal_circumference_mm(mask01: np.ndarray, spacing: Tuple[float, float]) -> float:
    """AC in mm (reference ``measure_ac_mm``): perimeter of the ellipse fitted to the largest contour, semi-axes scaled by
    the pixel spacing; contours of fewer than five points fall back to their arc length at the mean spacing."""
    import cv2
    outline = _main_contour(mask01, cv2.CHAIN_APPROX_NONE)
    if outline is None:
        return 0.0
    if len(outline) < 5:
        return cv2.arcLength(outline, True) * float(sum(spacing) / 2)
    _centre, (d0, d1), _angle = cv2.fitEllipse(outline)
    return ramanujan_perimeter(d0 / 2 * spacing[0], d1 / 2 * spacing[1])


def mask_volume(mask: np.ndarray, frame: int, n_frames: int) -> np.ndarray:
    """``[n_frames, H, W]`` uint8, label 2 on the selected frame, zero elsewhere (all zero for an out-of-range frame)."""
    volume = np.zeros((n_frames,) + mask.shape, np.uint8)
    if 0 <= frame < n_frames:
        np.multiply(mask > 0, 2, out=volume[frame], casting="unsafe")
    return volume


def write_case_outputs(mask: np.ndarray, frame: int, sweep_path, out_dir, header: Optional[dict] = None) -> Path:
    """``<out_dir>/<case>/images/fetal-abdomen-segmentation/output.mha`` (geometry of the input sweep, uncompressed, as the
    reference's ``sitk.WriteImage(out, path, False)``) and ``<out_dir>/<case>/fetal-abdomen-frame-number.json``."""
    sweep_path = Path(sweep_path)
    if header is None:
        header = read_mha(sweep_path)[1]
    n_frames = int(header["DimSize"].split()[2])
    spacing = tuple(float(v) for v in header.get("ElementSpacing", "1 1 1").split())
    case_dir = Path(out_dir) / sweep_path.stem
    seg_dir = case_dir / "images" / "fetal-abdomen-segmentation"
    seg_dir.mkdir(parents=True, exist_ok=True)
    write_mha(seg_dir / "output.mha", mask_volume(mask, frame, n_frames), spacing=spacing, compress=False)
    (case_dir / "fetal-abdomen-frame-number.json").write_text(json.dumps(frame, indent=2))
    return case_dir


# reference spellings (attention_aspp_unet_pipeline_stage.py:340-374,383-397; test_ablation.py:373-418)
refine_mask = clean_mask
_circularity_score = circularity
select_best = pick_frame
_ellipse_circum = ramanujan_perimeter
measure_ac_mm = abdominal_circumference_mm
convert_mask_2d_to_3d = mask_volume
write_output_mha_and_json = write_case_outputs


# ------------------------------------------------------------------------------------------------ device part
class PipelinePredictor:
    """Batched slice loop of the CLI: conditioning, resize, flip-TTA inference, resize back, blur and threshold on the
    device; the binary masks come back for the host clean-up."""

    def __init__(self, net: AttentionASPPUNet, device: str | torch.device = "cuda", batch: int = 60, threads: Optional[int] = None):
        self.seg = FetalAbdomenSegmentation(net=net, device=device, batch=batch)
        self.net, self.device, self.batch = self.seg.net, self.seg.device, int(batch)
        self.threads = threads or min(32, os.cpu_count() or 1)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _logits(self, x: torch.Tensor) -> torch.Tensor:
        out = self.net(x)
        return out if isinstance(out, torch.Tensor) else out[0]

    @torch.no_grad()
    def predict_prob_tta(self, x: torch.Tensor) -> torch.Tensor:
        """``sigmoid((net(x) + flip(net(flip(x, [-1])), [-1])) / 2)`` for a device batch ``x``: uint8 ``[B,H,W]`` (values are
        divided by 255 by the network's first kernel) or float32 ``[B,1,H,W]``.  Returns float32 ``[B,H,W]`` on the device."""
        L, hnd = _capi.lib(), self.net.engine_handle()              # created by FetalAbdomenSegmentation (net.prepare)
        B, H, W = x.shape[0], x.shape[-2], x.shape[-1]
        mirrored = torch.empty_like(x)
        kind = _capi.AAU_X_U8 if x.dtype == torch.uint8 else _capi.AAU_X_F32
        _capi.check(hnd, L.aau_flip_w(hnd, x.data_ptr(), kind, B * H, W, mirrored.data_ptr(), self._stream()), "aau_flip_w")
        plain, flipped = self._logits(x), self._logits(mirrored)
        prob = torch.empty((B, H, W), dtype=torch.float32, device=self.device)
        _capi.check(hnd, L.aau_tta_prob(hnd, plain.data_ptr(), flipped.data_ptr(), B * H, W, prob.data_ptr(), self._stream()), "aau_tta_prob")
        return prob

    def resize_on_device(self, frames_u8: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
        """``cv2.resize(frame, (w, h), INTER_LINEAR)`` of device uint8 frames ``[n,H,W]`` (bit exact, ``aau_resize_u8``)."""
        n, H, W = frames_u8.shape
        out = torch.empty((n, size[0], size[1]), dtype=torch.uint8, device=self.device)
        hnd = self.net.engine_handle()
        _capi.check(hnd, _capi.lib().aau_resize_u8(hnd, frames_u8.data_ptr(), n, H, W, out.data_ptr(), size[0], size[1], self._stream()), "aau_resize_u8")
        return out

    def tail_on_device(self, prob: torch.Tensor, size: Tuple[int, int], thr: float) -> Tuple[torch.Tensor, torch.Tensor]:
        """``(cv2.GaussianBlur(cv2.resize(prob, (w, h)), (5, 5), 0) > thr)`` per slice -> (uint8 masks ``[n,h,w]``, int32
        areas ``[n]``), both on the device (``aau_tail_masks``)."""
        n, PH, PW = prob.shape
        masks = torch.empty((n, size[0], size[1]), dtype=torch.uint8, device=self.device)
        areas = torch.empty(n, dtype=torch.int32, device=self.device)
        hnd = self.net.engine_handle()
        import ctypes as C
        _capi.check(hnd, _capi.lib().aau_tail_masks(hnd, prob.data_ptr(), n, PH, PW, size[0], size[1], C.c_float(thr), masks.data_ptr(),
                                                   areas.data_ptr(), self._stream()), "aau_tail_masks")
        return masks, areas

    def condition(self, frames: np.ndarray) -> np.ndarray:
        """uint8 ``[N,H,W]`` conditioned frames on the host (device kernels for uint8 input, the reference's host calls otherwise)."""
        if frames.dtype == np.uint8:
            out = np.empty_like(frames)
            for s in range(0, frames.shape[0], self.batch):
                out[s:s + self.batch] = self.seg.condition_on_device(torch.from_numpy(np.ascontiguousarray(frames[s:s + self.batch])).to(self.device)).cpu().numpy()
            return out
        return np.rint(preprocess_sweep(frames) * 255.0).astype(np.uint8)

    @torch.no_grad()
    def raw_masks(self, frames: np.ndarray, thr: float = DEFAULT_THR) -> Tuple[np.ndarray, np.ndarray]:
        """Thresholded (not yet cleaned) masks ``[N,H,W]`` uint8 and their areas for raw frames: the device part of the
        slice loop (:492-498).  Only the frames go up and only the uint8 masks come down."""
        n, H, W = frames.shape
        masks = np.empty((n, H, W), np.uint8)
        areas = np.empty(n, np.int32)
        on_device = frames.dtype == np.uint8
        if not on_device:                                           # other voxel types: the reference's own host conditioning
            frames = self.condition(frames)
        for s in range(0, n, self.batch):
            x = torch.from_numpy(np.ascontiguousarray(frames[s:s + self.batch])).to(self.device)
            if on_device:
                x = self.seg.condition_on_device(x)
            small = self.resize_on_device(x, (IMG_SIZE, IMG_SIZE))
            m, a = self.tail_on_device(self.predict_prob_tta(small), (H, W), thr)
            masks[s:s + self.batch] = m.cpu().numpy()
            areas[s:s + self.batch] = a.cpu().numpy()
        return masks, areas

    def predict_masks(self, frames: np.ndarray, thr: float = DEFAULT_THR) -> np.ndarray:
        """Cleaned uint8 {0,1} masks ``[N,H,W]`` of raw frames, following the reference slice loop (:492-501)."""
        raw, _ = self.raw_masks(frames, thr)
        with ThreadPoolExecutor(self.threads) as pool:
            return np.stack(list(pool.map(clean_mask, raw)))

    def predict_case(self, sweep: np.ndarray, spacing_xy: Tuple[float, float], thr: float = DEFAULT_THR) -> dict:
        masks = self.predict_masks(sweep, thr)
        frame = pick_frame(masks, 5)
        return {"best_frame": frame, "mask": masks[frame], "ac_mm": round(abdominal_circumference_mm(masks[frame], spacing_xy), 1), "masks": masks}


# ------------------------------------------------------------------------------------------------ CLI-level driver
def _threshold(thr: Optional[float], thr_config: str) -> float:
    """Explicit value, else ``best_thr`` of the calibration file when it parses, else 0.48 (:405-411)."""
    if thr is not None:
        return thr
    try:
        return float(json.loads(Path(thr_config).read_text())["best_thr"])
    except Exception:
        return DEFAULT_THR


def _load_spacings(spacing_json: Optional[str]) -> Dict[str, Tuple[float, float]]:
    """``{case_id: (sx, sy)}`` from the optional json (values are ``[sx, sy]`` or ``{"spacing": [sx, sy]}``, :413-432)."""
    table: Dict[str, Tuple[float, float]] = {}
    if not spacing_json:
        return table
    try:
        raw = json.loads(Path(spacing_json).read_text())
    except Exception as err:
        print(f"cannot load spacing_json: {err}")
        return table
    for case_id, value in raw.items():
        if isinstance(value, dict):
            value = value.get("spacing")
        if isinstance(value, (list, tuple)) and len(value) >= 2:
            table[case_id] = (float(value[0]), float(value[1]))
    return table


def _split_frame_name(stem: str) -> Tuple[str, int]:
    """``<case>_s<frame>`` -> (case, frame); anything else -> (stem, -1)  (:463-470)."""
    case_id, sep, tail = stem.partition("_s")
    if not sep:
        return stem, -1
    try:
        return case_id, int(tail.split("_s")[0])
    except ValueError:
        return case_id, -1


def predict(input_dir, out_dir, *, net: AttentionASPPUNet, spacing_json: Optional[str] = None, thr: Optional[float] = None,
            thr_config: str = "./checkpoints/thr.json", batch: int = 60) -> List[Tuple[str, int, float]]:
    """The reference CLI's ``predict(args)`` over a directory of ``.mha`` sweeps and ``.png/.jpg`` frames: masks, per-case
    output.mha + json and ``ac_results.csv`` (``case_id, frame_idx, ac_mm``) in ``out_dir``."""
    import cv2
    cut = _threshold(thr, thr_config)
    spacings = _load_spacings(spacing_json)
    engine = PipelinePredictor(net, batch=batch)
    out_dir = Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    results: List[Tuple[str, int, float]] = []
    for path in sorted(Path(input_dir).iterdir()):
        kind = path.suffix.lower()
        if kind in (".png", ".jpg", ".jpeg"):
            frame = cv2.imread(str(path), cv2.IMREAD_GRAYSCALE)
            mask = engine.predict_masks(frame[None], cut)[0]
            cv2.imwrite(str(out_dir / f"{path.stem}_mask.png"), mask * 255)
            case_id, frame_idx = _split_frame_name(path.stem)
            if case_id in spacings:
                results.append((case_id, frame_idx, round(abdominal_circumference_mm(mask, spacings[case_id]), 1)))
            else:
                print(f"no spacing for {case_id}, skip AC")
        elif kind == ".mha":
            sweep, header = read_mha(path)
            sx, sy = [float(v) for v in header.get("ElementSpacing", "1 1 1").split()][:2]
            case = engine.predict_case(sweep, (sx, sy), cut)
            write_case_outputs(case["mask"], case["best_frame"], path, out_dir, header)
            results.append((path.stem, int(case["best_frame"]), case["ac_mm"]))
            print(f"{path.stem}: best_frame={case['best_frame']}, AC={case['ac_mm']:.1f} mm")
    if results:
        with open(out_dir / "ac_results.csv", "w", newline="") as f:
            writer = csv.writer(f)
            writer.writerow(["case_id", "frame_idx", "ac_mm"])
            writer.writerows(results)
    return results
