"""Sweep-level host layer: the reference's ``FetalAbdomenSegmentation`` / ``select_fetal_abdomen_mask_and_frame``
surface (model_attention_aspp.py:33-97) on top of the B200 engine.

Two ways in:

* :meth:`FetalAbdomenSegmentation.segment_sweep` -- the fast path.  A ``uint8 [N,H,W]`` sweep goes to the GPU
  in pinned, double-buffered batches; logits, per-frame areas and the arg-max never leave the device; only
  ``int32 areas[N]``, the best index and ONE ``uint8 [H,W]`` mask come back.  The 3x3 dilation + largest
  8-connected component of that single frame run on the host with scipy, exactly as the reference does
  (model_attention_aspp.py:80-85).
* :meth:`postprocess` / :func:`select_fetal_abdomen_mask_and_frame` -- signature-compatible mirrors taking host
  numpy volumes, for code written against the reference.  Their threshold / area / arg-max arithmetic still runs
  in the CUDA kernels (``aau_frame_scores``); there is no numpy fallback for it.

Frames of a sweep are independent, so a sweep (or a list of cases) is sharded across GPUs by giving every rank a
contiguous block of frames (``frame_range``); the only cross-rank step is a host gather of ``areas`` followed by
the same first-max arg-max (:func:`merge_shard_scores`).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

try:
    import _capi
    from attention_aspp_unet import AttentionASPPUNet
except ImportError:                                     # package-style import
    from . import _capi                                 # type: ignore
    from .attention_aspp_unet import AttentionASPPUNet  # type: ignore

__all__ = ["FetalAbdomenSegmentation", "select_fetal_abdomen_mask_and_frame", "merge_shard_scores", "largest_component",
           "preprocess_sweep", "load_image_file_as_array", "crop_roi_224", "logit_cutoff"]

ROI = 224                                               # model_attention_aspp.py:20-30
N_SAMPLED_FRAMES = 128                                  # model_attention_aspp.py:45


def preprocess_sweep(frames: np.ndarray) -> np.ndarray:
    """Per-frame min-max stretch to uint8, CLAHE (clip 1.0, 8x8 tiles), 3x3 median, ``/255`` -- the reference's
    input conditioning (model_attention_aspp.py:11-17, inference.py:147-190).  Host OpenCV calls, the same ones
    the reference makes (SURVEY.md section 8 f3 lists a GPU version as a later row).  Returns float32 [N,H,W]."""
    import cv2
    clahe = cv2.createCLAHE(clipLimit=1.0, tileGridSize=(8, 8))
    out = np.empty(frames.shape, np.float32)
    for i, sl in enumerate(frames):
        u8 = cv2.normalize(sl, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
        out[i] = cv2.medianBlur(clahe.apply(u8), 3)
    out /= 255.0
    return out


def load_image_file_as_array(*, location) -> np.ndarray:
    """``[1, N, H, W]`` float32 in [0,1] from a .mha sweep (model_attention_aspp.py:11-17)."""
    try:
        from metaimage import read_mha
    except ImportError:
        from .metaimage import read_mha                 # type: ignore
    arr, _ = read_mha(location)
    if arr.ndim != 3:
        raise ValueError(f"Expected 3-D image (frames, H, W), got {arr.shape}")
    return preprocess_sweep(arr)[None]


def crop_roi_224(img: np.ndarray):
    """224x224 patch centred on the centroid of the pixels brighter than 1.2x the frame mean, clamped to the
    frame, zero-padded when the frame is smaller (model_attention_aspp.py:20-30).  Returns ``(patch, (x0, y0))``."""
    h, w = img.shape
    ys, xs = np.where(img > img.mean() * 1.2)
    cx, cy = (w // 2, h // 2) if len(xs) == 0 else (int(xs.mean()), int(ys.mean()))
    x0, y0 = max(0, cx - ROI // 2), max(0, cy - ROI // 2)
    x0, y0 = min(x0, w - ROI), min(y0, h - ROI)
    patch = img[y0:y0 + ROI, x0:x0 + ROI]
    if patch.shape != (ROI, ROI):
        full = np.zeros((ROI, ROI), img.dtype)
        full[:patch.shape[0], :patch.shape[1]] = patch
        patch = full
    return patch, (x0, y0)


def largest_component(frame: np.ndarray) -> np.ndarray:
    """3x3 dilation, 8-connected labelling, keep the largest component (model_attention_aspp.py:80-85).

    Host integer work on ONE frame per sweep; scipy is what the reference itself calls."""
    import scipy.ndimage as ndi
    structure = np.ones((3, 3), dtype=np.uint8)
    frame = ndi.binary_dilation(frame, structure=structure, iterations=1)
    labeled, n = ndi.label(frame, structure=structure)
    if n:
        sizes = ndi.sum(frame, labeled, index=range(1, n + 1))
        frame = labeled == (np.argmax(sizes) + 1)
    return frame.astype(np.uint8)


def merge_shard_scores(shard_areas: Sequence[np.ndarray]) -> Tuple[np.ndarray, int]:
    """Host gather step of a multi-GPU sweep: concatenate the per-rank ``areas`` blocks (rank order == frame
    order) and take the first maximum, as ``areas.argmax()`` does on one device.  Returns ``(areas, idx)`` with
    ``idx == -1`` when every frame is empty (model_attention_aspp.py:95-96)."""
    areas = np.concatenate([np.asarray(a, dtype=np.int64) for a in shard_areas]) if len(shard_areas) else np.zeros(0, np.int64)
    if areas.size == 0 or areas.max() == 0:
        return areas, -1
    return areas, int(areas.argmax())


_CUTOFFS = {}


def logit_cutoff(prob_thr: float) -> float:
    """The fp32 logit ``c`` with ``torch.sigmoid(l) > prob_thr  <=>  l > c`` for THIS host's ``torch.sigmoid`` (the
    reference's own arithmetic, model_attention_aspp.py:54,71): sigmoid is monotone, so the threshold on the probability
    is a threshold on the logit, found once per threshold by bisection over the fp32 number line (SURVEY.md identity
    i7).  The device then only compares (``AAU_IN_LOGIT_CUT``): no transcendental per pixel, and the decision is the
    one the reference's fp32 sigmoid makes, bit for bit."""
    t = float(np.float32(prob_thr))
    c = _CUTOFFS.get(t)
    if c is not None:
        return c

    def to_float(k: int) -> float:                       # ordered integer image of the fp32 number line
        bits = k if k >= 0 else (0x80000000 | (-k))
        return float(np.array([bits], np.uint32).view(np.float32)[0])

    def passes(k: int) -> bool:                          # 64 copies: the vectorised ATen kernel, not its scalar tail
        x = torch.full((64,), to_float(k), dtype=torch.float32)
        return bool((torch.sigmoid(x).numpy() > np.float32(t))[0])

    lo, hi = -0x7f800000, 0x7f800000
    if passes(lo):
        c = float("-inf")
    elif not passes(hi):
        c = float("inf")
    else:
        while hi - lo > 1:
            mid = (lo + hi) // 2
            if passes(mid):
                hi = mid
            else:
                lo = mid
        c = to_float(lo)
    _CUTOFFS[t] = c
    return c


class _Scores:
    """Device-side threshold / area / arg-max through libaau (aau_frame_scores, aau_best_frame)."""

    def __init__(self, net: AttentionASPPUNet, device: torch.device):
        self.net, self.device = net, device
        net.prepare(device)

    def run(self, values: torch.Tensor, kind: int, thr: float, areas: torch.Tensor, best: Optional[torch.Tensor],
            mask: Optional[torch.Tensor]):
        n, h, w = values.shape
        if kind == _capi.AAU_IN_LOGITS:                   # threshold in logit space, calibrated on the host's sigmoid
            kind, thr = _capi.AAU_IN_LOGIT_CUT, logit_cutoff(thr)
        st = _capi.lib().aau_frame_scores(self.net.engine_handle(), values.data_ptr(), kind, n, h, w, C.c_float(thr), areas.data_ptr(),
                                          best.data_ptr() if best is not None else None, mask.data_ptr() if mask is not None else None,
                                          torch.cuda.current_stream(self.device).cuda_stream)
        _capi.check(self.net.engine_handle(), st, "aau_frame_scores")

    def best(self, areas: torch.Tensor, best: torch.Tensor):
        st = _capi.lib().aau_best_frame(self.net.engine_handle(), areas.data_ptr(), areas.numel(), best.data_ptr(),
                                        torch.cuda.current_stream(self.device).cuda_stream)
        _capi.check(self.net.engine_handle(), st, "aau_best_frame")


class FetalAbdomenSegmentation:
    """Mirror of the reference wrapper (model_attention_aspp.py:33-89) driving the B200 engine.

    ``net`` may be passed in (already holding weights) or is built as the wrapper does:
    ``AttentionASPPUNet(in_ch=1, num_classes=1, base=base)`` + ``load_state_dict(torch.load(path), strict=False)``.
    """

    PROB_THRESHOLD = 0.05                                  # model_attention_aspp.py:71

    DEFAULT_CHECKPOINT = "checkpoints/best_model.pth"      # model_attention_aspp.py:34

    def __init__(self, checkpoint_path: Optional[str] = None, *, net: Optional[AttentionASPPUNet] = None, base: int = 16,
                 device: str | torch.device = "cuda", batch: int = 56, act_dtype: str = "fp16"):
        if not torch.cuda.is_available():
            raise RuntimeError("FetalAbdomenSegmentation (B200 engine) needs a CUDA device; there is no CPU fallback")
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if net is None:
            # as the reference (model_attention_aspp.py:34-37): the checkpoint is mandatory -- `torch.load` raises when the
            # file is missing; a network with random weights is never built silently (pass `net=` to bring your own)
            path = self.DEFAULT_CHECKPOINT if checkpoint_path is None else checkpoint_path
            net = AttentionASPPUNet(in_ch=1, num_classes=1, base=base, act_dtype=act_dtype)
            miss, unexp = net.load_state_dict(torch.load(path, map_location="cpu"), strict=False)
            print(f"[DEBUG] load_state — missing:{len(miss)} unexpected:{len(unexp)}")
        self.net = net.eval()
        self.batch = int(batch)
        self._scores = _Scores(self.net, self.device)
        self._pinned = None
        self._staging = None
        self.last = {}

    # ------------------------------------------------------------------------------------------------
    def _net_logits(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        res = self.net(x, out=out)
        return res if isinstance(res, torch.Tensor) else res[0]

    def condition_on_device(self, frames_u8: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The reference's per-frame conditioning (min-max stretch, CLAHE(1.0, 8x8), 3x3 median) on device uint8 frames
        ``[n,H,W]``, bit exact with the OpenCV calls of model_attention_aspp.py:11-17 (``aau_condition_frames``).
        The result is what ``preprocess_sweep`` returns times 255, and is fed to the network as uint8."""
        assert frames_u8.dtype == torch.uint8 and frames_u8.is_cuda and frames_u8.dim() == 3 and frames_u8.is_contiguous()
        n, H, W = frames_u8.shape
        if out is None:
            out = torch.empty_like(frames_u8)
        L, hnd = _capi.lib(), self.net.engine_handle()
        need = L.aau_condition_workspace_bytes(hnd, n)
        if getattr(self, "_cond_ws", None) is None or self._cond_ws.numel() < need:
            self._cond_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        st = L.aau_condition_frames(hnd, frames_u8.data_ptr(), n, H, W, out.data_ptr(), self._cond_ws.data_ptr(), self._cond_ws.numel(),
                                    torch.cuda.current_stream(self.device).cuda_stream)
        _capi.check(hnd, st, "aau_condition_frames")
        return out

    # ------------------------------------------------------------------------------------------------
    # sweep driver: everything a sweep needs from the device is ENQUEUED by _enqueue_sweep (no host synchronisation), the
    # host tail (one D2H wait + connected components of ONE mask) runs in _finish_sweep.  segment_sweep = both back to back;
    # segment_sweeps pipelines them so that the host tail of sweep k overlaps the kernels of sweep k + 1.
    class _Slot:
        """Per-sweep device / pinned-host buffers (two slots alternate so that a finished sweep can still be read while
        the next one is running)."""

        def __init__(self):
            self.logits = self.scores = self.mask = self.h_scores = self.h_mask = None
            self.done = torch.cuda.Event()

    def _slot(self, which: int, n: int, H: int, W: int) -> "FetalAbdomenSegmentation._Slot":
        if getattr(self, "_slots", None) is None:
            self._slots = [None, None]
        sl = self._slots[which]
        if sl is None:
            sl = self._slots[which] = FetalAbdomenSegmentation._Slot()
        dev = self.device
        # every frame's logits stay resident (1.67 MB per 562x744 frame; 1.4 GB per 840-frame sweep of 180 GB)
        if sl.logits is None or sl.logits.shape != (max(n, 1), 1, H, W):
            sl.logits = torch.empty((max(n, 1), 1, H, W), dtype=torch.float32, device=dev)
            sl.scores = torch.zeros(max(n, 1) + 2, dtype=torch.int32, device=dev)   # areas[n] | {best index, best area}: ONE D2H copy
            sl.mask = torch.empty((H, W), dtype=torch.uint8, device=dev)
            sl.h_scores = torch.empty(max(n, 1) + 2, dtype=torch.int32).pin_memory()
            sl.h_mask = torch.empty((H, W), dtype=torch.uint8).pin_memory()
        return sl

    @torch.no_grad()
    def _enqueue_sweep(self, volume, frame_range, prob_thr, finalize, condition, which: int = 0) -> dict:
        thr = self.PROB_THRESHOLD if prob_thr is None else float(prob_thr)
        vol = torch.from_numpy(volume) if isinstance(volume, np.ndarray) else volume
        lo, hi = (0, vol.shape[0]) if frame_range is None else frame_range
        n, H, W = hi - lo, vol.shape[1], vol.shape[2]
        dev = self.device
        is_u8 = vol.dtype == torch.uint8
        if not is_u8:
            vol = vol.float()
        sl = self._slot(which, n, H, W)
        areas, best = sl.scores[: max(n, 1)], sl.scores[max(n, 1):]
        B = max(1, min(self.batch, n))
        self._logits_all = sl.logits
        # double-buffered pinned staging -> device input, copies on a side stream
        shape = (2, B, H, W) if is_u8 else (2, B, 1, H, W)
        main = torch.cuda.current_stream(dev)
        if self._pinned is None or self._pinned.shape != torch.Size(shape) or self._pinned.dtype != vol.dtype:
            main.synchronize()                                           # a sweep still in flight may be reading the old staging
            self._pinned = torch.empty(shape, dtype=vol.dtype).pin_memory()
            self._staging = torch.empty(shape, dtype=vol.dtype, device=dev)
            self._batch_no = 0
        if getattr(self, "_copy_stream", None) is None:                # stream and events live as long as the wrapper
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._events = [torch.cuda.Event() for _ in range(4)]
            self._batch_no = 0
        copy_stream = self._copy_stream
        ready, consumed = self._events[:2], self._events[2:]
        h2d = 0
        starts = list(range(0, n, B))

        def stage(i):
            s = starts[i]
            b = min(B, n - s)
            k = self._batch_no + i                                       # batches are numbered across sweeps: slot k & 1 was last used by batch k - 2
            slot = k & 1
            src = vol[lo + s: lo + s + b]
            if vol.is_pinned():
                host = src if is_u8 else src.unsqueeze(1)
            else:
                host = self._pinned[slot, :b]
                if k >= 2:
                    consumed[slot].synchronize()                         # the pinned slot must have been copied out
                host.copy_(src if is_u8 else src.unsqueeze(1))
            with torch.cuda.stream(copy_stream):
                if k >= 2:
                    copy_stream.wait_event(consumed[slot])               # the kernels that read this device slot are done
                self._staging[slot, :b].copy_(host, non_blocking=True)
                ready[slot].record(copy_stream)
            return b

        if starts:
            nb = stage(0)
        for i, s in enumerate(starts):
            b, slot = nb, (self._batch_no + i) & 1
            main.wait_event(ready[slot])
            x = self._staging[slot, :b]
            h2d += x.numel() * x.element_size()
            if condition:
                if not is_u8:
                    raise ValueError("device-side conditioning takes uint8 sweeps (use preprocess_sweep on the host otherwise)")
                if getattr(self, "_cond_out", None) is None or self._cond_out.shape != self._staging.shape[1:]:
                    self._cond_out = torch.empty(self._staging.shape[1:], dtype=torch.uint8, device=dev)
                x = self.condition_on_device(x, out=self._cond_out[:b])
            logits = self._net_logits(x, out=sl.logits[s: s + b])
            consumed[slot].record(main)
            if i + 1 < len(starts):
                nb = stage(i + 1)                                        # (after `consumed` of the batch two back has been recorded)
            self._scores.run(logits[:, 0], _capi.AAU_IN_LOGITS, thr, areas[s: s + b], None, None)
        self._batch_no += len(starts)
        job = {"n": n, "lo": lo, "H": H, "W": W, "thr": thr, "finalize": finalize, "slot": sl, "h2d": h2d,
               "launches": len(starts) * (self.net.num_launches() + 1) + 2}
        if n == 0:
            return job
        self._scores.best(areas[:n], best)
        if finalize:                                                     # the selected frame's mask, picked on the device
            hnd = self.net.engine_handle()
            st = _capi.lib().aau_best_frame_mask(hnd, sl.logits.data_ptr(), _capi.AAU_IN_LOGIT_CUT, n, H, W, C.c_float(logit_cutoff(thr)),
                                                 best.data_ptr(), sl.mask.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
            _capi.check(hnd, st, "aau_best_frame_mask")
            sl.h_mask.copy_(sl.mask, non_blocking=True)
        sl.h_scores.copy_(sl.scores, non_blocking=True)                  # D2H: 4*n + 8 bytes (+ H*W mask bytes), no host sync here
        sl.done.record(main)
        return job

    def _finish_sweep(self, job: dict) -> dict:
        n, lo, H, W, sl = job["n"], job["lo"], job["H"], job["W"], job["slot"]
        out = {"n_frames": n, "h2d_bytes": job["h2d"], "launches": job["launches"]}
        if n == 0:
            out.update(areas=np.zeros(0, np.int32), best_idx=-1, best_area=0, mask=np.zeros((H, W), np.uint8), d2h_bytes=0)
            return out
        sl.done.synchronize()
        host_scores = sl.h_scores.numpy()
        bi, ba = int(host_scores[-2]), int(host_scores[-1])
        out.update(areas=host_scores[:n].copy(), best_area=ba, best_local=bi, d2h_bytes=4 * n + 8)
        if not job["finalize"]:
            return out
        if ba == 0:
            out.update(best_idx=-1, mask=np.zeros((H, W), np.uint8))
            return out
        out["mask"] = largest_component(sl.h_mask.numpy())               # host integer work on ONE frame (model_attention_aspp.py:80-85)
        out["d2h_bytes"] += H * W
        out["best_idx"] = lo + bi
        self.last = out
        return out

    def segment_sweep(self, volume, frame_range: Optional[Tuple[int, int]] = None, prob_thr: Optional[float] = None,
                      finalize: bool = True, condition: bool = False):
        """Segment ``volume[frame_range]`` (``uint8`` or float ``[N,H,W]``, host numpy / pinned tensor) and select
        the best frame.  Returns a dict: ``areas`` (int32 numpy, this shard), ``best_idx`` (index into the full
        sweep, -1 if empty), ``best_area``, ``mask`` (uint8 [H,W], post-processed as the reference) and timing
        counters.  With ``finalize=False`` only ``areas`` are produced (multi-GPU shards: the caller gathers them,
        picks the global frame with :func:`merge_shard_scores` and asks the owner rank for :meth:`frame_mask`).
        ``condition=True`` (uint8 sweeps) runs the reference's frame conditioning on the device between the H2D copy
        and the network, so a RAW sweep goes in."""
        return self._finish_sweep(self._enqueue_sweep(volume, frame_range, prob_thr, finalize, condition, which=0))

    def segment_sweeps(self, volumes, prob_thr: Optional[float] = None, condition: bool = False):
        """Generator over the results of ``segment_sweep`` for a sequence of cases, software-pipelined: the device part of
        case k + 1 is enqueued before the host waits for case k, so the host tail (the D2H wait and the connected
        components of the selected frame) overlaps the next case's kernels.  Same results as calling ``segment_sweep`` per
        case (many-case inference, BASELINE config[2])."""
        pending = None
        for k, vol in enumerate(volumes):
            job = self._enqueue_sweep(vol, None, prob_thr, True, condition, which=k & 1)
            if pending is not None:
                yield self._finish_sweep(pending)
            pending = job
        if pending is not None:
            yield self._finish_sweep(pending)

    @torch.no_grad()
    def predict(self, input_img_path, save_probabilities: bool = False) -> np.ndarray:
        """The reference wrapper's ``predict`` (model_attention_aspp.py:41-65): read the sweep, condition every frame,
        sample 128 frames (``linspace``), crop a 224x224 ROI per frame, run the network + sigmoid, paste the ROI
        probabilities back (``cv2.resize`` to the ROI extent) into a zero ``[128, H, W]`` float32 volume.
        ``input_img_path`` is the list ``inference.py`` passes (first entry is used)."""
        import cv2
        from pathlib import Path
        path = Path(input_img_path[0] if isinstance(input_img_path, (list, tuple)) else input_img_path)
        self.case_id = path.stem
        vol = load_image_file_as_array(location=path)
        return self.predict_array(vol[0], save_probabilities=save_probabilities)

    @torch.no_grad()
    def predict_array(self, frames01: np.ndarray, save_probabilities: bool = False) -> np.ndarray:
        """``predict`` from an already conditioned float32 ``[N, H, W]`` volume in [0,1]."""
        import cv2
        from pathlib import Path
        idxs = np.linspace(0, frames01.shape[0] - 1, N_SAMPLED_FRAMES).astype(int)
        vol = frames01[idxs]
        N, H, W = vol.shape
        patches, coords = [], []
        for sl in vol:
            p, xy = crop_roi_224(sl)
            patches.append(p)
            coords.append(xy)
        x = torch.from_numpy(np.stack(patches)).unsqueeze(1).to(self.device)      # [128,1,224,224] float32
        prob = torch.empty((N, ROI, ROI), dtype=torch.float32, device=self.device)
        B = max(1, self.batch)
        for i in range(0, N, B):
            logits = self._net_logits(x[i:i + B])
            st = _capi.lib().aau_sigmoid(self.net.engine_handle(), logits.data_ptr(), logits.numel(), prob[i:i + B].data_ptr(),
                                         torch.cuda.current_stream(self.device).cuda_stream)
            _capi.check(self.net.engine_handle(), st, "aau_sigmoid")
        prob_roi = prob.cpu().numpy()
        prob_full = np.zeros((N, H, W), np.float32)
        for i, (x0, y0) in enumerate(coords):
            h_roi, w_roi = min(ROI, H - y0), min(ROI, W - x0)
            prob_full[i, y0:y0 + h_roi, x0:x0 + w_roi] = cv2.resize(prob_roi[i], (w_roi, h_roi))
        if save_probabilities:
            Path("output/probabilities").mkdir(parents=True, exist_ok=True)
            np.save(f"output/probabilities/{getattr(self, 'case_id', 'case')}_prob.npy", prob_full)
        return prob_full

    def _mask_from_logits(self, logits_1hw: torch.Tensor, thr: float) -> np.ndarray:
        _, H, W = logits_1hw.shape
        m = torch.empty((1, H, W), dtype=torch.uint8, device=self.device)
        a = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._scores.run(logits_1hw.contiguous(), _capi.AAU_IN_LOGITS, thr, a, None, m)
        return largest_component(m[0].cpu().numpy())

    @torch.no_grad()
    def frame_mask(self, volume, frame: int, prob_thr: Optional[float] = None) -> np.ndarray:
        """Post-processed mask of one frame (used by the owner rank after a multi-GPU gather)."""
        thr = self.PROB_THRESHOLD if prob_thr is None else float(prob_thr)
        vol = torch.from_numpy(volume) if isinstance(volume, np.ndarray) else volume
        x = vol[frame: frame + 1].to(self.device)
        x = x if x.dtype == torch.uint8 else x.float().unsqueeze(1)
        return self._mask_from_logits(self._net_logits(x)[:, 0], thr)

    # ------------------------------------------------------------------------------------------------
    # signature-compatible mirrors of the reference helpers (host numpy in / out)
    def postprocess(self, probability_map: np.ndarray, prob_thr: Optional[float] = None) -> np.ndarray:
        """model_attention_aspp.py:69-89: ``(prob > 0.05)``, frame with the largest area, all-zero volume if it is
        empty, else dilation + largest component of that frame; every other frame zero."""
        thr = self.PROB_THRESHOLD if prob_thr is None else float(prob_thr)
        prob = np.ascontiguousarray(probability_map, dtype=np.float32)
        n, H, W = prob.shape
        areas = torch.zeros(n, dtype=torch.int32, device=self.device)
        best = torch.zeros(2, dtype=torch.int32, device=self.device)
        step = max(1, (256 << 20) // (H * W * 4))
        for s in range(0, n, step):
            chunk = torch.from_numpy(prob[s: s + step]).to(self.device)
            self._scores.run(chunk, _capi.AAU_IN_PROB, thr, areas[s: s + chunk.shape[0]], None, None)
        self._scores.best(areas, best)
        bi, ba = [int(v) for v in best.cpu().numpy()]
        out = np.zeros((n, H, W), np.uint8)
        if ba == 0:
            return out
        m = torch.empty((1, H, W), dtype=torch.uint8, device=self.device)
        a = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._scores.run(torch.from_numpy(prob[bi: bi + 1]).to(self.device), _capi.AAU_IN_PROB, thr, a, None, m)
        out[bi] = largest_component(m[0].cpu().numpy())
        return out

    def select(self, mask_3d: np.ndarray):
        return select_fetal_abdomen_mask_and_frame(mask_3d, _engine=self)


def select_fetal_abdomen_mask_and_frame(mask_3d: np.ndarray, _engine: Optional[FetalAbdomenSegmentation] = None):
    """model_attention_aspp.py:91-97.  2-D input -> ``(mask > 0, 0)``; otherwise the frame with the largest
    ``sum`` (first on ties), ``(zeros, -1)`` when that sum is 0.  The per-frame sums and the arg-max run in the
    CUDA kernels (``aau_frame_scores`` with ``AAU_IN_U8`` + ``aau_best_frame``)."""
    mask_3d = np.asarray(mask_3d)
    if mask_3d.ndim == 2:
        return (mask_3d > 0).astype(np.uint8), 0
    if _engine is None:
        _engine = _default_engine()
    dev = _engine.device
    n, H, W = mask_3d.shape
    vol = torch.from_numpy(np.ascontiguousarray(mask_3d, dtype=np.uint8)).to(dev)
    areas = torch.zeros(n, dtype=torch.int32, device=dev)
    best = torch.zeros(2, dtype=torch.int32, device=dev)
    binm = torch.empty((n, H, W), dtype=torch.uint8, device=dev)
    _engine._scores.run(vol, _capi.AAU_IN_U8, 0.0, areas, best, binm)
    bi, ba = [int(v) for v in best.cpu().numpy()]
    if ba == 0:
        return np.zeros((H, W), np.uint8), -1
    return binm[bi].cpu().numpy(), bi


_DEFAULT = None


def _default_engine() -> FetalAbdomenSegmentation:
    global _DEFAULT
    if _DEFAULT is None:
        _DEFAULT = FetalAbdomenSegmentation(net=AttentionASPPUNet(base_c=16))
    return _DEFAULT
