"""Container entry point: sweep in, ``<case>.mha`` mask volume + ``fetal-abdomen-frame-number.json`` out.

Mirror of the reference's ``inference.py`` (run(): :50-133, write_array_as_image_file: :208-254,
convert_2d_mask_to_3d: :257-273, write_json_file: :136-139, get_image_file_path: :196-199) driving the B200
engine.  Environment: ``CASE_ID`` (output file stem, default ``output``), ``AAU_CHECKPOINT`` (state dict to load; default ``checkpoints/best_model.pth`` as the reference, a missing file raises),
``AAU_BASE_C`` (default 16, the wrapper's ``base``), ``AAU_SWEEP_MODE``:

* ``roi224`` (default) -- the reference wrapper's recipe: 128 sampled frames, 224x224 ROI per frame, paste back
  (``FetalAbdomenSegmentation.predict`` / ``postprocess`` / ``select_fetal_abdomen_mask_and_frame``).
* ``full`` -- every frame of the sweep at full resolution through ``segment_sweep`` (the path bench.py measures);
  the frame number then indexes the original sweep.

Outputs follow the reference byte semantics: uint8 volume ``[n_frames, H, W]`` with the selected frame's mask as
{0,1} and every other frame zero, spacing 0.28^3, zlib-compressed MetaImage; the JSON file holds the bare integer
(``-1`` and an all-zero volume when nothing was segmented).
"""
from __future__ import annotations

import json
import os
from glob import glob
from pathlib import Path

import numpy as np

try:
    from metaimage import read_mha, write_mha
    from fetal_abdomen import FetalAbdomenSegmentation, select_fetal_abdomen_mask_and_frame, preprocess_sweep
except ImportError:                                      # package-style import
    from .metaimage import read_mha, write_mha          # type: ignore
    from .fetal_abdomen import FetalAbdomenSegmentation, select_fetal_abdomen_mask_and_frame, preprocess_sweep  # type: ignore

INPUT_PATH = Path("./test/input")
OUTPUT_PATH = Path("./test/output")
SPACING = (0.28, 0.28, 0.28)


def get_image_file_path(*, location) -> list:
    location = Path(location)
    return glob(str(location / "*.tiff")) + glob(str(location / "*.mha"))


def write_json_file(*, location, content) -> None:
    Path(location).parent.mkdir(parents=True, exist_ok=True)
    with open(location, "w") as f:
        f.write(json.dumps(content, indent=4))


def convert_2d_mask_to_3d(*, mask_2d: np.ndarray, frame_number, number_of_frames: int) -> np.ndarray:
    """Zero volume with ``mask_2d`` (values 1 -> 2, anything else -> 0) at ``frame_number``; ``-1`` gives the
    all-zero volume, any other out-of-range or ``None`` frame number is a ``ValueError``."""
    marked = np.where(mask_2d == 1, 2, 0).astype(np.uint8)
    volume = np.zeros((number_of_frames,) + marked.shape, dtype=np.uint8)
    if frame_number == -1:
        return volume
    if frame_number is None or not (0 <= frame_number < number_of_frames):
        raise ValueError(f"frame_number must be between -1 and {number_of_frames - 1}, got {frame_number}.")
    volume[frame_number] = marked
    return volume


def write_array_as_image_file(*, location, array: np.ndarray, frame_number=None, number_of_frames: int = 128,
                              filename: str = "output.mha") -> Path:
    location = Path(location)
    location.mkdir(parents=True, exist_ok=True)
    array = np.squeeze(array)
    assert array.ndim == 2, f"Expected a 2D array, got {array.ndim}D."
    volume = convert_2d_mask_to_3d(mask_2d=array.astype(np.float32), frame_number=frame_number, number_of_frames=number_of_frames)
    volume = np.where(volume > 0.5, 1, 0).astype(np.uint8)
    assert set(np.unique(volume)).issubset({0, 1})
    write_mha(location / filename, volume, spacing=SPACING, compress=True)
    return location / filename


def _nearest_resize(mask: np.ndarray, h: int, w: int) -> np.ndarray:
    import cv2
    return cv2.resize(mask.astype("uint8"), (w, h), interpolation=cv2.INTER_NEAREST)


def run(case_id: str | None = None, *, algorithm: FetalAbdomenSegmentation | None = None, input_path=None, output_path=None,
        mode: str | None = None) -> int:
    case_id = case_id or os.getenv("CASE_ID", "output")
    input_path = Path(input_path or INPUT_PATH)
    output_path = Path(output_path or OUTPUT_PATH)
    mode = mode or os.getenv("AAU_SWEEP_MODE", "roi224")
    paths = get_image_file_path(location=input_path / "images/stacked-fetal-ultrasound")
    if not paths:
        raise FileNotFoundError(f"no .mha / .tiff sweep under {input_path / 'images/stacked-fetal-ultrasound'}")
    if algorithm is None:
        # AAU_CHECKPOINT unset -> the reference's default path; a missing file raises (never random weights)
        algorithm = FetalAbdomenSegmentation(os.getenv("AAU_CHECKPOINT") or FetalAbdomenSegmentation.DEFAULT_CHECKPOINT,
                                             base=int(os.getenv("AAU_BASE_C", "16")))
    sweep, _ = read_mha(paths[0])
    n_frames, ref_h, ref_w = sweep.shape
    if mode == "full":
        if sweep.dtype == np.uint8:                                 # raw sweep in, conditioning on the device (bit exact with OpenCV)
            res = algorithm.segment_sweep(np.ascontiguousarray(sweep), condition=True)
        else:                                                       # other voxel types: the reference's own host calls
            conditioned = preprocess_sweep(sweep)
            res = algorithm.segment_sweep(np.ascontiguousarray(np.rint(conditioned * 255.0).astype(np.uint8)))
        segmentation, frame_number = res["mask"], res["best_idx"]
    else:
        prob = algorithm.predict(paths, save_probabilities=bool(int(os.getenv("AAU_SAVE_PROB", "0"))))
        post = algorithm.postprocess(prob)
        segmentation, frame_number = select_fetal_abdomen_mask_and_frame(post, _engine=algorithm)
    if segmentation.shape != (ref_h, ref_w):
        segmentation = _nearest_resize(segmentation, ref_h, ref_w)
    segmentation = (segmentation > 0).astype("uint8")
    write_array_as_image_file(location=output_path / "images/fetal-abdomen-segmentation", array=segmentation,
                              frame_number=frame_number, number_of_frames=n_frames, filename=f"{case_id}.mha")
    write_json_file(location=output_path / "fetal-abdomen-frame-number.json", content=int(frame_number))
    print(f"[inference] case {case_id}: frame {frame_number}, {int(segmentation.sum())} mask pixels, mode {mode}")
    return 0


if __name__ == "__main__":
    raise SystemExit(run())
