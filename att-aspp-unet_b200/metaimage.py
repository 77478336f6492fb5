"""Minimal MetaImage (.mha) reader / writer for ultrasound sweeps and segmentation volumes.

The reference reads sweeps with ``SimpleITK.ReadImage`` and writes the mask volume with
``SimpleITK.WriteImage(image, path, useCompression=True)`` after ``SetSpacing([0.28, 0.28, 0.28])``
(inference.py:91-95, 236-246).  SimpleITK is not part of this image, and the only things the hot path needs from it
are "give me the ``[frames, H, W]`` array of a single-file MetaImage" and "write a uint8 volume, zlib-compressed,
with a spacing": a text header of ``Key = Value`` lines ending in ``ElementDataFile = LOCAL`` followed by the raw
(or zlib-deflated) little-endian voxels in x-fastest order.  Files written here open in ITK / SimpleITK / ITK-SNAP;
files written by ITK with one local data block are read back.
"""
from __future__ import annotations

import zlib
from pathlib import Path
from typing import Dict, Sequence, Tuple

import numpy as np

_MET_TO_NP = {
    "MET_UCHAR": np.uint8, "MET_CHAR": np.int8, "MET_USHORT": np.uint16, "MET_SHORT": np.int16,
    "MET_UINT": np.uint32, "MET_INT": np.int32, "MET_ULONG_LONG": np.uint64, "MET_LONG_LONG": np.int64,
    "MET_FLOAT": np.float32, "MET_DOUBLE": np.float64,
}
_NP_TO_MET = {np.dtype(v): k for k, v in _MET_TO_NP.items()}


class MetaImageError(ValueError):
    pass


def read_mha(path) -> Tuple[np.ndarray, Dict[str, str]]:
    """Return ``(array, header)``.  The array is indexed ``[z, y, x]`` (``[frames, H, W]`` for a sweep), exactly as
    ``SimpleITK.GetArrayFromImage`` returns it; ``header`` holds the raw ``Key = Value`` strings."""
    raw = Path(path).read_bytes()
    header: Dict[str, str] = {}
    pos = 0
    while True:
        end = raw.find(b"\n", pos)
        if end < 0:
            raise MetaImageError(f"{path}: no ElementDataFile line in the MetaImage header")
        line = raw[pos:end].decode("latin-1").strip()
        pos = end + 1
        if not line:
            continue
        if "=" not in line:
            raise MetaImageError(f"{path}: malformed header line {line!r}")
        key, value = (t.strip() for t in line.split("=", 1))
        header[key] = value
        if key == "ElementDataFile":
            break
    if header["ElementDataFile"] != "LOCAL":
        raise MetaImageError(f"{path}: only single-file MetaImages (ElementDataFile = LOCAL) are supported")
    if header.get("ObjectType", "Image") != "Image":
        raise MetaImageError(f"{path}: ObjectType {header.get('ObjectType')!r} is not an image")
    ndims = int(header["NDims"])
    dims = [int(t) for t in header["DimSize"].split()]
    if len(dims) != ndims:
        raise MetaImageError(f"{path}: DimSize does not have NDims entries")
    et = header["ElementType"]
    if et not in _MET_TO_NP:
        raise MetaImageError(f"{path}: unsupported ElementType {et}")
    channels = int(header.get("ElementNumberOfChannels", "1"))
    dtype = np.dtype(_MET_TO_NP[et])
    if header.get("BinaryDataByteOrderMSB", header.get("ElementByteOrderMSB", "False")).lower() == "true":
        dtype = dtype.newbyteorder(">")
    payload = raw[pos:]
    if header.get("CompressedData", "False").lower() == "true":
        n = header.get("CompressedDataSize")
        payload = zlib.decompress(payload[: int(n)] if n else payload)
    count = int(np.prod(dims)) * channels
    if len(payload) < count * dtype.itemsize:
        raise MetaImageError(f"{path}: data block is shorter than DimSize says")
    arr = np.frombuffer(payload, dtype=dtype, count=count)
    shape = list(reversed(dims)) + ([channels] if channels > 1 else [])
    return arr.reshape(shape).astype(dtype.newbyteorder("="), copy=True), header        # own, writable memory


def write_mha(path, array: np.ndarray, spacing: Sequence[float] = (1.0, 1.0, 1.0), compress: bool = True,
              level: int = 2) -> int:
    """Write ``array`` (indexed ``[z, y, x]``) as a single-file MetaImage.  ``spacing`` is given in ITK order
    (x, y, z), as ``image.SetSpacing`` takes it.  Returns the number of bytes written."""
    arr = np.ascontiguousarray(array)
    if arr.dtype not in _NP_TO_MET:
        raise MetaImageError(f"unsupported dtype {arr.dtype}")
    if arr.dtype.byteorder == ">":
        arr = arr.astype(arr.dtype.newbyteorder("<"))
    nd = arr.ndim
    if len(spacing) != nd:
        raise MetaImageError("spacing must have one entry per dimension")
    data = arr.tobytes()
    eye = " ".join("1" if i == j else "0" for i in range(nd) for j in range(nd))
    zeros = " ".join("0" for _ in range(nd))
    lines = ["ObjectType = Image", f"NDims = {nd}", "BinaryData = True", "BinaryDataByteOrderMSB = False"]
    if compress:
        data = zlib.compress(data, level)
        lines += ["CompressedData = True", f"CompressedDataSize = {len(data)}"]
    else:
        lines += ["CompressedData = False"]
    lines += [f"TransformMatrix = {eye}", f"Offset = {zeros}", f"CenterOfRotation = {zeros}"]
    if nd == 3:
        lines.append("AnatomicalOrientation = RAI")
    lines += ["ElementSpacing = " + " ".join(repr(float(s)) for s in spacing),
              "DimSize = " + " ".join(str(d) for d in reversed(arr.shape)),
              f"ElementType = {_NP_TO_MET[arr.dtype]}", "ElementDataFile = LOCAL"]
    blob = ("\n".join(lines) + "\n").encode("ascii") + data
    Path(path).write_bytes(blob)
    return len(blob)
