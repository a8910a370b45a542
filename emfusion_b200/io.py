"""On-disk formats either side of the path -- host code, no device work:

 * TUM RGB-D container (`associations.txt` + 16-bit PNG depth scaled by 5000), as emf::TUMRGBDReader reads it
   (reference src/utils/TUMRGBDReader.cpp:38-104), plus a writer so that the synthetic stream of BASELINE.json's config 2
   can round-trip through the reference's own input format;
 * the raw volume dump of emf::EMFusion::writeVolume (src/core/EMFusion.cpp:1302-1313): int32 resolution[3], size_t element
   size, float voxel size, then the array -- and its inverse, which gives parity tests a fixture format the reference can
   produce.
The Co-Fusion reader (src/utils/ImageReader.cpp:41-116) takes OpenEXR depth; no EXR codec exists in this image, so only
its file naming (`Color%04d.png`, `Depth%04d.exr`) and its `depth > 100 -> 0` rule are restated (cofusion_names, cofusion_clean).
"""
from __future__ import annotations

import os
import struct
from typing import List, Sequence, Tuple

import numpy as np

TUM_DEPTH_SCALE = np.float32(1.0) / np.float32(5000.0)    # `1/5000.f` (TUMRGBDReader.cpp:101)


class TUMRGBDReader:
    """emf::TUMRGBDReader: parse `associations.txt`, hand out (rgb, depth in metres as float32) frames."""

    def __init__(self, path: str):
        self.path = path if path.endswith(os.sep) else path + os.sep
        self.rgbFileNames: List[str] = []
        self.depthFileNames: List[str] = []
        self.frameRate = 0.0
        self.minBufferSize = 0
        self._readFileAssociations(os.path.join(self.path, "associations.txt"))

    # -- src/utils/TUMRGBDReader.cpp:38-92
    def _readFileAssociations(self, filename: str):
        try:
            with open(filename, "r") as fh:
                lines = fh.read().split("\n")
        except OSError as e:
            raise RuntimeError("Could not open association file!") from e
        start = end = 0.0
        rgb_first = True
        first = True
        for line in lines:
            entries = line.replace("\t", " ").split(" ")      # boost::split on blanks and tabs (no token compression)
            if len(entries) != 4:
                first = False
                continue
            if first:
                rgb_first = entries[1].startswith("rgb/")
                start = float(entries[0])
                first = False
            else:
                end = float(entries[0])
            if rgb_first:
                self.rgbFileNames.append(entries[1]); self.depthFileNames.append(entries[3])
            else:
                self.rgbFileNames.append(entries[3]); self.depthFileNames.append(entries[1])
        if self.rgbFileNames and end != start:
            self.frameRate = len(self.rgbFileNames) / (end - start)
            self.minBufferSize = int(round(self.frameRate))

    def numFrames(self) -> int:
        return len(self.depthFileNames)

    # -- src/utils/TUMRGBDReader.cpp:94-104
    def readFrame(self, index: int) -> Tuple[np.ndarray, np.ndarray]:
        from PIL import Image
        rgb = np.asarray(Image.open(self.path + self.rgbFileNames[index]).convert("RGB"))[..., ::-1].copy()   # BGR, as cv::imread
        d16 = np.asarray(Image.open(self.path + self.depthFileNames[index]))
        if d16.dtype != np.uint16:
            d16 = d16.astype(np.uint16)
        depth = d16.astype(np.float32) * TUM_DEPTH_SCALE      # convertTo(CV_32FC1, 1/5000.f)
        return rgb, depth


def write_tum_stream(path: str, depths: Sequence[np.ndarray], rgbs: Sequence[np.ndarray] = None, fps: float = 30.0,
                     t0: float = 1305031102.175304) -> List[np.ndarray]:
    """Write frames as a TUM RGB-D sequence (rgb/<t>.png, depth/<t>.png 16-bit = round(metres * 5000), associations.txt).
    Returns the depth images as a reader will see them (quantised), for comparisons."""
    from PIL import Image
    os.makedirs(os.path.join(path, "rgb"), exist_ok=True)
    os.makedirs(os.path.join(path, "depth"), exist_ok=True)
    seen = []
    with open(os.path.join(path, "associations.txt"), "w") as fh:
        for i, d in enumerate(depths):
            t = t0 + i / fps
            name = f"{t:.6f}.png"
            q = np.clip(np.rint(np.asarray(d, dtype=np.float64) * 5000.0), 0, 65535).astype(np.uint16)
            Image.fromarray(q).save(os.path.join(path, "depth", name))
            rgb = rgbs[i] if rgbs is not None else np.zeros(d.shape + (3,), dtype=np.uint8)
            Image.fromarray(np.ascontiguousarray(rgb[..., ::-1])).save(os.path.join(path, "rgb", name))
            fh.write(f"{t:.6f} rgb/{name} {t:.6f} depth/{name}\n")
            seen.append(q.astype(np.float32) * TUM_DEPTH_SCALE)
    return seen


# -- src/core/EMFusion.cpp:1302-1313
def writeVolume(filename: str, vol: np.ndarray, resolution, voxelSize: float):
    """vol: (Rz*Ry, Rx[, channels]) array in the reference's layout; element size = bytes per voxel"""
    vol = np.ascontiguousarray(vol)
    n = int(resolution[0]) * int(resolution[1]) * int(resolution[2])
    elem = vol.nbytes // n
    with open(filename, "wb") as fh:
        fh.write(struct.pack("<3i", *[int(r) for r in resolution]))
        fh.write(struct.pack("<Q", elem))
        fh.write(struct.pack("<f", float(voxelSize)))
        fh.write(vol.tobytes())


def readVolume(filename: str):
    """-> (array (Rz*Ry, Rx[, channels]) float32, resolution (x, y, z), voxelSize)"""
    with open(filename, "rb") as fh:
        res = struct.unpack("<3i", fh.read(12))
        (elem,) = struct.unpack("<Q", fh.read(8))
        (voxel,) = struct.unpack("<f", fh.read(4))
        data = np.frombuffer(fh.read(), dtype=np.float32)
    ch = elem // 4
    if data.size != res[0] * res[1] * res[2] * ch:
        raise ValueError("truncated volume file")
    shape = (res[1] * res[2], res[0]) if ch == 1 else (res[1] * res[2], res[0], ch)
    return data.reshape(shape).copy(), res, voxel


# -- src/utils/ImageReader.cpp:41-116 (naming and the depth rule only; EXR decoding is not available here)
def cofusion_names(colorpath: str, depthpath: str, index: int) -> Tuple[str, str]:
    return os.path.join(colorpath, f"Color{index:04d}.png"), os.path.join(depthpath, f"Depth{index:04d}.exr")


def cofusion_clean(depth: np.ndarray) -> np.ndarray:
    d = np.asarray(depth, dtype=np.float32).copy()
    d[d > 100] = 0
    return d
