"""On-disk formats either side of the path -- host code, no device work:

 * TUM RGB-D container (`associations.txt` + 16-bit PNG depth scaled by 5000), as emf::TUMRGBDReader reads it
   (reference src/utils/TUMRGBDReader.cpp:38-104), plus a writer so that the synthetic stream of BASELINE.json's config 2
   can round-trip through the reference's own input format;
 * the raw volume dump of emf::EMFusion::writeVolume (src/core/EMFusion.cpp:1302-1313): int32 resolution[3], size_t element
   size, float voxel size, then the array -- and its inverse, which gives parity tests a fixture format the reference can
   produce.
The Co-Fusion reader (src/utils/ImageReader.cpp:41-116: `Color%04d.png`, `Depth%04d.exr` float metres, `depth > 100 -> 0`)
takes OpenEXR depth through cv::imread.  No EXR codec exists in this image, so a minimal one is written out here
(read_exr / write_exr: single-part scan-line files, FLOAT or HALF channels, compression NONE / ZIPS / ZIP -- what depth
dumps use); ImageReader mirrors the reference class on top of it.
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


# ------------------------------------------------------------------------------------------------
# OpenEXR, the subset depth images use: single-part scan-line files; channels FLOAT (32 bit) or HALF (16 bit);
# compression NONE (0), ZIPS (2: one scan line per chunk) or ZIP (3: 16 scan lines per chunk).
# Layout (OpenEXR file layout specification): magic 20000630, version word, attributes `name\0 type\0 size value`
# terminated by \0, one uint64 offset per chunk, chunks `y, size, data`; inside a chunk the scan lines follow each other,
# each holding its channels in alphabetical order; ZIP data is zlib-deflated after a byte reordering (even / odd bytes
# split) and a delta predictor.
# ------------------------------------------------------------------------------------------------
_EXR_MAGIC = 20000630


def _exr_unzip(blob: bytes, raw_size: int) -> bytes:
    import zlib
    if len(blob) >= raw_size:           # stored uncompressed when deflate does not help
        return blob
    t = np.frombuffer(zlib.decompress(blob), dtype=np.uint8).astype(np.int64)
    if t.size != raw_size:
        raise ValueError("EXR: chunk inflates to the wrong size")
    t = (np.cumsum(t) - 128 * np.arange(t.size)) & 0xFF        # predictor: t[i] = t[i-1] + t[i] - 128
    t = t.astype(np.uint8)
    half = (raw_size + 1) // 2
    out = np.empty(raw_size, dtype=np.uint8)
    out[0::2] = t[:half]
    out[1::2] = t[half:]
    return out.tobytes()


def _exr_zip(raw: bytes) -> bytes:
    import zlib
    a = np.frombuffer(raw, dtype=np.uint8)
    t = np.concatenate([a[0::2], a[1::2]]).astype(np.int64)
    d = t.copy()
    d[1:] = (t[1:] - t[:-1] + 128 + 256) & 0xFF
    blob = zlib.compress(d.astype(np.uint8).tobytes(), 6)
    return blob if len(blob) < len(raw) else raw


def read_exr(filename: str, channel: str = None) -> np.ndarray:
    """-> (H, W) float32: the named channel, or the file's only channel, or the first of Z / Y / R / G / B that exists."""
    with open(filename, "rb") as fh:
        buf = fh.read()
    magic, version = struct.unpack_from("<iI", buf, 0)
    if magic != _EXR_MAGIC:
        raise ValueError("not an OpenEXR file")
    if version & 0x1A00:
        raise ValueError("EXR: tiled / deep / multi-part files are not supported")
    pos = 8
    attrs = {}
    while buf[pos] != 0:
        e = buf.index(b"\0", pos); name = buf[pos:e].decode(); pos = e + 1
        e = buf.index(b"\0", pos); typ = buf[pos:e].decode(); pos = e + 1
        (size,) = struct.unpack_from("<i", buf, pos); pos += 4
        attrs[name] = (typ, buf[pos:pos + size]); pos += size
    pos += 1
    chans = []                                  # (name, pixel type, x sampling, y sampling), in file (alphabetical) order
    cb = attrs["channels"][1]
    q = 0
    while cb[q] != 0:
        e = cb.index(b"\0", q); cname = cb[q:e].decode(); q = e + 1
        ptype, _, xs, ys = struct.unpack_from("<iIii", cb, q); q += 16
        chans.append((cname, ptype, xs, ys))
    comp = attrs["compression"][1][0]
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    lines_per_chunk = {0: 1, 2: 1, 3: 16}.get(comp)
    if lines_per_chunk is None:
        raise ValueError(f"EXR: compression {comp} is not supported (NONE, ZIPS, ZIP are)")
    if any(xs != 1 or ys != 1 for _, _, xs, ys in chans):
        raise ValueError("EXR: sub-sampled channels are not supported")
    names = [c[0] for c in chans]
    if channel is None:
        channel = names[0] if len(names) == 1 else next((c for c in ("Z", "Y", "R", "G", "B") if c in names), names[0])
    if channel not in names:
        raise ValueError(f"EXR: no channel {channel!r} (has {names})")
    bpp = {0: 4, 1: 2, 2: 4}
    line_bytes = sum(bpp[c[1]] * w for c in chans)
    n_chunks = (h + lines_per_chunk - 1) // lines_per_chunk
    offsets = struct.unpack_from(f"<{n_chunks}Q", buf, pos)
    out = np.zeros((h, w), dtype=np.float32)
    for off in offsets:
        y, size = struct.unpack_from("<ii", buf, off)
        n_lines = min(lines_per_chunk, y1 - y + 1)
        raw = buf[off + 8:off + 8 + size]
        if comp != 0:
            raw = _exr_unzip(raw, n_lines * line_bytes)
        for ln in range(n_lines):
            o = ln * line_bytes
            for cname, ptype, _, _ in chans:
                nb = bpp[ptype] * w
                if cname == channel:
                    dt = {0: "<u4", 1: "<f2", 2: "<f4"}[ptype]
                    out[y - y0 + ln] = np.frombuffer(raw, dtype=dt, count=w, offset=o).astype(np.float32)
                o += nb
    return out


def write_exr(filename: str, img: np.ndarray, channel: str = "Y", compression: int = 3):
    """(H, W) float32 -> single-channel FLOAT scan-line EXR (compression 0 NONE, 2 ZIPS, 3 ZIP).  Channel "Y": what cv::imread
    (the reference's reader, src/utils/ImageReader.cpp:108) turns into a CV_32FC1 image."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    h, w = img.shape
    lpc = {0: 1, 2: 1, 3: 16}[compression]

    def attr(name, typ, val):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(val)) + val

    head = struct.pack("<iI", _EXR_MAGIC, 2)
    head += attr("channels", "chlist", channel.encode() + b"\0" + struct.pack("<iIii", 2, 0, 1, 1) + b"\0")
    head += attr("compression", "compression", bytes([compression]))
    head += attr("dataWindow", "box2i", struct.pack("<4i", 0, 0, w - 1, h - 1))
    head += attr("displayWindow", "box2i", struct.pack("<4i", 0, 0, w - 1, h - 1))
    head += attr("lineOrder", "lineOrder", bytes([0]))
    head += attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
    head += attr("screenWindowCenter", "v2f", struct.pack("<2f", 0.0, 0.0))
    head += attr("screenWindowWidth", "float", struct.pack("<f", 1.0))
    head += b"\0"
    chunks = []
    for y in range(0, h, lpc):
        raw = img[y:y + lpc].astype("<f4").tobytes()
        data = raw if compression == 0 else _exr_zip(raw)
        chunks.append(struct.pack("<ii", y, len(data)) + data)
    off = len(head) + 8 * len(chunks)
    table = b""
    for c in chunks:
        table += struct.pack("<Q", off)
        off += len(c)
    with open(filename, "wb") as fh:
        fh.write(head + table + b"".join(chunks))


class ImageReader:
    """emf::ImageReader (reference src/utils/ImageReader.cpp:41-116): a Co-Fusion style directory pair, `colour/Color%04d.png`
    and `depth/Depth%04d.exr` (float32 metres); depth > 100 m counts as missing."""

    def __init__(self, colorpath: str, depthpath: str):
        self.colorpath, self.depthpath = colorpath, depthpath
        rgbs = len([f for f in os.listdir(colorpath) if f.endswith(".png")])
        depths = len([f for f in os.listdir(depthpath) if f.endswith(".exr")])
        if rgbs != depths:
            raise RuntimeError("Different number of rgb and depth files!")
        self._n = rgbs
        self.currFrame = 0
        while not all(os.path.exists(f) for f in cofusion_names(colorpath, depthpath, self.currFrame)):   # :72-89
            self.currFrame += 1
            if self.currFrame >= rgbs:
                raise RuntimeError("Could not find starting index!")

    def numFrames(self) -> int:
        return self._n

    def readFrame(self, index: int) -> Tuple[np.ndarray, np.ndarray]:
        from PIL import Image
        cname, dname = cofusion_names(self.colorpath, self.depthpath, index)
        depth = read_exr(dname)
        if depth.dtype != np.float32:
            raise ValueError("Unsupported depth-files")
        rgb = np.asarray(Image.open(cname).convert("RGB"))[..., ::-1].copy()
        return rgb, cofusion_clean(depth)


def write_cofusion_stream(path: str, depths: Sequence[np.ndarray], rgbs: Sequence[np.ndarray] = None, start: int = 0,
                          compression: int = 3) -> Tuple[str, str]:
    """frames -> `<path>/colour/Color%04d.png`, `<path>/depth/Depth%04d.exr`; returns (colorpath, depthpath)"""
    from PIL import Image
    cp, dp = os.path.join(path, "colour"), os.path.join(path, "depth")
    os.makedirs(cp, exist_ok=True); os.makedirs(dp, exist_ok=True)
    for i, d in enumerate(depths):
        cname, dname = cofusion_names(cp, dp, start + i)
        write_exr(dname, d, compression=compression)
        rgb = rgbs[i] if rgbs is not None else np.zeros(d.shape + (3,), dtype=np.uint8)
        Image.fromarray(np.ascontiguousarray(rgb[..., ::-1])).save(cname)
    return cp, dp


# -- src/utils/ImageReader.cpp:41-116: file naming and the depth rule
def cofusion_names(colorpath: str, depthpath: str, index: int) -> Tuple[str, str]:
    return os.path.join(colorpath, f"Color{index:04d}.png"), os.path.join(depthpath, f"Depth{index:04d}.exr")


def cofusion_clean(depth: np.ndarray) -> np.ndarray:
    d = np.asarray(depth, dtype=np.float32).copy()
    d[d > 100] = 0
    return d
