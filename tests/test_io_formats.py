"""CPU: the on-disk formats either side of the path (emfusion_b200/io.py): a synthetic stream round-trips through the TUM
RGB-D container exactly as emf::TUMRGBDReader would read it, and volume dumps round-trip through writeVolume's layout."""
import os
import struct

import numpy as np
import pytest

from emfusion_b200 import io as emfio
from emfusion_b200.synth import Scene


def test_tum_round_trip(tmp_path):
    scene = Scene(n_objects=2, width=160, height=120, seed=1)
    frames = [scene.render(f) for f in range(4)]
    rgbs = [np.dstack([(i * 40).astype(np.uint8)] * 3) for _, i in frames]
    seen = emfio.write_tum_stream(str(tmp_path), [d for d, _ in frames], rgbs, fps=30.0)
    rd = emfio.TUMRGBDReader(str(tmp_path))
    assert rd.numFrames() == 4 and all(n.startswith("rgb/") for n in rd.rgbFileNames)
    assert all(n.startswith("depth/") for n in rd.depthFileNames)
    assert abs(rd.frameRate - 4 / (3 / 30.0)) < 1e-2 and rd.minBufferSize == round(rd.frameRate)
    for f in range(4):
        rgb, depth = rd.readFrame(f)
        assert depth.dtype == np.float32 and depth.shape == (120, 160)
        assert np.array_equal(depth, seen[f])
        assert np.abs(depth - frames[f][0]).max() <= 0.5 / 5000 + 1e-6      # 0.2 mm quantisation
        assert np.array_equal(rgb, rgbs[f])
    # metres = uint16 * (1/5000.f) in float32, as convertTo computes it
    assert seen[0].max() > 1.0 and np.float32(5000) * emfio.TUM_DEPTH_SCALE == np.float32(1.0)


def test_tum_depth_first_associations(tmp_path):
    os.makedirs(tmp_path / "rgb"); os.makedirs(tmp_path / "depth")
    (tmp_path / "associations.txt").write_text("1.0 depth/a.png 1.01 rgb/a.png\n1.5 depth/b.png 1.51 rgb/b.png\n\nbroken line\n")
    rd = emfio.TUMRGBDReader(str(tmp_path))
    assert rd.rgbFileNames == ["rgb/a.png", "rgb/b.png"] and rd.depthFileNames == ["depth/a.png", "depth/b.png"]
    assert rd.frameRate == pytest.approx(2 / 0.5)
    with pytest.raises(RuntimeError):
        emfio.TUMRGBDReader(str(tmp_path / "nowhere"))


def test_volume_dump_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    res = (6, 5, 4)
    for ch in (1, 2, 3):
        shape = (res[1] * res[2], res[0]) if ch == 1 else (res[1] * res[2], res[0], ch)
        vol = rng.standard_normal(shape).astype(np.float32)
        fn = str(tmp_path / f"v{ch}.bin")
        emfio.writeVolume(fn, vol, res, 0.0125)
        raw = open(fn, "rb").read()
        assert struct.unpack("<3i", raw[:12]) == res and struct.unpack("<Q", raw[12:20])[0] == 4 * ch
        assert len(raw) == 24 + vol.nbytes
        back, r2, vs = emfio.readVolume(fn)
        assert r2 == res and vs == np.float32(0.0125) and np.array_equal(back, vol)


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_exr_reader_against_opencv_fixtures():
    """files written by OpenCV's EXR codec (tests/golden/generate_exr_golden.py; the codec behind the reference's cv::imread)
    decode to the exact pixel values: FLOAT and HALF channels, compression NONE / ZIPS / ZIP"""
    want = np.load(os.path.join(GOLDEN, "cv2_depth.npy"))
    for name in ("zip", "zips", "none"):
        got = emfio.read_exr(os.path.join(GOLDEN, f"cv2_depth_{name}.exr"))
        assert got.dtype == np.float32 and np.array_equal(got, want), name
    half = emfio.read_exr(os.path.join(GOLDEN, "cv2_depth_half_zip.exr"))
    assert np.array_equal(half, want.astype(np.float16).astype(np.float32))


def test_exr_round_trip_and_opencv_reads_ours(tmp_path):
    rng = np.random.default_rng(3)
    img = (rng.random((33, 47)) * 6).astype(np.float32)
    for comp in (0, 2, 3):
        fn = str(tmp_path / f"d{comp}.exr")
        emfio.write_exr(fn, img, compression=comp)
        assert np.array_equal(emfio.read_exr(fn), img)
    with pytest.raises(ValueError):
        emfio.read_exr(os.path.join(GOLDEN, "cv2_depth.npy"))
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    try:
        import cv2
        back = cv2.imread(str(tmp_path / "d3.exr"), cv2.IMREAD_UNCHANGED)
    except Exception:
        back = None
    if back is not None:       # the reference's own reader sees a CV_32FC1 image with the same values
        assert back.dtype == np.float32 and back.shape == img.shape and np.array_equal(back, img)


def test_cofusion_stream_round_trip(tmp_path):
    """config 3's container: colour/Color%04d.png + depth/Depth%04d.exr, starting index found like the reference does"""
    scene = Scene(n_objects=2, width=160, height=120, seed=1)
    frames = [scene.render(f) for f in range(3)]
    depths = [d.copy() for d, _ in frames]
    depths[1][0, 0] = 500.0                                   # > 100 m: treated as missing (ImageReader.cpp:113)
    cp, dp = emfio.write_cofusion_stream(str(tmp_path), depths, start=2)
    rd = emfio.ImageReader(cp, dp)
    assert rd.numFrames() == 3 and rd.currFrame == 2
    for f in range(3):
        rgb, depth = rd.readFrame(2 + f)
        want = depths[f].copy(); want[want > 100] = 0
        assert depth.dtype == np.float32 and np.array_equal(depth, want)
    os.remove(os.path.join(dp, "Depth0004.exr"))
    with pytest.raises(RuntimeError):
        emfio.ImageReader(cp, dp)


def test_cofusion_rules():
    c, d = emfio.cofusion_names("/a/colour", "/a/depth", 7)
    assert c.endswith("Color0007.png") and d.endswith("Depth0007.exr")
    assert np.array_equal(emfio.cofusion_clean(np.array([1.0, 101.0, 100.0], np.float32)), np.array([1.0, 0.0, 100.0], np.float32))
