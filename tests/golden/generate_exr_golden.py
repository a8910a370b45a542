"""Writes tests/golden/cv2_depth_*.exr with OpenCV's own EXR codec (cv2.imwrite -- the codec behind the reference's
cv::imread, src/utils/ImageReader.cpp:108) and the pixel values as .npy: the fixtures that pin emfusion_b200.io.read_exr.
Run once where the cv2 wheel has OpenEXR support:  OPENCV_IO_ENABLE_OPENEXR=1 python tests/golden/generate_exr_golden.py"""
import os
os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2
import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(7)
img = (rng.random((24, 40)) * 4.0).astype(np.float32)
img[2, 3] = 250.0      # beyond Co-Fusion's 100 m rule
img[5:9, 10:20] = 0.0
np.save(os.path.join(here, "cv2_depth.npy"), img)
for name, args in (("zip", [cv2.IMWRITE_EXR_COMPRESSION, 3]), ("zips", [cv2.IMWRITE_EXR_COMPRESSION, 2]), ("none", [cv2.IMWRITE_EXR_COMPRESSION, 0]),
                   ("half_zip", [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_HALF, cv2.IMWRITE_EXR_COMPRESSION, 3])):
    assert cv2.imwrite(os.path.join(here, f"cv2_depth_{name}.exr"), img, args)
print("written")
