"""CPU: the class-level drop-in include/emf_b200_cv.hpp -- emf::TSDF / emf::ObjTSDF with the reference's own signatures over
the C ABI -- compiles next to the reference's class headers, and every public method src/core/EMFusion.cpp calls has the same
member-function type as the reference's declaration (static_asserts in tests/csrc/cv_adapter_check.cpp).  OpenCV-with-CUDA,
Eigen and Sophus are not installable in this image: the type stand-in of oracle/shim_full is used for both sides."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_INC = "/root/reference/include"


@pytest.mark.skipif(not os.path.isdir(REF_INC), reason="reference headers not present on this box")
@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_cv_adapter_matches_reference_class_surface(tmp_path):
    obj = tmp_path / "cv_adapter_check.o"
    cmd = ["g++", "-std=c++17", "-c", "-w", "-I", os.path.join(ROOT, "oracle", "shim_full"), "-I", REF_INC,
           "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
           os.path.join(ROOT, "tests", "csrc", "cv_adapter_check.cpp"), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    syms = subprocess.run(["nm", "-C", "-u", str(obj)], capture_output=True, text=True).stdout
    # the adapter's inline bodies reach the product only through the C ABI
    for c_abi in ("emf_update_tsdf", "emf_raycast_tsdf", "emf_compute_association", "emf_track_iterate", "emf_update_fgbg_probs"):
        assert c_abi in syms, f"{c_abi} not referenced by the adapter"
