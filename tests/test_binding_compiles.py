"""CPU: the reference-side binding (bindings/emf_b200_opencv_binding.cpp) compiles against
the reference's own operator headers -- i.e. it defines the six level-1 operators with exactly the reference's
signatures -- using the type-only OpenCV stand-in of oracle/shim (real OpenCV is not installable in this image)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_INC = "/root/reference/include"


@pytest.mark.skipif(not os.path.isdir(REF_INC), reason="reference headers not present on this box")
@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_opencv_binding_compiles_against_reference_headers(tmp_path):
    obj = tmp_path / "binding.o"
    cmd = ["g++", "-std=c++17", "-c", "-w", "-DEMF_B200_BINDING_NO_FRAME_OPS", "-I", os.path.join(ROOT, "oracle", "shim"),
           "-I", REF_INC, "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
           os.path.join(ROOT, "bindings", "emf_b200_opencv_binding.cpp"), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    syms = subprocess.run(["nm", "-C", str(obj)], capture_output=True, text=True).stdout
    for name in ("emf::cuda::TSDF::updateTSDF(", "emf::cuda::TSDF::computeTSDFGrads(", "emf::cuda::TSDF::raycastTSDF(",
                 "emf::cuda::TSDF::getVolumeVals(", "emf::cuda::TSDF::copyValues(", "emf::cuda::ObjTSDF::updateFgBgProbs("):
        assert any(name in ln and " T " in ln for ln in syms.splitlines()), f"{name} not defined by the binding"
    for c_abi in ("emf_update_tsdf", "emf_compute_tsdf_grads", "emf_raycast_tsdf", "emf_get_volume_vals",
                  "emf_update_fgbg_probs", "emf_copy_values"):
        assert any(ln.strip().endswith("U " + c_abi) for ln in syms.splitlines()), f"{c_abi} not referenced"
