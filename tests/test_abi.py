"""CPU: the C-ABI shared library loads, exports exactly the symbols include/emf_b200.h declares, rejects bad
arguments before touching the device, and the product package has one compute path (no oracle, no fallback).
No compute call is made here -- there is no GPU in this container."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from emfusion_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "emf_b200.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        build.build_product()
    return _lib.lib()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"EMF_API\s+(?:const\s+)?\w+\s*\*?\s*(emf_\w+)\s*\(", src)))


def test_header_declares_the_bound_symbols():
    assert declared_symbols() == _lib.EXPORTED


def test_library_exports_every_declared_symbol(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, f"not exported: {missing}"
    # nothing but the C ABI leaks out of the library (hidden visibility for everything else)
    leaked = [s for s in exported if not s.startswith("emf_") and not s.startswith("_")]
    assert not leaked, leaked


def test_every_header_citation_points_into_the_reference_tree():
    src = open(HEADER).read()
    cites = re.findall(r"((?:src|include)/[\w/]+\.(?:cpp|cu|cuh|h)):(\d+)", src)
    assert len(cites) >= 15
    ref = "/root/reference"
    if os.path.isdir(ref):
        for path, line in cites:
            p = os.path.join(ref, path)
            assert os.path.exists(p), p
            assert int(line) <= sum(1 for _ in open(p, errors="replace")), (path, line)


def test_version(lib):
    assert lib.emf_version().decode().startswith("emf_b200 ") and "sm_100a" in lib.emf_version().decode()


def test_only_sm_100a_code_is_embedded():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_invalid_arguments_are_rejected_without_a_device(lib):
    img = _lib.Image(None, 0, 0, 0)
    K = (C.c_float * 9)(525, 0, 319.5, 0, 525, 239.5, 0, 0, 1)
    pose = _lib.Pose()
    res = (C.c_int * 3)(64, 64, 64)
    assert lib.emf_compute_points(img, img, K, None) == _lib.EMF_ERR_INVALID
    assert lib.emf_update_tsdf(img, img, None, None, pose, K, res, 0.01, 0.1, 64.0, None) == _lib.EMF_ERR_INVALID
    assert lib.emf_compute_tsdf_grads(None, None, res, None) == _lib.EMF_ERR_INVALID
    assert lib.emf_get_volume_vals(None, img, pose, res, 0.01, img, None) == _lib.EMF_ERR_INVALID
    assert lib.emf_compute_fg_probs(None, 10, None, None, None) == _lib.EMF_ERR_INVALID
    assert lib.emf_assoc_weights(0, None, None, img, None, None, 0, None, None) == _lib.EMF_ERR_INVALID
    assert lib.emf_integrate_volumes(0, None, None, K, img, None, 64.0, None) == _lib.EMF_ERR_INVALID
    assert lib.emf_raycast_volumes(0, None, None, K, None, None, None, None, None, None, None) == _lib.EMF_ERR_INVALID
    vols = (_lib.Volume * 1)()
    poses = (_lib.Pose * 1)()
    imgs = (_lib.Image * 1)()
    prm = _lib.TsdfParams(64, 0.02, 0.8, 1.0)
    too_many = _lib.EMF_MAX_VOLUMES + 1
    pts = _lib.Image(0x1000, 640 * 12, 640, 480)   # plausible header; never dereferenced on the host
    assert lib.emf_assoc_weights(too_many, vols, poses, pts, prm, imgs, 0, None, None) == _lib.EMF_ERR_UNSUPPORTED
    assert lib.emf_assoc_weights(1, vols, poses, pts, prm, imgs, 7, None, None) == _lib.EMF_ERR_INVALID
    # a volume with a degenerate resolution
    bad = (C.c_int * 3)(1, 64, 64)
    assert lib.emf_compute_tsdf_grads(0x1000, 0x2000, bad, None) == _lib.EMF_ERR_INVALID
    with pytest.raises(_lib.EmfError):
        _lib.check(_lib.EMF_ERR_INVALID, "x")


def test_screen_rect_host_helper(lib):
    K = (C.c_float * 9)(525, 0, 319.5, 0, 525, 239.5, 0, 0, 1)
    res = (C.c_int * 3)(64, 64, 64)
    out = (C.c_int * 4)()
    pose = _lib.Pose()
    pose.R[:] = [1, 0, 0, 0, 1, 0, 0, 0, 1]
    # camera -> volume: volume centre 2 m in front of the camera
    pose.t[:] = [0, 0, -2.0]
    assert lib.emf_volume_screen_rect(res, 0.01, pose, K, 640, 480, out) == _lib.EMF_OK
    x0, y0, x1, y1 = list(out)
    b = 31 * 0.01
    near = 2.0 - b
    exp_half = 525 * b / near
    assert x0 <= 319.5 - exp_half and x1 >= 319.5 + exp_half and y0 <= 239.5 - exp_half and y1 >= 239.5 + exp_half
    assert x1 - x0 <= 2 * exp_half + 8 and 0 <= x0 < x1 <= 640 and 0 <= y0 < y1 <= 480
    # camera inside the box: full frame
    pose.t[:] = [0, 0, 0]
    assert lib.emf_volume_screen_rect(res, 0.01, pose, K, 640, 480, out) == _lib.EMF_OK
    assert list(out) == [0, 0, 640, 480]
    # box entirely off-screen to the right: empty rectangle, clamped
    pose.t[:] = [-50.0, 0, -2.0]
    assert lib.emf_volume_screen_rect(res, 0.01, pose, K, 640, 480, out) == _lib.EMF_OK
    assert out[2] - out[0] == 0 or out[0] >= 640 - 1
    assert lib.emf_volume_screen_rect(None, 0.01, pose, K, 640, 480, out) == _lib.EMF_ERR_INVALID


def test_product_never_touches_the_oracle_and_has_no_fallback(monkeypatch, tmp_path):
    pkg = os.path.join(ROOT, "emfusion_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle_c" not in txt and "emf_oracle" not in txt and "libemf_ref" not in txt, f
                assert "import triton" not in txt and "torch.compile" not in txt, f
    # a missing library is a loud error, not a fallback
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.EmfError, match="no CPU fallback"):
        _lib.lib()


def test_ops_refuse_cpu_tensors():
    import torch
    from emfusion_b200 import ops
    with pytest.raises(_lib.EmfError):
        ops.image(torch.zeros((4, 4)))
    with pytest.raises(_lib.EmfError):
        ops.volume(torch.zeros(8), torch.zeros(8), (2, 2, 2), 0.1, 1.0)


def test_struct_layouts_match_the_header():
    """sizeof/offsetof of the POD descriptors as a C compiler sees them == the ctypes mirror."""
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "emf_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu\n", sizeof(emf_image), sizeof(emf_pose), sizeof(emf_tsdf_params), sizeof(emf_volume));
  printf("%zu %zu %zu %zu %zu %zu %zu\n", offsetof(emf_volume, weights), offsetof(emf_volume, grads),
         offsetof(emf_volume, fg_probs), offsetof(emf_volume, res), offsetof(emf_volume, voxel_size),
         offsetof(emf_volume, truncdist), offsetof(emf_volume, id));
  return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.run(["/usr/bin/gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        a, b = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    assert [int(x) for x in a.split()] == [C.sizeof(_lib.Image), C.sizeof(_lib.Pose), C.sizeof(_lib.TsdfParams),
                                           C.sizeof(_lib.Volume)]
    V = _lib.Volume
    assert [int(x) for x in b.split()] == [V.weights.offset, V.grads.offset, V.fg_probs.offset, V.res.offset,
                                           V.voxel_size.offset, V.truncdist.offset, V.id.offset]


def test_new_entry_points_reject_invalid_arguments_without_a_device(lib):
    """tracker, resize, pre-filter and exchange entry points: argument errors are detected on the host"""
    img = _lib.Image(None, 0, 0, 0)
    K = (C.c_float * 9)(525, 0, 319.5, 0, 525, 239.5, 0, 0, 1)
    res = (C.c_int * 3)(64, 64, 64)
    off = (C.c_int * 3)(0, 0, 0)
    vols, poses, imgs = (_lib.Volume * 1)(), (_lib.Pose * 1)(), (_lib.Image * 1)()
    modes = (C.c_int * 1)(1)
    pts = _lib.Image(0x1000, 640 * 12, 640, 480)
    assert lib.emf_track_workspace_bytes(0) == 0 and lib.emf_track_workspace_bytes(_lib.EMF_MAX_VOLUMES + 1) == 0
    assert lib.emf_track_workspace_bytes(2) > lib.emf_track_workspace_bytes(1) > 0
    assert lib.emf_track_linearise(0, vols, poses, modes, pts, K, imgs, 0.2, 64.0, imgs, None, None, None, 0x1000, 0x1000, 1 << 20,
                                   None) == _lib.EMF_ERR_INVALID
    assert lib.emf_track_linearise(1, vols, poses, modes, pts, K, imgs, 0.2, 64.0, imgs, None, None, None, 0x1000, 0x1000, 8,
                                   None) == _lib.EMF_ERR_INVALID          # workspace too small
    assert lib.emf_track_linearise(_lib.EMF_MAX_VOLUMES + 1, vols, poses, modes, pts, K, imgs, 0.2, 64.0, imgs, None, None, None,
                                   0x1000, 0x1000, 1 << 30, None) == _lib.EMF_ERR_UNSUPPORTED
    bad_mode = (C.c_int * 1)(5)
    assert lib.emf_track_linearise(1, vols, poses, bad_mode, pts, K, imgs, 0.2, 64.0, imgs, None, None, None, 0x1000, 0x1000,
                                   1 << 20, None) == _lib.EMF_ERR_INVALID
    lm = _lib.TrackLMParams(1e3, 1e-8, 1e-8, 2.0, 0.2, 64.0)
    assert lib.emf_track_iterate(1, vols, None, poses, pts, K, imgs, lm, imgs, 0x1000, 0x1000, 1 << 20, 1, None) == _lib.EMF_ERR_INVALID
    assert lib.emf_copy_values(0x1000, 0x2000, 4, off, res, res, None) == _lib.EMF_ERR_UNSUPPORTED
    assert lib.emf_copy_values(0x1000, 0x1000, 1, off, res, res, None) == _lib.EMF_ERR_INVALID    # in place
    assert lib.emf_resize_volume(0x1000, 0x2000, 0x3000, res, 0x4000, 0x5000, None, res, off, None) == _lib.EMF_ERR_INVALID
    assert lib.emf_preprocess_depth(img, img, None, K, 7, 0.04, 4.5, None) == _lib.EMF_ERR_INVALID
    d = _lib.Image(0x1000, 640 * 4, 640, 480)
    assert lib.emf_preprocess_depth(d, d, None, K, 7, 0.04, 4.5, None) == _lib.EMF_ERR_INVALID       # in place
    d2 = _lib.Image(0x9000, 640 * 4, 640, 480)
    assert lib.emf_preprocess_depth(d, d2, None, K, 31, 0.04, 4.5, None) == _lib.EMF_ERR_UNSUPPORTED  # window too large
    assert lib.emf_xchg_signal(0, None, 1, None) == _lib.EMF_ERR_INVALID
    assert lib.emf_xchg_wait(None, 1, 1, None, 1.0, None) == _lib.EMF_ERR_INVALID
    assert lib.emf_xchg_sum_images(0, None, img, None) == _lib.EMF_ERR_INVALID
    assert lib.emf_xchg_open(None, None) == _lib.EMF_ERR_INVALID


def test_track_state_layout_matches_the_header():
    import numpy as np
    from emfusion_b200 import ops
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "emf_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(emf_track_state), offsetof(emf_track_state, t), offsetof(emf_track_state, R_old),
         offsetof(emf_track_state, mu), offsetof(emf_track_state, A), offsetof(emf_track_state, x), offsetof(emf_track_state, err),
         offsetof(emf_track_state, converged));
  printf("%zu\n", sizeof(emf_track_lm_params));
  return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.run(["/usr/bin/gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        a, b = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    dt = ops.TRACK_STATE_DTYPE
    assert [int(x) for x in a.split()] == [dt.itemsize, dt.fields["t"][1], dt.fields["R_old"][1], dt.fields["mu"][1], dt.fields["A"][1],
                                           dt.fields["x"][1], dt.fields["err"][1], dt.fields["converged"][1]]
    assert int(b) == C.sizeof(_lib.TrackLMParams)
