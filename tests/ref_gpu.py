"""ctypes binding of oracle/_ref/libemf_ref.so -- the reference's own CUDA kernels compiled unchanged
against the type shim (oracle/Makefile).  TEST INFRASTRUCTURE: used by -m gpu tests, the golden-vector
generator and bench.py --impl reference.  Operates on torch CUDA tensors (device pointers only)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libemf_ref.so")
# the same launch structure over THIS REPO's level-1 operators (oracle/Makefile: b200ops) -- bench.py --impl unchanged-caller
B200OPS_PATH = os.path.join(ROOT, "oracle", "_ref", "libemf_ref_b200ops.so")


def use_library(path: str) -> None:
    """load another build of ref_driver.cu instead of the reference kernels (before the first call)"""
    global REF_PATH, _lib
    REF_PATH, _lib = path, None



def available() -> bool:
    return os.path.exists(REF_PATH)


class VolDesc(C.Structure):
    _fields_ = [("tsdf", C.c_void_p), ("weights", C.c_void_p), ("grads", C.c_void_p), ("fg_probs", C.c_void_p),
                ("res", C.c_int * 3), ("voxel", C.c_float), ("trunc", C.c_float), ("id", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(REF_PATH)
        vp, ci, cf = C.c_void_p, C.c_int, C.c_float
        L.emfref_update_tsdf.argtypes = [vp, vp, ci, ci, vp, vp, vp, vp, vp, vp, cf, cf, cf, vp]
        L.emfref_update_gradients.argtypes = [vp, vp, vp, vp]
        L.emfref_raycast.argtypes = [vp, vp, vp, vp, vp, vp, vp, ci, ci, vp, vp, vp, vp, cf, cf, vp]
        L.emfref_get_volume_vals.argtypes = [vp, vp, ci, ci, vp, vp, vp, cf, vp, vp]
        L.emfref_update_fgbg.argtypes = [vp, vp, ci, ci, vp, vp, vp, vp, vp, vp, vp, cf, vp]
        L.emfref_frame_create.argtypes = [ci, ci, ci, C.POINTER(VolDesc)]
        L.emfref_frame_create.restype = vp
        L.emfref_frame_destroy.argtypes = [vp]
        L.emfref_frame_destroy.restype = None
        L.emfref_frame_update_fg_masks.argtypes = [vp]
        L.emfref_frame_assoc.argtypes = [vp, vp, vp, vp, cf, cf, cf]
        L.emfref_frame_raycast.argtypes = [vp, vp, vp, vp, ci, ci]
        L.emfref_frame_integrate.argtypes = [vp, vp, vp, vp, vp, cf, ci]
        for n in ("assoc_ptr", "vol_ray", "vol_vert", "vol_norm", "vol_mask"):
            f = getattr(L, "emfref_frame_" + n)
            f.argtypes = [vp, ci]
            f.restype = vp
        for n in ("ray", "vert", "norm", "seg"):
            f = getattr(L, "emfref_frame_" + n)
            f.argtypes = [vp]
            f.restype = vp
        L.emfref_frame_visible.argtypes = [vp, ci]
        L.emfref_frame_set_visible.argtypes = [vp, ci, ci]
        L.emfref_frame_set_visible.restype = None
        L.emfref_frame_fill_assoc.argtypes = [vp, ci, cf]
        L.emfref_memcpy_d2d.argtypes = [vp, vp, C.c_size_t]
        L.emfref_compute_points.argtypes = [vp, vp, ci, ci, vp]
        L.emfref_copy_values.argtypes = [vp, vp, ci, vp, vp, vp]
        if hasattr(L, "emfref_marching_cubes"):
            L.emfref_marching_cubes.argtypes = [vp, vp, vp, vp, cf, vp]
            L.emfref_marching_cubes.restype = vp
            L.emfref_mesh_fetch.argtypes = [vp, vp, vp, vp]
        if not hasattr(L, "emfref_tracker_create"):      # (a build without the tracker operators)
            _lib = L
            return _lib
        L.emfref_tracker_create.argtypes = [ci, ci]
        L.emfref_tracker_create.restype = vp
        L.emfref_tracker_destroy.argtypes = [vp]
        L.emfref_tracker_destroy.restype = None
        L.emfref_tracker_ptr.argtypes = [vp, ci]
        L.emfref_tracker_ptr.restype = vp
        L.emfref_tracker_linearise.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, cf, cf, cf, vp, vp, vp]
        L.emfref_tracker_error.argtypes = [vp, vp, vp, vp, vp, vp, cf, vp]
        _lib = L
    return _lib


def _h(a, dtype=np.float32):
    """host array -> (keepalive, pointer)"""
    arr = np.ascontiguousarray(np.asarray(a, dtype=dtype).reshape(-1))
    return arr, arr.ctypes.data


def _s():
    return torch.cuda.current_stream().cuda_stream


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError(f"reference {what} failed rc={rc}")


def compute_points(depth, points, K):
    h, w = depth.shape
    kk, pk = _h(K)
    _chk(lib().emfref_compute_points(depth.data_ptr(), points.data_ptr(), w, h, pk), "computePoints")


def update_tsdf(depth, assoc, tsdf, weights, R, t, K, res, voxel, trunc, maxw):
    h, w = depth.shape
    kr, pr = _h(R); kt, pt = _h(t); kk, pk = _h(K); ks, ps = _h(res, np.int32)
    _chk(lib().emfref_update_tsdf(depth.data_ptr(), assoc.data_ptr(), w, h, tsdf.data_ptr(), weights.data_ptr(), pr, pt,
                                  pk, ps, voxel, trunc, maxw, _s()), "updateTSDF")


def update_gradients(tsdf, grads, res):
    ks, ps = _h(res, np.int32)
    _chk(lib().emfref_update_gradients(tsdf.data_ptr(), grads.data_ptr(), ps, _s()), "updateGradients")


def raycast(tsdf, grads, weights, ray, vert, norm, mask, R, t, K, res, voxel, trunc):
    h, w = ray.shape
    kr, pr = _h(R); kt, pt = _h(t); kk, pk = _h(K); ks, ps = _h(res, np.int32)
    _chk(lib().emfref_raycast(tsdf.data_ptr(), grads.data_ptr(), weights.data_ptr(), ray.data_ptr(), vert.data_ptr(),
                              norm.data_ptr(), mask.data_ptr(), w, h, pr, pt, pk, ps, voxel, trunc, _s()), "raycastTSDF")


def get_volume_vals(vol, points, R, t, res, voxel, vals):
    h, w = vals.shape
    kr, pr = _h(R); kt, pt = _h(t); ks, ps = _h(res, np.int32)
    _chk(lib().emfref_get_volume_vals(vol.data_ptr(), points.data_ptr(), w, h, pr, pt, ps, voxel, vals.data_ptr(), _s()),
         "getVolumeVals")


def update_fgbg(mask, occluded, tsdf, weights, fgbg, R, t, K, res, voxel):
    h, w = mask.shape
    kr, pr = _h(R); kt, pt = _h(t); kk, pk = _h(K); ks, ps = _h(res, np.int32)
    _chk(lib().emfref_update_fgbg(mask.data_ptr(), occluded.data_ptr(), w, h, tsdf.data_ptr(), weights.data_ptr(),
                                  fgbg.data_ptr(), pr, pt, pk, ps, voxel, _s()), "updateFgBgProbs")


class RefFrame:
    """The reference's per-frame hot path (launch structure of src/core/EMFusion.cpp) over caller volumes.
    vols: list of dicts(tsdf, weights, grads, fg_probs|None, res, voxel, trunc, id); [0] = background."""

    def __init__(self, w, h, vols):
        self.w, self.h, self.n_vol = w, h, len(vols)
        arr = (VolDesc * len(vols))()
        for i, v in enumerate(vols):
            arr[i].tsdf = v["tsdf"].data_ptr(); arr[i].weights = v["weights"].data_ptr()
            arr[i].grads = v["grads"].data_ptr()
            arr[i].fg_probs = v["fg_probs"].data_ptr() if v.get("fg_probs") is not None else None
            arr[i].res[:] = [int(r) for r in v["res"]]
            arr[i].voxel = v["voxel"]; arr[i].trunc = v["trunc"]; arr[i].id = v["id"]
        self._keep = vols
        torch.cuda.synchronize()
        self.hd = lib().emfref_frame_create(w, h, len(vols), arr)
        self.update_fg_masks()

    def close(self):
        if self.hd:
            torch.cuda.synchronize()
            lib().emfref_frame_destroy(self.hd)
            self.hd = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update_fg_masks(self):
        torch.cuda.synchronize()
        _chk(lib().emfref_frame_update_fg_masks(self.hd), "fg masks")

    def assoc(self, points, R_co, t_co, sigma=0.02, alpha=0.8, uni=1.0):
        torch.cuda.synchronize()
        kr, pr = _h(R_co); kt, pt = _h(t_co)
        _chk(lib().emfref_frame_assoc(self.hd, points.data_ptr(), pr, pt, sigma, alpha, uni), "assoc")

    def raycast(self, R_co, t_co, K, boundary=20, visibility_thresh=1600):
        kr, pr = _h(R_co); kt, pt = _h(t_co); kk, pk = _h(K)
        _chk(lib().emfref_frame_raycast(self.hd, pr, pt, pk, boundary, visibility_thresh), "raycast")

    def integrate(self, depth, R_oc, t_oc, K, maxw=64.0, use_vis=True):
        kr, pr = _h(R_oc); kt, pt = _h(t_oc); kk, pk = _h(K)
        _chk(lib().emfref_frame_integrate(self.hd, depth.data_ptr(), pr, pt, pk, maxw, 1 if use_vis else 0), "integrate")

    def assoc_image(self, i):
        return _from_ptr(lib().emfref_frame_assoc_ptr(self.hd, i), (self.h, self.w), torch.float32)

    def vol_outputs(self, i):
        L = lib()
        return dict(ray=_from_ptr(L.emfref_frame_vol_ray(self.hd, i), (self.h, self.w), torch.float32),
                    vert=_from_ptr(L.emfref_frame_vol_vert(self.hd, i), (self.h, self.w, 3), torch.float32),
                    norm=_from_ptr(L.emfref_frame_vol_norm(self.hd, i), (self.h, self.w, 3), torch.float32),
                    mask=_from_ptr(L.emfref_frame_vol_mask(self.hd, i), (self.h, self.w), torch.uint8))

    def composite(self):
        L = lib()
        return dict(ray=_from_ptr(L.emfref_frame_ray(self.hd), (self.h, self.w), torch.float32),
                    vert=_from_ptr(L.emfref_frame_vert(self.hd), (self.h, self.w, 3), torch.float32),
                    norm=_from_ptr(L.emfref_frame_norm(self.hd), (self.h, self.w, 3), torch.float32),
                    seg=_from_ptr(L.emfref_frame_seg(self.hd), (self.h, self.w), torch.uint8))

    def download_composite(self, seg_host, ray_host):
        """blocking D2H copies of modelSegmentation / raylengths into (pinned) host tensors"""
        L = lib()
        _chk(L.emfref_memcpy_d2d(seg_host.data_ptr(), L.emfref_frame_seg(self.hd), seg_host.numel()), "memcpy")
        _chk(L.emfref_memcpy_d2d(ray_host.data_ptr(), L.emfref_frame_ray(self.hd), ray_host.numel() * 4), "memcpy")

    def visible(self, i):
        return bool(lib().emfref_frame_visible(self.hd, i))

    def set_visible(self, i, v):
        lib().emfref_frame_set_visible(self.hd, i, 1 if v else 0)

    def fill_assoc(self, i, v):
        _chk(lib().emfref_frame_fill_assoc(self.hd, i, v), "fill_assoc")


def _from_ptr(ptr, shape, dtype):
    """Copy device memory at ptr into a fresh torch tensor (D2D)."""
    out = torch.empty(shape, dtype=dtype, device="cuda")
    torch.cuda.synchronize()
    _chk(lib().emfref_memcpy_d2d(out.data_ptr(), ptr, out.numel() * out.element_size()), "memcpy")
    return out


def copy_values(src, dst, channels, offset, src_res, dst_res):
    ko, po = _h(offset, np.int32); ks, ps = _h(src_res, np.int32); kd, pd = _h(dst_res, np.int32)
    _chk(lib().emfref_copy_values(src.data_ptr(), dst.data_ptr(), channels, po, ps, pd), "copyValues")


class RefTracker:
    """the tracking members of one emf::TSDF + the reference's launch chain of one iteration (oracle/ref_driver.cu section 4)"""
    GRADS, TSDF_VALS, INT_WEIGHTS, TRACK_WEIGHTS, AS, BS, A_GPU, B_GPU = range(8)

    def __init__(self, w, h):
        self.w, self.h = w, h
        self.handle = lib().emfref_tracker_create(w, h)

    def close(self):
        if self.handle:
            lib().emfref_tracker_destroy(self.handle)
            self.handle = None

    def linearise(self, tsdf, grads_vol, weights, points, assoc, R, t, res, voxel, huber=0.2, maxw=64.0):
        """-> (A (6,6), b (6,), err) as the reference downloads them"""
        kr, pr = _h(R); kt, pt = _h(t); ks, ps = _h(res, np.int32)
        A = np.zeros(36, dtype=np.float32); b = np.zeros(6, dtype=np.float32); e = np.zeros(1, dtype=np.float32)
        _chk(lib().emfref_tracker_linearise(self.handle, tsdf.data_ptr(), grads_vol.data_ptr(), weights.data_ptr(),
                                            points.data_ptr(), assoc.data_ptr(), pr, pt, ps, voxel, huber, maxw,
                                            A.ctypes.data, b.ctypes.data, e.ctypes.data), "tracker linearise")
        return A.reshape(6, 6), b, float(e[0])

    def error(self, tsdf, points, R, t, res, voxel):
        kr, pr = _h(R); kt, pt = _h(t); ks, ps = _h(res, np.int32)
        e = np.zeros(1, dtype=np.float32)
        _chk(lib().emfref_tracker_error(self.handle, tsdf.data_ptr(), points.data_ptr(), pr, pt, ps, voxel, e.ctypes.data),
             "tracker error")
        return float(e[0])

    def image(self, what, shape, dtype=torch.float32):
        """copy of one of the tracker's device buffers as a torch tensor"""
        out = torch.empty(shape, dtype=dtype, device="cuda")
        _chk(lib().emfref_memcpy_d2d(out.data_ptr(), lib().emfref_tracker_ptr(self.handle, what), out.numel() * out.element_size()),
             "memcpy")
        return out


MESH_PATH = os.path.join(ROOT, "oracle", "_ref", "libemf_ref_mesh.so")
_mesh = None


def mesh_lib():
    """the reference's TSDF.cu compiled with a 64-register cap (oracle/Makefile says why): its marching-cubes launcher only"""
    global _mesh
    if _mesh is None:
        _mesh = C.CDLL(MESH_PATH)
        _mesh.emfref_marching_cubes.restype = C.c_void_p
        _mesh.emfref_marching_cubes.argtypes = [C.c_void_p] * 4 + [C.c_float, C.c_void_p]
        _mesh.emfref_mesh_fetch.restype = C.c_int
        _mesh.emfref_mesh_fetch.argtypes = [C.c_void_p] * 4
    return _mesh


def marching_cubes(tsdf, grads, mask, res, voxel):
    """the reference's emf::cuda::TSDF::marchingCubes -> (vertices (n, 3), normals (n, 3), triangles (m, 4) int32) CUDA tensors"""
    counts = np.zeros(2, dtype=np.int32)
    r = np.asarray(res, dtype=np.int32)
    h = mesh_lib().emfref_marching_cubes(tsdf.data_ptr(), grads.data_ptr(), mask.data_ptr(), r.ctypes.data, float(voxel), counts.ctypes.data)
    nv, nt = int(counts[0]), int(counts[1])
    v = torch.empty((nv, 3), dtype=torch.float32, device=tsdf.device)
    n = torch.empty((nv, 3), dtype=torch.float32, device=tsdf.device)
    t = torch.empty((nt // 4, 4), dtype=torch.int32, device=tsdf.device)
    _chk(mesh_lib().emfref_mesh_fetch(h, v.data_ptr(), n.data_ptr(), t.data_ptr()), "mesh_fetch")
    return v, n, t
