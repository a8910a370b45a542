"""ctypes/numpy binding of the C oracle (oracle/emf_oracle.c).  TEST INFRASTRUCTURE: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")

_f = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u8 = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64 = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


class Oracle:
    def __init__(self, path: str):
        L = C.CDLL(path)
        self.L = L
        self.path = path
        L.emfo_compute_points.argtypes = [_f, C.c_int, C.c_int, _f, _f]
        L.emfo_update_tsdf.argtypes = [_f, _f, C.c_int, C.c_int, _f, _f, _f, _f, _f, _i32, C.c_float, C.c_float,
                                       C.c_float, C.c_void_p]
        L.emfo_compute_grads.argtypes = [_f, _f, _i32]
        L.emfo_raycast.argtypes = [_f, _f, _f, _f, _f, _f, _u8, C.c_int, C.c_int, _f, _f, _f, _i32, C.c_float,
                                   C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        L.emfo_get_volume_vals.argtypes = [_f, _f, C.c_int, C.c_int, _f, _f, _i32, C.c_float, _f, C.c_void_p]
        L.emfo_assoc_volume.argtypes = [_f, C.c_void_p, _f, C.c_int, C.c_int, _f, _f, _i32, C.c_float, C.c_float,
                                        C.c_float, C.c_float, C.c_float, _f, C.c_void_p, _f]
        L.emfo_normalise.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.emfo_composite.argtypes = [C.c_int, _i32] + [C.POINTER(C.c_void_p)] * 4 + [_f, _f, _f, _u8, C.c_int, C.c_int,
                                                                                   C.c_int, _f, _f, _f, _u8, _i64]
        L.emfo_update_fgbg.argtypes = [_u8, _u8, C.c_int, C.c_int, _f, _f, _f, _f, _f, _f, _i32, C.c_float]
        L.emfo_compute_fg_probs.argtypes = [_f, C.c_int64, _f, _u8]
        L.emfo_raycast_weights.argtypes = [_f, _u8, C.c_int64, _f]
        _d = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
        L.emfo_track_linearise.argtypes = [_f, _f, _f, _f, _f, C.c_int, C.c_int, _f, _f, _i32, C.c_float, C.c_float,
                                           C.c_float, _f, _f, _f, _f, _d, _d, _d, _d]
        L.emfo_track_linearise.restype = None
        L.emfo_preprocess_depth.argtypes = [_f, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _f]
        L.emfo_preprocess_depth.restype = None
        L.emfo_uses_fma.restype = C.c_int
        for n in ("emfo_compute_points", "emfo_update_tsdf", "emfo_compute_grads", "emfo_raycast",
                  "emfo_get_volume_vals", "emfo_assoc_volume", "emfo_normalise", "emfo_composite",
                  "emfo_update_fgbg", "emfo_compute_fg_probs", "emfo_raycast_weights"):
            getattr(L, n).restype = None

    # ---- helpers: R (9,), t (3,), K (9,) float32; res int32 (3,)
    @staticmethod
    def _p(a):
        return np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))

    @staticmethod
    def _r(res):
        return np.ascontiguousarray(np.asarray(res, dtype=np.int32))

    def compute_points(self, depth, K):
        h, w = depth.shape
        out = np.zeros((h, w, 3), dtype=np.float32)
        self.L.emfo_compute_points(np.ascontiguousarray(depth), w, h, self._p(K), out)
        return out

    def update_tsdf(self, depth, assoc, tsdf, weights, R, t, K, res, voxel, trunc, maxw, counts=False):
        h, w = depth.shape
        c = np.zeros(6, dtype=np.int64)
        self.L.emfo_update_tsdf(np.ascontiguousarray(depth), np.ascontiguousarray(assoc), w, h, tsdf, weights,
                                self._p(R), self._p(t), self._p(K), self._r(res), voxel, trunc, maxw,
                                c.ctypes.data if counts else None)
        return c

    def compute_grads(self, tsdf, res):
        g = np.empty((tsdf.size, 3), dtype=np.float32)
        self.L.emfo_compute_grads(tsdf, g, self._r(res))
        return g

    def raycast(self, tsdf, grads, weights, R, t, K, res, voxel, trunc, w, h, raylengths=None, step_stats=False):
        ray = np.zeros((h, w), dtype=np.float32) if raylengths is None else np.ascontiguousarray(raylengths).copy()
        vert = np.zeros((h, w, 3), dtype=np.float32)
        norm = np.zeros((h, w, 3), dtype=np.float32)
        mask = np.zeros((h, w), dtype=np.uint8)
        hit = np.full((h, w, 3), -1, dtype=np.int32)
        steps = np.zeros(2, dtype=np.int64)
        step_img = np.zeros((h, w, 3), dtype=np.int32) if step_stats else None
        self.L.emfo_raycast(tsdf, grads, weights, ray, vert, norm, mask, w, h, self._p(R), self._p(t), self._p(K),
                            self._r(res), voxel, trunc, hit.ctypes.data, steps.ctypes.data,
                            step_img.ctypes.data if step_stats else None)
        return dict(ray=ray, vert=vert, norm=norm, mask=mask, hit=hit, steps=steps, step_img=step_img)

    def get_volume_vals(self, vol, points, R, t, res, voxel):
        h, w = points.shape[:2]
        out = np.empty((h, w), dtype=np.float32)
        nin = np.zeros(1, dtype=np.int64)
        self.L.emfo_get_volume_vals(vol, np.ascontiguousarray(points), w, h, self._p(R), self._p(t), self._r(res),
                                    voxel, out, nin.ctypes.data)
        return out, int(nin[0])

    def assoc_volume(self, tsdf, fg_probs, points, R, t, res, voxel, trunc, sigma=0.02, alpha=0.8, uni=1.0):
        h, w = points.shape[:2]
        out = np.empty((h, w), dtype=np.float32)
        m = np.empty((h, w), dtype=np.uint8)
        scratch = np.empty((h, w), dtype=np.float32)
        self.L.emfo_assoc_volume(tsdf, fg_probs.ctypes.data if fg_probs is not None else None,
                                 np.ascontiguousarray(points), w, h, self._p(R), self._p(t), self._r(res), voxel,
                                 trunc, sigma, alpha, uni, out, m.ctypes.data, scratch)
        return out, m

    def normalise(self, imgs, div0_is_zero=True):
        """imgs: list of (h, w) float32 arrays, normalised in place; returns the normaliser."""
        n = imgs[0].size
        ptrs = (C.c_void_p * len(imgs))(*[a.ctypes.data for a in imgs])
        norm = np.empty(imgs[0].shape, dtype=np.float32)
        self.L.emfo_normalise(ptrs, len(imgs), n, 1 if div0_is_zero else 0, norm.ctypes.data)
        return norm

    def composite(self, ids, obj_ray, obj_vert, obj_norm, obj_mask, bg_ray, bg_vert, bg_norm, bg_mask, boundary):
        h, w = bg_ray.shape
        k = len(ids)
        mk = lambda arrs: (C.c_void_p * max(k, 1))(*[a.ctypes.data for a in arrs])
        ray = np.zeros((h, w), dtype=np.float32)
        vert = np.zeros((h, w, 3), dtype=np.float32)
        norm = np.zeros((h, w, 3), dtype=np.float32)
        seg = np.zeros((h, w), dtype=np.uint8)
        vis = np.zeros(max(k, 1), dtype=np.int64)
        self.L.emfo_composite(k, np.asarray(ids if k else [0], dtype=np.int32), mk(obj_ray), mk(obj_vert),
                              mk(obj_norm), mk(obj_mask), bg_ray, bg_vert, bg_norm, bg_mask, w, h, boundary, ray,
                              vert, norm, seg, vis)
        return dict(ray=ray, vert=vert, norm=norm, seg=seg, vis=vis[:k])

    def update_fgbg(self, mask, occluded, tsdf, weights, fgbg, R, t, K, res, voxel):
        h, w = mask.shape
        self.L.emfo_update_fgbg(np.ascontiguousarray(mask), np.ascontiguousarray(occluded), w, h, tsdf, weights, fgbg,
                                self._p(R), self._p(t), self._p(K), self._r(res), voxel)

    def compute_fg_probs(self, fgbg):
        n = fgbg.size // 2
        p = np.empty(n, dtype=np.float32)
        m = np.empty(n, dtype=np.uint8)
        self.L.emfo_compute_fg_probs(fgbg, n, p, m)
        return p, m

    def raycast_weights(self, weights, fg_vol_mask):
        out = np.empty_like(weights)
        self.L.emfo_raycast_weights(weights, fg_vol_mask, weights.size, out)
        return out

    def track_linearise(self, tsdf, grads_vol, weights, points, assoc, R, t, res, voxel, huber=0.2, maxw=64.0):
        """one tracker iteration's device part; sums in float64"""
        h, w = points.shape[:2]
        n = h * w
        g6 = np.empty((n, 6), dtype=np.float32)
        vals = np.empty((h, w), dtype=np.float32)
        iw = np.empty((h, w), dtype=np.float32)
        tw = np.empty((h, w), dtype=np.float32)
        A = np.zeros(36); b = np.zeros(6); err = np.zeros(1); wmax = np.zeros(1)
        self.L.emfo_track_linearise(tsdf, np.ascontiguousarray(grads_vol).reshape(-1), weights, np.ascontiguousarray(points),
                                    np.ascontiguousarray(assoc), w, h, self._p(R), self._p(t), self._r(res), voxel, huber,
                                    maxw, g6, vals, iw, tw, A, b, err, wmax)
        return dict(grads=g6, tsdfVals=vals, intWeights=iw, trackWeights=tw, A=A.reshape(6, 6), b=b, err=float(err[0]),
                    wmax=float(wmax[0]))

    def preprocess_depth(self, raw, kernel_size=7, sigma_depth=0.04, sigma_spatial=4.5):
        h, w = raw.shape
        out = np.empty((h, w), dtype=np.float32)
        self.L.emfo_preprocess_depth(np.ascontiguousarray(raw, dtype=np.float32), w, h, kernel_size, sigma_depth, sigma_spatial, out)
        return out

    def uses_fma(self) -> bool:
        return bool(self.L.emfo_uses_fma())


def build():
    subprocess.run(["make", "-C", ODIR, "-s", "_build/libemf_oracle.so", "_build/libemf_oracle_nofma.so"], check=True)


def load(nofma: bool = False) -> Oracle:
    name = "libemf_oracle_nofma.so" if nofma else "libemf_oracle.so"
    path = os.path.join(ODIR, "_build", name)
    if not os.path.exists(path):
        build()
    return Oracle(path)
