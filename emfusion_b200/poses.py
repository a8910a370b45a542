"""Rigid-pose algebra of the host mirror (stand-in for cv::Affine3f on the reference side).

The reference composes relative poses on the host before every operator call:
T_OC = cam_pose^-1 * pose (src/core/TSDF.cpp:112) and T_CO = pose^-1 * cam_pose
(src/core/TSDF.cpp:141,162).  Here poses are kept in float64 and rounded to float32 once,
when handed to the C ABI (which takes the already-composed relative pose, like the reference's
level-1 operators do).
"""
from __future__ import annotations

import numpy as np


def _mm(A, B):
    """A @ B for (..., 3, 3) float64 arrays as ((a0 b0 + a1 b1) + a2 b2): a fixed evaluation order, so that the scalar
    path (Affine) and the batched path (rel_pose_arrays) give the same bits."""
    return (A[..., :, 0, None] * B[..., None, 0, :] + A[..., :, 1, None] * B[..., None, 1, :]) + A[..., :, 2, None] * B[..., None, 2, :]


def _mv(A, v):
    return (A[..., :, 0] * v[..., None, 0] + A[..., :, 1] * v[..., None, 1]) + A[..., :, 2] * v[..., None, 2]


class Affine:
    __slots__ = ("R", "t")

    def __init__(self, R=None, t=None):
        self.R = np.eye(3) if R is None else np.asarray(R, dtype=np.float64).reshape(3, 3)
        self.t = np.zeros(3) if t is None else np.asarray(t, dtype=np.float64).reshape(3)

    @staticmethod
    def identity() -> "Affine":
        return Affine()

    @staticmethod
    def translation(t) -> "Affine":
        return Affine(None, t)

    @staticmethod
    def from_rvec(rvec, t=None) -> "Affine":
        """Rodrigues vector -> rotation (as cv::Affine3f(rvec, t))."""
        r = np.asarray(rvec, dtype=np.float64)
        th = np.linalg.norm(r)
        if th < 1e-12:
            return Affine(None, t)
        k = r / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * (Kx @ Kx)
        return Affine(R, t)

    def inv(self) -> "Affine":
        Rt = np.ascontiguousarray(self.R.T)
        return Affine(Rt, -_mv(Rt, self.t))

    def __mul__(self, other: "Affine") -> "Affine":
        return Affine(_mm(self.R, other.R), _mv(self.R, other.t) + self.t)

    def rotation32(self) -> np.ndarray:
        return np.ascontiguousarray(self.R, dtype=np.float32).reshape(9)

    def translation32(self) -> np.ndarray:
        return np.ascontiguousarray(self.t, dtype=np.float32)

    def copy(self) -> "Affine":
        return Affine(self.R.copy(), self.t.copy())


def rel_pose_OC(cam_pose: Affine, pose: Affine) -> Affine:
    """cam_pose.inv() * pose -- volume -> camera (integrate, fg/bg update)."""
    return cam_pose.inv() * pose


def rel_pose_CO(cam_pose: Affine, pose: Affine) -> Affine:
    """pose.inv() * cam_pose -- camera -> volume (raycast, association)."""
    return pose.inv() * cam_pose


def rel_pose_arrays(cam_pose: Affine, poses):
    """Relative poses of many volumes at once: -> (T_co, T_oc), each (n, 12) float32 = n packed emf_pose (R row-major,
    then t).  Same arithmetic as rel_pose_CO / rel_pose_OC (bit-identical), without n Python round trips."""
    n = len(poses)
    Rv = np.empty((n, 3, 3)); tv = np.empty((n, 3))
    for i, p in enumerate(poses):
        Rv[i] = p.R; tv[i] = p.t
    Rc, tc = cam_pose.R, cam_pose.t
    # T_co = pose^-1 * cam
    Rvt = np.ascontiguousarray(np.swapaxes(Rv, 1, 2))
    tvi = -_mv(Rvt, tv)
    R_co = _mm(Rvt, Rc[None])
    t_co = _mv(Rvt, tc[None]) + tvi
    # T_oc = cam^-1 * pose
    Rct = np.ascontiguousarray(Rc.T)
    tci = -_mv(Rct, tc)
    R_oc = _mm(Rct[None], Rv)
    t_oc = _mv(Rct[None], tv) + tci[None]
    T_co = np.concatenate([R_co.reshape(n, 9), t_co], axis=1).astype(np.float32)
    T_oc = np.concatenate([R_oc.reshape(n, 9), t_oc], axis=1).astype(np.float32)
    return np.ascontiguousarray(T_co), np.ascontiguousarray(T_oc)


def pack_poses(poses) -> np.ndarray:
    """[Affine] -> (n, 12) float32 = n packed emf_pose (R row-major, then t)"""
    out = np.empty((len(poses), 12), dtype=np.float32)
    for i, p in enumerate(poses):
        out[i, :9] = p.R.reshape(9)
        out[i, 9:] = p.t
    return out
