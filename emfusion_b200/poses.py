"""Rigid-pose algebra of the host mirror (stand-in for cv::Affine3f on the reference side).

The reference composes relative poses on the host before every operator call:
T_OC = cam_pose^-1 * pose (src/core/TSDF.cpp:112) and T_CO = pose^-1 * cam_pose
(src/core/TSDF.cpp:141,162).  Here poses are kept in float64 and rounded to float32 once,
when handed to the C ABI (which takes the already-composed relative pose, like the reference's
level-1 operators do).
"""
from __future__ import annotations

import numpy as np


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
        Rt = self.R.T
        return Affine(Rt, -Rt @ self.t)

    def __mul__(self, other: "Affine") -> "Affine":
        return Affine(self.R @ other.R, self.R @ other.t + self.t)

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
