"""Deterministic synthetic RGB-D stream (SURVEY.md section 8d "Synthetic inputs").

An analytic scene -- three axis-aligned room planes plus K moving spheres -- rendered per pixel
to float32 z-depth in metres, with ground-truth camera / object poses (tracking bypassed) and
analytic instance masks (Mask R-CNN bypassed).  Pure numpy; used by tests, smoke() and bench.py to
produce the inputs both arms consume.  Nothing here is on the timed path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from .poses import Affine


@dataclass
class Sphere:
    centre0: np.ndarray      # world position at frame 0
    radius: float
    velocity: np.ndarray     # m / frame
    spin: np.ndarray         # Rodrigues vector / frame (invisible for a sphere; exercises pose math)

    def centre(self, frame: int) -> np.ndarray:
        return self.centre0 + self.velocity * frame

    def pose(self, frame: int) -> Affine:
        return Affine.from_rvec(self.spin * frame, self.centre(frame))


class Scene:
    """Room: back wall z = 4.2, floor y = +1.05 (camera y points down), side wall x = -1.9."""
    BACK_Z, FLOOR_Y, SIDE_X = 4.2, 1.05, -1.9

    def __init__(self, n_objects: int = 0, width: int = 640, height: int = 480, seed: int = 0,
                 radius_range: Tuple[float, float] = (0.15, 0.35), noise_sigma: float = 0.0,
                 dropout: float = 0.0, intr: Optional[np.ndarray] = None):
        self.w, self.h = width, height
        f = 525.0 * width / 640.0
        self.K = (np.array([[f, 0, width / 2 - 0.5], [0, f, height / 2 - 0.5], [0, 0, 1]], dtype=np.float32)
                  if intr is None else np.asarray(intr, dtype=np.float32))
        self.noise_sigma, self.dropout = noise_sigma, dropout
        self.seed = seed
        rng = np.random.default_rng(seed)
        self.spheres: List[Sphere] = []
        if n_objects > 0:
            cols = int(np.ceil(np.sqrt(n_objects * 2.0)))
            rows = int(np.ceil(n_objects / cols))
            lo, hi = radius_range
            if n_objects > 12:   # keep a crowded scene inside the frustum
                lo, hi = min(lo, 0.10), min(hi, 0.20)
            for k in range(n_objects):
                cu, cv = k % cols, k // cols
                u = ((cu + 0.5 + rng.uniform(-0.25, 0.25)) / cols) * 2 - 1
                v = ((cv + 0.5 + rng.uniform(-0.25, 0.25)) / rows) * 2 - 1
                z = rng.uniform(1.4, 3.2)
                c = np.array([u * z * 0.48, v * z * 0.33, z])
                r = rng.uniform(lo, hi)
                vel = rng.uniform(-1, 1, 3) * np.array([0.002, 0.001, 0.002])
                spin = rng.uniform(-1, 1, 3) * 0.004
                self.spheres.append(Sphere(c, float(r), vel, spin))

    # -- ground-truth trajectory: 10 cm circle in x/y with a small yaw/pitch wobble
    def cam_pose(self, frame: int) -> Affine:
        th = 2 * np.pi * frame / 60.0
        t = np.array([0.1 * np.cos(th) - 0.1, 0.1 * np.sin(th), 0.0])
        rvec = np.array([0.01 * np.sin(th), 0.015 * np.sin(2 * th), 0.005 * np.sin(th)])
        return Affine.from_rvec(rvec, t)

    def render(self, frame: int) -> Tuple[np.ndarray, np.ndarray]:
        """-> (depth (h, w) float32 metres, instance (h, w) uint8: 0 = room, k+1 = sphere k)."""
        cam = self.cam_pose(frame)
        xs, ys = np.meshgrid(np.arange(self.w, dtype=np.float64), np.arange(self.h, dtype=np.float64))
        dc = np.stack([(xs - self.K[0, 2]) / self.K[0, 0], (ys - self.K[1, 2]) / self.K[1, 1], np.ones_like(xs)], -1)
        dw = dc @ cam.R.T
        o = cam.t
        best = np.full((self.h, self.w), np.inf)
        with np.errstate(divide="ignore", invalid="ignore"):
            for axis, c, sign in ((2, self.BACK_Z, 1), (1, self.FLOOR_Y, 1), (0, self.SIDE_X, -1)):
                lam = (c - o[axis]) / dw[..., axis]
                ok = (sign * dw[..., axis] > 1e-9) & (lam > 0)
                best = np.where(ok & (lam < best), lam, best)
        inst = np.zeros((self.h, self.w), dtype=np.uint8)
        a = np.sum(dw * dw, -1)
        for k, s in enumerate(self.spheres):
            oc = o - s.centre(frame)
            b = 2 * (dw @ oc)
            cc = oc @ oc - s.radius ** 2
            disc = b * b - 4 * a * cc
            with np.errstate(invalid="ignore"):
                lam = (-b - np.sqrt(disc)) / (2 * a)
            ok = (disc > 0) & (lam > 0) & (lam < best)
            best = np.where(ok, lam, best)
            inst = np.where(ok, np.uint8(k + 1), inst)
        depth = np.where(np.isfinite(best), best, 0.0)   # dc.z == 1 => lambda is camera z-depth
        if self.noise_sigma > 0:
            depth = depth + np.random.default_rng(self.seed * 7919 + 1 + frame).normal(0, self.noise_sigma, depth.shape) * (depth > 0)
        if self.dropout > 0:
            drop = np.random.default_rng(self.seed * 7919 + 2 + 31 * frame).random(depth.shape) < self.dropout
            depth = np.where(drop, 0.0, depth)
        return depth.astype(np.float32), inst

    # -- object volumes as the reference would create them (src/core/EMFusion.cpp:537-547):
    #    world-aligned at creation, metric size volPad * largest extent
    def object_voxel_size(self, k: int, vol_res: int, vol_pad: float = 2.0) -> float:
        return float(np.float32(vol_pad * 2 * self.spheres[k].radius / vol_res))

    def object_pose(self, k: int, frame: int) -> Affine:
        return self.spheres[k].pose(frame)
