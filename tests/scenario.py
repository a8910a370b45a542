"""Small deterministic scenarios shared by the parity tests: synthetic frames plus the volume state the
ORACLE reaches after integrating them (CPU).  Sizes are chosen so the oracle finishes in seconds."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from emfusion_b200.poses import Affine, rel_pose_CO, rel_pose_OC
from emfusion_b200.synth import Scene


@dataclass
class VolState:
    res: tuple
    voxel: float
    trunc: float
    pose: Affine
    vid: int
    tsdf: np.ndarray
    weights: np.ndarray
    fgbg: Optional[np.ndarray] = None
    fg_probs: Optional[np.ndarray] = None

    @property
    def n(self):
        return self.res[0] * self.res[1] * self.res[2]


@dataclass
class Scenario:
    name: str
    scene: Scene
    w: int
    h: int
    K: np.ndarray
    bg: VolState
    objs: List[VolState] = field(default_factory=list)
    depths: List[np.ndarray] = field(default_factory=list)
    insts: List[np.ndarray] = field(default_factory=list)

    def cam(self, f):
        return self.scene.cam_pose(f)

    def vols(self):
        return [self.bg] + self.objs


def R9(a: Affine):
    return a.rotation32()


def T3(a: Affine):
    return a.translation32()


def make(name, oracle, width, height, bg_res, n_obj, obj_res, n_frames=2, seed=0, dropout=0.0, noise=0.0,
         bg_size=5.12, integrate_frames=1) -> Scenario:
    scene = Scene(n_objects=n_obj, width=width, height=height, seed=seed, dropout=dropout, noise_sigma=noise)
    bg_res = tuple(bg_res)
    s = float(np.float32(bg_size / bg_res[0]))
    bg = VolState(bg_res, s, float(np.float32(10.0 * np.float32(s))), Affine.translation([0, 0, bg_size / 2]), 0,
                  np.zeros(int(np.prod(bg_res)), np.float32), np.zeros(int(np.prod(bg_res)), np.float32))
    sc = Scenario(name, scene, width, height, scene.K, bg)
    for k in range(n_obj):
        r = tuple(obj_res)
        vs = scene.object_voxel_size(k, r[0])
        n = int(np.prod(r))
        sc.objs.append(VolState(r, vs, float(np.float32(10.0 * np.float32(vs))), scene.object_pose(k, 0), k + 1,
                                np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(2 * n, np.float32),
                                np.zeros(n, np.float32)))
    for f in range(n_frames):
        d, i = scene.render(f)
        sc.depths.append(d)
        sc.insts.append(i)
    ones = np.ones((height, width), np.float32)
    zeros_u8 = np.zeros((height, width), np.uint8)
    for f in range(integrate_frames):
        cam = sc.cam(f)
        for v in sc.vols():
            if v.vid > 0:
                v.pose = scene.object_pose(v.vid - 1, f)
            T = rel_pose_OC(cam, v.pose)
            oracle.update_tsdf(sc.depths[f], ones, v.tsdf, v.weights, R9(T), T3(T), sc.K, v.res, v.voxel, v.trunc, 64.0)
        for v in sc.objs:
            T = rel_pose_OC(cam, v.pose)
            m = (sc.insts[f] == v.vid).astype(np.uint8)
            oracle.update_fgbg(m, zeros_u8, v.tsdf, v.weights, v.fgbg, R9(T), T3(T), sc.K, v.res, v.voxel)
            v.fg_probs, _ = oracle.compute_fg_probs(v.fgbg)
    return sc
