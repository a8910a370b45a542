"""Host mirror of the reference's volume classes -- same names, argument order and meaning as
emf::TSDF (reference include/EMFusion/core/TSDF.h:39-328, src/core/TSDF.cpp) and emf::ObjTSDF
(include/EMFusion/core/ObjTSDF.h:33-217, src/core/ObjTSDF.cpp), hot-path subset only.

Storage is torch CUDA memory in the reference's layout (rows = z*Ry + y, cols = x); compute is
the C ABI.  Differences from the reference that do not change results:
 * tsdfGrads is materialised lazily (getGrads()); the raycast takes forward differences on the
   fly, so updateGradients() is O(1) unless materialize_grads=True;
 * ObjTSDF keeps no raycastWeights / fgVolMask volumes -- the fgProb > 0.5 mask is applied inside
   the raycast kernel.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np
import torch

from . import ops
from .poses import Affine, rel_pose_CO, rel_pose_OC


@dataclass
class TSDFParams:
    """reference include/EMFusion/core/data.h:32-71 (defaults of the paper experiments)."""
    tau: float = 1e3
    eps1: float = 1e-8
    eps2: float = 1e-8
    nu_init: float = 2.0
    huberThresh: float = 0.2
    maxTSDFWeight: float = 64.0
    assocSigma: float = 0.02
    alpha: float = 0.8
    uniPrior: float = 1.0

    def c(self):
        return ops.tsdf_params(self.maxTSDFWeight, self.assocSigma, self.alpha, self.uniPrior)


@dataclass
class Params:
    """reference include/EMFusion/core/data.h:76-199 (fields the hot path consumes)."""
    frameSize: Tuple[int, int] = (640, 480)   # (width, height)
    intr: np.ndarray = None
    globalVolumeDims: Tuple[int, int, int] = (512, 512, 512)
    globalVoxelSize: float = 5.12 / 512
    globalRelTruncDist: float = 10.0
    objVolumeDims: Tuple[int, int, int] = (64, 64, 64)
    objRelTruncDist: float = 10.0
    volumePose: Affine = None
    volPad: float = 2.0
    visibilityThresh: int = 40 * 40
    boundary: int = 20
    tsdfParams: TSDFParams = field(default_factory=TSDFParams)

    def __post_init__(self):
        w, h = self.frameSize
        if self.intr is None:
            f = 525.0 * w / 640.0
            self.intr = np.array([[f, 0, w / 2 - 0.5], [0, f, h / 2 - 0.5], [0, 0, 1]], dtype=np.float32)
        if self.volumePose is None:
            vol = self.globalVoxelSize * self.globalVolumeDims[0]
            self.volumePose = Affine.translation([0, 0, vol / 2])


class TSDF:
    def __init__(self, volumeRes, voxelSize: float, truncdist: float, pose: Affine, params: TSDFParams,
                 frameSize, device="cuda", materialize_grads: bool = False, accelerate: bool = False):
        self.params = params
        self.volumeRes = tuple(int(r) for r in volumeRes)
        self.voxelSize = float(np.float32(voxelSize))
        self.truncdist = float(np.float32(truncdist))
        self.frameSize = tuple(frameSize)
        self.device = torch.device(device)
        rx, ry, rz = self.volumeRes
        self.tsdfVol = torch.empty((ry * rz, rx), dtype=torch.float32, device=self.device)
        self.tsdfWeights = torch.empty((ry * rz, rx), dtype=torch.float32, device=self.device)
        # acceleration state (no reference counterpart; results are unchanged): constant-segment bitmaps
        # maintained by the integrate kernel and the brick map the raycast jumps through
        self.constBits = self.brickMap = None
        if accelerate and rx % 4 == 0:
            words = ops.bitmapWords(self.volumeRes)
            self.constBits = torch.empty((3 * words,), dtype=torch.int32, device=self.device)
            self.brickMap = torch.empty((ops.brickMapBytes(self.volumeRes),), dtype=torch.uint8, device=self.device)
        self.materialize_grads = materialize_grads
        self._grads: Optional[torch.Tensor] = None
        self._grads_dirty = True
        self.id = 0
        self.reset(pose)

    # -- src/core/TSDF.cpp:74-79
    def reset(self, pose: Affine):
        self.tsdfVol.zero_()
        self.tsdfWeights.zero_()
        if self.constBits is not None:
            ops.resetBitmaps(self.c_volume())   # every segment is "all 0"; the brick map is rebuilt
        if self._grads is not None:
            self._grads.zero_()
        self._grads_dirty = False if self._grads is not None else True
        self.pose = pose.copy()

    def getCorners(self):
        c = (np.array(self.volumeRes, dtype=np.float32) - 1) * np.float32(self.voxelSize) / 2
        return -c, c

    def getVolumeSize(self):
        return np.array(self.volumeRes, dtype=np.float32) * np.float32(self.voxelSize)

    def getVolumeRes(self):
        return self.volumeRes

    def getVoxelSize(self):
        return self.voxelSize

    def getTruncDist(self):
        return self.truncdist

    def getPose(self) -> Affine:
        return self.pose

    def numVoxels(self) -> int:
        return self.volumeRes[0] * self.volumeRes[1] * self.volumeRes[2]

    # -- src/core/TSDF.cpp:108-118
    def integrate(self, depth, weights, cam_pose: Affine, intr, stream=None):
        v = [self.c_volume()]
        ops.integrateVolumes(v, [rel_pose_OC(cam_pose, self.pose)], intr, depth, [weights],
                             self.params.maxTSDFWeight, stream)
        ops.updateBrickMaps(v, stream)
        self._grads_dirty = True

    # -- src/core/TSDF.cpp:120-123
    def updateGradients(self, stream=None):
        if self.materialize_grads:
            self._materialize(stream)

    def _materialize(self, stream=None):
        if self._grads is None:
            rx, ry, rz = self.volumeRes
            self._grads = torch.empty((ry * rz, rx, 3), dtype=torch.float32, device=self.device)
        ops.computeTSDFGrads(self.tsdfVol, self._grads, self.volumeRes, stream)
        self._grads_dirty = False

    def getGrads(self) -> torch.Tensor:
        """tsdfGrads (float3 per voxel), brought up to date on demand."""
        if self._grads is None or self._grads_dirty:
            self._materialize()
        return self._grads

    def _raycast_grads(self):
        return self._grads if (self.materialize_grads and self._grads is not None and not self._grads_dirty) else None

    def c_volume(self, with_grads: bool = False):
        return ops.volume(self.tsdfVol, self.tsdfWeights, self.volumeRes, self.voxelSize, self.truncdist,
                          grads=self._raycast_grads() if with_grads else None, fg_probs=self._fg(), vid=self.id,
                          const_bits=self.constBits, brick_map=self.brickMap, fg_box=self._fg_box())

    def _fg(self):
        return None

    def _fg_box(self):
        return None

    # -- src/core/TSDF.cpp:125-156
    def computeAssociation(self, points, cam_pose: Affine, associationWeights, stream=None, associationMask=None):
        ops.computeAssociation(self.c_volume(), points, rel_pose_CO(cam_pose, self.pose), self.params.c(),
                               associationWeights, associationMask, stream)

    # -- src/core/TSDF.cpp:158-168
    def raycast(self, cam_pose: Affine, intr, raylengths, vertices, normals, mask, stream=None, hit_voxel=None):
        ops.raycastTSDF(self.tsdfVol, self._raycast_grads(), self.tsdfWeights, raylengths, vertices, normals, mask,
                        rel_pose_CO(cam_pose, self.pose), intr, self.volumeRes, self.voxelSize, self.truncdist,
                        fgProbs=self._fg(), hit_voxel=hit_voxel, stream=stream)

    # -- virtual accessors (src/core/TSDF.cpp:346-373): host copies of the volumes
    def getTSDF(self) -> np.ndarray:
        return self.tsdfVol.cpu().numpy()

    def getWeightsVol(self) -> np.ndarray:
        return self.tsdfWeights.cpu().numpy()

    # -- src/core/TSDF.cpp:356-373, src/core/ObjTSDF.cpp:247-268: marching cubes over the voxels with weight > 0 (objects: and
    #    fgProb > 0.5); cv::viz::Mesh's cloud / normals / polygons as arrays
    def getMesh(self):
        v, n, t = ops.marchingCubes(ops.volume(self.tsdfVol, self.tsdfWeights, self.volumeRes, self.voxelSize, self.truncdist,
                                               fg_probs=self._fg(), vid=self.id))
        return {"cloud": v, "normals": n, "polygons": t}


class ObjTSDF(TSDF):
    nextID = 0   # static counter, incremented only by the constructor (src/core/ObjTSDF.cpp:28,34)

    def __init__(self, volumeRes, voxelSize, truncdist, pose, params, frameSize, device="cuda",
                 materialize_grads: bool = False, accelerate: bool = False):
        rx, ry, rz = (int(r) for r in volumeRes)
        dev = torch.device(device)
        self.fgBgProbs = torch.empty((ry * rz, rx, 2), dtype=torch.float32, device=dev)
        self.fgProbs = torch.empty((ry * rz, rx), dtype=torch.float32, device=dev)
        # voxel bounds of {fgProb > 0.5}: rays that miss them cannot hit (raycast cull; results unchanged)
        self.fgBox = torch.tensor([1, 1, 1, 0, 0, 0], dtype=torch.int32, device=dev)
        self.classProbs = []
        self.exCount = 1
        self.nonExCount = 0
        super().__init__(volumeRes, voxelSize, truncdist, pose, params, frameSize, device, materialize_grads,
                         accelerate)
        ObjTSDF.nextID += 1
        self.id = ObjTSDF.nextID

    def __eq__(self, other):
        return isinstance(other, ObjTSDF) and self.id == other.id

    def __hash__(self):
        return hash(self.id)

    def getID(self) -> int:
        return self.id

    # -- src/core/ObjTSDF.cpp:56-59
    def reset(self, pose: Affine):
        super().reset(pose)
        self.fgBgProbs.zero_()
        self.fgProbs.zero_()
        self.fgBox.copy_(torch.tensor([1, 1, 1, 0, 0, 0], dtype=torch.int32))   # empty

    def getExProb(self) -> float:
        return float(self.exCount) / (self.exCount + self.nonExCount)

    def updateExProb(self, exists: bool):
        self.exCount += int(exists)
        self.nonExCount += 1 - int(exists)

    def _fg(self):
        return self.fgProbs

    def _fg_box(self):
        return self.fgBox

    # -- src/core/ObjTSDF.cpp:80-165
    def resize(self, p10, p90, volPad: float, stream=None) -> np.ndarray:
        """Grow / recentre the grid so that the box [p10, p90] (object coordinates) fits with padding volPad.  Returns the
        shift of the volume centre (zeros if the box was already contained).  Host arithmetic as in the reference (float32,
        cv::Vec3i conversions round to nearest even); the device part is one launch (emf_resize_volume).  A caller that holds
        this volume in an engine must re-submit the volume list afterwards (the arrays are new allocations)."""
        f32 = np.float32
        p10, p90 = np.asarray(p10, dtype=f32), np.asarray(p90, dtype=f32)
        res = np.array(self.volumeRes, dtype=f32)
        vs = f32(self.voxelSize)
        high = ((res - f32(1)) / f32(2)) * vs
        low = -high
        if bool(np.all(p10 >= low) and np.all(p90 <= high)):
            return np.zeros(3, dtype=f32)
        newCenter = (p10 + p90) / f32(2)
        pixOffset = np.rint(newCenter / vs).astype(np.int64)             # cv::Vec3i(Vec3f): saturate_cast<int> = cvRound
        newCenter = pixOffset.astype(f32) * vs
        self.pose = Affine(self.pose.R, self.pose.t + self.pose.R @ newCenter.astype(np.float64))   # pose.translate(R * c)
        newDims = p90 - p10
        newVolSize = f32(volPad) * newDims.max() / vs
        n = (int(np.ceil(newVolSize)) + 1) // 2 * 2
        newRes = np.array([n, n, n], dtype=np.int64)
        pixOffset = pixOffset - np.rint((newRes - np.array(self.volumeRes, dtype=np.int64)) * 0.5).astype(np.int64)   # Vec3i / 2
        dev = self.device
        newVol = torch.empty((n * n, n), dtype=torch.float32, device=dev)
        newWeights = torch.empty((n * n, n), dtype=torch.float32, device=dev)
        newFgBg = torch.empty((n * n, n, 2), dtype=torch.float32, device=dev)
        ops.resizeVolume(self.tsdfVol, self.tsdfWeights, self.fgBgProbs, self.volumeRes, newVol, newWeights, newFgBg,
                         (n, n, n), [int(v) for v in pixOffset], stream)
        self.tsdfVol, self.tsdfWeights, self.fgBgProbs = newVol, newWeights, newFgBg
        self.fgProbs = torch.empty((n * n, n), dtype=torch.float32, device=dev)
        self.volumeRes = (n, n, n)
        self._grads = None            # rebuilt on demand from the new tsdf (getGrads)
        self._grads_dirty = True
        if self.constBits is not None:
            if n % 4 == 0:
                self.constBits = torch.zeros((3 * ops.bitmapWords(self.volumeRes),), dtype=torch.int32, device=dev)   # nothing certified
                self.brickMap = torch.zeros((ops.brickMapBytes(self.volumeRes),), dtype=torch.uint8, device=dev)
            else:
                self.constBits = self.brickMap = None
        self.computeFgProbs(stream)
        return newCenter

    # -- src/core/ObjTSDF.cpp:167-179
    def integrateMask(self, mask, occluded_mask, cam_pose: Affine, intr, stream=None):
        ops.updateFgBgProbs(mask, occluded_mask, self.tsdfVol, self.tsdfWeights, self.fgBgProbs,
                            rel_pose_OC(cam_pose, self.pose), intr, self.volumeRes, self.voxelSize, stream)
        self.computeFgProbs(stream)

    # -- src/core/ObjTSDF.cpp:218-226
    def computeFgProbs(self, stream=None):
        ops.computeFgProbs(self.fgBgProbs, self.fgProbs, None, stream, fgBox=self.fgBox, volumeRes=self.volumeRes)

    def getFgProbVol(self) -> np.ndarray:
        return self.fgProbs.cpu().numpy()

    def getFgVolMask(self) -> torch.Tensor:
        return (self.fgProbs > 0.5).to(torch.uint8) * 255
