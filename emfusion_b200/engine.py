"""Frame-level host mirror of the three hot methods of emf::EMFusion
(reference src/core/EMFusion.cpp): computeAssociationWeights (:635-670), raycast (:726-795)
and integrateDepth (:865-889), each as ONE or TWO batched launches over every volume of the
frame, plus the multi-GPU sharding of object volumes.

Sharding: rank r owns objects r, r+G, r+2G, ... (list order); rank 0 also owns the background.
The association needs one exchange -- the per-pixel normaliser, all-reduced (sum) over ranks --
and the composite one gather of the per-rank (raylength, vertex, normal, mask) images to rank 0.
Everything else is rank-local with no collective on the data path.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import ops
from .poses import Affine, rel_pose_CO, rel_pose_OC
from .volume import ObjTSDF, Params, TSDF


class EMFusionEngine:
    def __init__(self, params: Params, device="cuda", rank: int = 0, world_size: int = 1, group=None,
                 materialize_grads: bool = False, accelerate: bool = False, replicate_background: bool = False):
        self.params = params
        self.replicate_background = bool(replicate_background) and world_size > 1
        self.accelerate = accelerate
        self.device = torch.device(device)
        self.rank, self.world = rank, world_size
        self.group = group
        self.materialize_grads = materialize_grads
        w, h = params.frameSize
        self.w, self.h = w, h
        self.pose = Affine.identity()
        self.frameCount = 0
        f32, u8 = torch.float32, torch.uint8
        dev = self.device
        self.background: Optional[TSDF] = None
        if rank == 0 or self.replicate_background:
            # float * float as in the reference (src/core/EMFusion.cpp:31)
            self.background = TSDF(params.globalVolumeDims, params.globalVoxelSize,
                                   float(np.float32(params.globalRelTruncDist) * np.float32(params.globalVoxelSize)),
                                   params.volumePose,
                                   params.tsdfParams, params.frameSize, dev, materialize_grads, accelerate)
        self.objects: List[ObjTSDF] = []          # local shard, list order
        self.all_ids: List[int] = []              # global list order (ids), identical on every rank
        self.depth = torch.zeros((h, w), dtype=f32, device=dev)
        self.points = torch.zeros((h, w, 3), dtype=f32, device=dev)
        self.raylengths = torch.zeros((h, w), dtype=f32, device=dev)
        self.vertices = torch.zeros((h, w, 3), dtype=f32, device=dev)
        self.normals = torch.zeros((h, w, 3), dtype=f32, device=dev)
        self.modelSegmentation = torch.zeros((h, w), dtype=u8, device=dev)
        self.bg_raylengths = torch.zeros((h, w), dtype=f32, device=dev)
        self.bg_vertices = torch.zeros((h, w, 3), dtype=f32, device=dev)
        self.bg_normals = torch.zeros((h, w, 3), dtype=f32, device=dev)
        self.bg_mask = torch.zeros((h, w), dtype=u8, device=dev)
        self.associationNorm = torch.zeros((h, w), dtype=f32, device=dev)
        self.bg_associationWeights = torch.ones((h, w), dtype=f32, device=dev)
        self.associationWeights = {}               # id -> (h, w) f32
        self.obj_raylengths, self.obj_vertices, self.obj_normals, self.obj_modelSegmentation = {}, {}, {}, {}
        self.vis_count = torch.zeros((EMFusionEngine.MAX_OBJ,), dtype=torch.int32, device=dev)
        self.vis_objs = set()
        self._created = set()
        self._gather_bufs = None

    MAX_OBJ = 96

    # ---- object lifecycle (hot-path subset of createObj, src/core/EMFusion.cpp:908-920) --------
    def owner_of(self, list_index: int) -> int:
        return list_index % self.world

    def add_object(self, obj_pose: Affine, voxelSize: float, volumeRes=None) -> Optional[ObjTSDF]:
        """Register a new object on every rank; only the owning rank allocates its volume."""
        idx = len(self.all_ids)
        ObjTSDF.nextID = max(ObjTSDF.nextID, self.all_ids[-1] if self.all_ids else 0)
        res = tuple(volumeRes or self.params.objVolumeDims)
        new_id = ObjTSDF.nextID + 1
        self.all_ids.append(new_id)
        if self.owner_of(idx) != self.rank:
            ObjTSDF.nextID = new_id
            return None
        obj = ObjTSDF(res, voxelSize, float(np.float32(self.params.objRelTruncDist) * np.float32(voxelSize)), obj_pose,
                      self.params.tsdfParams,
                      self.params.frameSize, self.device, self.materialize_grads, self.accelerate)
        assert obj.id == new_id
        self.objects.append(obj)
        h, w, dev = self.h, self.w, self.device
        self.obj_raylengths[obj.id] = torch.zeros((h, w), dtype=torch.float32, device=dev)
        self.obj_vertices[obj.id] = torch.zeros((h, w, 3), dtype=torch.float32, device=dev)
        self.obj_normals[obj.id] = torch.zeros((h, w, 3), dtype=torch.float32, device=dev)
        self.obj_modelSegmentation[obj.id] = torch.zeros((h, w), dtype=torch.uint8, device=dev)
        self.associationWeights[obj.id] = torch.ones((h, w), dtype=torch.float32, device=dev)
        self.vis_objs.add(obj.id)
        self._created.add(obj.id)     # integrated at the next integrateDepth whatever a raycast in between says (EMFusion.cpp:550,918)
        return obj

    def local_volumes(self) -> List[TSDF]:
        return ([self.background] if self.background is not None else []) + list(self.objects)

    def num_voxels_local(self) -> int:
        return sum(v.numVoxels() for v in self.local_volumes())

    # ---- frame input ------------------------------------------------------------------------
    def set_depth(self, depth: torch.Tensor):
        """depth (h, w) f32 on this device (already preprocessed); computes the point image."""
        self.depth = depth
        ops.computePoints(self.depth, self.points, self.params.intr)

    # ---- EMFusion::computeAssociationWeights (src/core/EMFusion.cpp:635-670) -----------------
    def _assoc_images(self, vols):
        return [self.bg_associationWeights if v.id == 0 else self.associationWeights[v.id] for v in vols]

    def computeAssociationWeights(self):
        vols = self.local_volumes()
        prm = self.params.tsdfParams.c()
        if self.world == 1:
            ops.assocWeights([v.c_volume() for v in vols], [rel_pose_CO(self.pose, v.pose) for v in vols],
                             self.points, prm, self._assoc_images(vols), mode=0, norm=self.associationNorm)
            return
        import torch.distributed as dist
        if vols:
            ops.assocWeights([v.c_volume() for v in vols], [rel_pose_CO(self.pose, v.pose) for v in vols],
                             self.points, prm, self._assoc_images(vols), mode=1, norm=self.associationNorm)
        else:
            self.associationNorm.zero_()
        dist.all_reduce(self.associationNorm, op=dist.ReduceOp.SUM, group=self.group)
        if vols:
            ops.assocNormalise(self._assoc_images(vols), self.associationNorm)

    # ---- EMFusion::raycast (src/core/EMFusion.cpp:726-795) -----------------------------------
    def _rects(self, vols):
        return [ops.volumeScreenRect(v.volumeRes, v.voxelSize, rel_pose_CO(self.pose, v.pose), self.params.intr,
                                     self.w, self.h) for v in vols]

    def raycast(self):
        vols = self.local_volumes()
        rects = self._rects(vols)
        if vols:
            ray = [self.bg_raylengths if v.id == 0 else self.obj_raylengths[v.id] for v in vols]
            vert = [self.bg_vertices if v.id == 0 else self.obj_vertices[v.id] for v in vols]
            norm = [self.bg_normals if v.id == 0 else self.obj_normals[v.id] for v in vols]
            mask = [self.bg_mask if v.id == 0 else self.obj_modelSegmentation[v.id] for v in vols]
            if self.background is not None and list(rects[0]) != [0, 0, self.w, self.h]:
                self.bg_mask.zero_()   # the composite reads the background's mask over the whole frame
            ops.raycastVolumes([v.c_volume(with_grads=True) for v in vols],
                               [rel_pose_CO(self.pose, v.pose) for v in vols], self.params.intr, rects, ray, vert,
                               norm, mask)
        if self.world == 1:
            objs = self.objects
            o_rects = rects[1:] if self.background is not None else rects
            ops.raycastComposite([o.id for o in objs], o_rects, [self.obj_raylengths[o.id] for o in objs],
                                 [self.obj_vertices[o.id] for o in objs], [self.obj_normals[o.id] for o in objs],
                                 [self.obj_modelSegmentation[o.id] for o in objs], self.bg_raylengths,
                                 self.bg_vertices, self.bg_normals, self.bg_mask, self.params.boundary,
                                 self.raylengths, self.vertices, self.normals, self.modelSegmentation,
                                 self.vis_count)
            self._update_visibility([o.id for o in objs])
            return
        self._raycast_composite_distributed(rects)

    def _update_visibility(self, ids: Sequence[int]):
        """vis_objs from the device counters -- the one device->host read of the frame (the reference
        does one blocking countNonZero per object, src/core/EMFusion.cpp:778-791)."""
        n = len(ids)
        self.vis_objs = set()
        if n == 0:
            return
        counts = self.vis_count[:n].cpu().numpy()
        for i, c in zip(ids, counts):
            if int(c) > self.params.visibilityThresh:
                self.vis_objs.add(i)
        self.vis_objs |= self._created

    def _raycast_composite_distributed(self, rects):
        raise NotImplementedError("the staged engine is single-GPU; the multi-GPU frame is NativeEngine's (csrc/xchg.cu, emf_composite_merge)")

    # ---- EMFusion::integrateDepth (src/core/EMFusion.cpp:865-889) ----------------------------
    def integrateDepth(self, only_visible: bool = True):
        vols = [v for v in self.local_volumes() if v.id == 0 or not only_visible or v.id in self.vis_objs or v.id in self._created]
        self._created = set()
        if not vols:
            return
        cv = [v.c_volume() for v in vols]
        ops.integrateVolumes(cv, [rel_pose_OC(self.pose, v.pose) for v in vols],
                             self.params.intr, self.depth, self._assoc_images(vols),
                             self.params.tsdfParams.maxTSDFWeight)
        ops.updateBrickMaps(cv)
        for v in vols:
            v._grads_dirty = True
            v.updateGradients()

    # ---- hot subset of EMFusion::processFrame (src/core/EMFusion.cpp:70-129), poses given ------
    def processFrame(self, depth: torch.Tensor, cam_pose: Optional[Affine] = None,
                     obj_poses: Optional[dict] = None):
        self.set_depth(depth)
        if cam_pose is not None:
            self.pose = cam_pose
        if obj_poses:
            for o in self.objects:
                if o.id in obj_poses:
                    o.pose = obj_poses[o.id]
        if self.frameCount > 0:
            self.computeAssociationWeights()
            self.raycast()
        self.integrateDepth(only_visible=self.frameCount > 0)
        self.frameCount += 1
