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

    def _raycast_composite_distributed(self, rects):
        """Each rank pre-composites its own objects (list order is preserved inside a shard), the
        per-rank results are gathered on rank 0 and merged there in (raylength, list-order) order --
        which is what the reference's sequential 'strictly nearer or first in list wins' loop computes."""
        import torch.distributed as dist
        h, w, dev = self.h, self.w, self.device
        objs = self.objects
        o_rects = rects[1:] if self.background is not None else rects
        # local pre-composite against an empty background
        if self._gather_bufs is None:
            z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=dev)
            self._gather_bufs = dict(ray=z(h, w), vert=z(h, w, 3), norm=z(h, w, 3), seg=z(h, w, dt=torch.uint8),
                                     zray=z(h, w), zvert=z(h, w, 3), zmask=z(h, w, dt=torch.uint8),
                                     win=z(h, w, dt=torch.int32))
        g = self._gather_bufs
        # per-rank winner: encode the LIST INDEX (global) of the winner so rank 0 can break ties
        ops.raycastComposite([o.id for o in objs], o_rects, [self.obj_raylengths[o.id] for o in objs],
                             [self.obj_vertices[o.id] for o in objs], [self.obj_normals[o.id] for o in objs],
                             [self.obj_modelSegmentation[o.id] for o in objs], g["zray"], g["zvert"], g["zvert"],
                             g["zmask"], self.params.boundary, g["ray"], g["vert"], g["norm"], g["seg"],
                             self.vis_count)
        packed = torch.cat([g["ray"].reshape(-1), g["vert"].reshape(-1), g["norm"].reshape(-1),
                            g["seg"].reshape(-1).to(torch.float32)])
        if self.rank == 0:
            bufs = [torch.empty_like(packed) for _ in range(self.world)]
            dist.gather(packed, bufs, dst=0, group=self.group)
            self._merge_on_root(bufs)
        else:
            dist.gather(packed, None, dst=0, group=self.group)
        # visibility is a property of the final segmentation: rank 0 counts and broadcasts
        n_all = len(self.all_ids)
        counts = torch.zeros((max(n_all, 1),), dtype=torch.int32, device=dev)
        if self.rank == 0 and n_all:
            seg = self.modelSegmentation
            b = self.params.boundary
            inner = seg[b:h - b, b:w - b].reshape(-1).to(torch.int64)
            hist = torch.bincount(inner, minlength=256)
            ids = torch.tensor([min(i, 255) for i in self.all_ids], device=dev)
            counts[:n_all] = hist[ids].to(torch.int32)
        dist.broadcast(counts, src=0, group=self.group)
        cs = counts.cpu().numpy()
        self.vis_objs = {i for i, c in zip(self.all_ids, cs) if int(c) > self.params.visibilityThresh}

    def _merge_on_root(self, bufs):
        h, w = self.h, self.w
        n = h * w
        id_to_idx = {min(i, 255): k for k, i in enumerate(self.all_ids)}
        lut = torch.full((256,), 1 << 30, dtype=torch.int64, device=self.device)
        for i, k in id_to_idx.items():
            lut[i] = k
        best_ray = torch.zeros((n,), dtype=torch.float32, device=self.device)
        best_idx = torch.full((n,), 1 << 30, dtype=torch.int64, device=self.device)
        best_seg = torch.zeros((n,), dtype=torch.uint8, device=self.device)
        best_vert = torch.zeros((n, 3), dtype=torch.float32, device=self.device)
        best_norm = torch.zeros((n, 3), dtype=torch.float32, device=self.device)
        for buf in bufs:
            ray = buf[:n]
            vert = buf[n:4 * n].reshape(n, 3)
            norm = buf[4 * n:7 * n].reshape(n, 3)
            seg = buf[7 * n:8 * n].to(torch.uint8)
            has = seg != 0
            idx = lut[seg.to(torch.int64)]
            empty = best_seg == 0
            # sequential rule in list order == lexicographic min over (ray, list index), with ray <= 0 always replaced
            take = has & (empty | (best_ray <= 0) & (idx > best_idx) | (ray < best_ray) |
                          ((ray == best_ray) & (idx < best_idx) & ~(best_ray <= 0)))
            best_ray = torch.where(take, ray, best_ray)
            best_idx = torch.where(take, idx, best_idx)
            best_seg = torch.where(take, seg, best_seg)
            best_vert = torch.where(take[:, None], vert, best_vert)
            best_norm = torch.where(take[:, None], norm, best_norm)
        bgm = self.bg_mask.reshape(-1) != 0
        take_bg = bgm & ((best_ray - self.bg_raylengths.reshape(-1)) > 0.05)
        seg = torch.where(take_bg, torch.zeros_like(best_seg), best_seg)
        no_obj = seg == 0
        bgv = torch.where(bgm[:, None], self.bg_vertices.reshape(n, 3), torch.zeros_like(best_vert))
        bgn = torch.where(bgm[:, None], self.bg_normals.reshape(n, 3), torch.zeros_like(best_norm))
        self.raylengths.copy_(best_ray.reshape(h, w))
        self.modelSegmentation.copy_(seg.reshape(h, w))
        self.vertices.copy_(torch.where(no_obj[:, None], bgv, best_vert).reshape(h, w, 3))
        self.normals.copy_(torch.where(no_obj[:, None], bgn, best_norm).reshape(h, w, 3))

    # ---- EMFusion::integrateDepth (src/core/EMFusion.cpp:865-889) ----------------------------
    def integrateDepth(self, only_visible: bool = True):
        vols = [v for v in self.local_volumes() if v.id == 0 or not only_visible or v.id in self.vis_objs]
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
