"""NativeEngine -- the frame-level host mirror on top of the C++ frame engine (csrc/engine.cu, emf_engine_* in
include/emf_b200.h): one C call per frame (five launches, no device->host read on the path) instead of the
per-stage Python calls of engine.EMFusionEngine.  Same attribute names as EMFusionEngine (which mirror emf::EMFusion's
members, reference include/EMFusion/core/EMFusion.h:452-471); the image attributes are torch views of engine-owned
memory.

Multi-GPU (world_size > 1): objects are sharded round-robin (rank 0 also owns the background); the frame runs in
three phase groups around the two exchanges of SURVEY.md section 8e -- all-reduce of the per-pixel normaliser, gather
of the per-rank pre-composited raycast to rank 0 -- both through torch.distributed (NCCL) on the engine's stream.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib, ops
from ._lib import Image, Pose, Volume, check
from .engine import EMFusionEngine
from .poses import Affine, rel_pose_arrays
from .volume import ObjTSDF, Params

# phases (include/emf_b200.h)
F_POINTS, F_ASSOC, F_ASSOC_PARTIAL, F_NORMALISE, F_RAYCAST, F_COMPOSITE, F_INTEGRATE, F_INTEGRATE_ALL = (
    0x1, 0x2, 0x4, 0x8, 0x10, 0x20, 0x40, 0x80)
F_COMPOSITE_NOBG = 0x100
F_ASSOC_PARTIAL_NOBG = 0x400
F_TIMED = 0x200
F_INTEGRATE_BG, F_INTEGRATE_OBJ = 0x800, 0x1000
OPT_RAY_CERTIFICATE = 1
F_ALL = F_POINTS | F_ASSOC | F_RAYCAST | F_COMPOSITE | F_INTEGRATE
IMG_POINTS, IMG_NORM, IMG_RAY, IMG_VERT, IMG_NORMALS, IMG_SEG = range(6)
IMG_VOL_ASSOC, IMG_VOL_RAY, IMG_VOL_VERT, IMG_VOL_NORMALS, IMG_VOL_MASK = range(16, 21)


class _DevMem:
    """device memory -> torch tensor without a copy (CUDA array interface)"""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class NativeEngine(EMFusionEngine):
    def __init__(self, params: Params, device="cuda", rank: int = 0, world_size: int = 1, group=None,
                 materialize_grads: bool = False, accelerate: bool = False, replicate_background: Optional[bool] = None,
                 peer_exchange: Optional[bool] = None):
        """replicate_background (multi-GPU; default: on when the frame height divides by the world size): every rank
        keeps and integrates its own copy of the background (replicas stay bit-identical: same inputs, deterministic
        kernels) and raycasts a band of image rows of it; rank 0 gathers the bands.  Off = the background lives on
        rank 0 only (BASELINE.json's layout), which bounds the speed-up by the background's share of the frame."""
        if replicate_background is None:
            replicate_background = world_size > 1 and params.frameSize[1] % world_size == 0
        # multi-GPU exchanges over NVLink peer memory (csrc/xchg.cu) instead of NCCL collectives; None = use it if the
        # peers' buffers can be opened (same node, CUDA IPC), else NCCL
        self._peer_wanted = peer_exchange
        self._px = None
        super().__init__(params, device, rank, world_size, group, materialize_grads, accelerate,
                         replicate_background=replicate_background)
        L = _lib.lib()
        cfg = _lib.EngineConfig()
        cfg.width, cfg.height = self.w, self.h
        cfg.K[:] = np.asarray(params.intr, dtype=np.float32).reshape(9).tolist()
        cfg.params = params.tsdfParams.c()
        cfg.boundary = int(params.boundary)
        cfg.visibility_thresh = int(params.visibilityThresh)
        self._e = L.emf_engine_create(C.byref(cfg))
        if not self._e:
            raise _lib.EmfError("emf_engine_create failed")
        self._L = L
        self._dirty = True
        if self.replicate_background:
            rows = self.h // self.world
            check(L.emf_engine_set_background_rows(self._e, rank * rows, (rank + 1) * rows), "emf_engine_set_background_rows")
        self._T = None
        self._T_key = None
        self._marks = None
        self._bg_integrated = False
        self._defer_integrate = True      # stage-level calls (raycast(); integrateDepth()) keep the reference's order
        self._stage = (C.c_float * 3)()
        self._counts = (C.c_int32 * _lib.EMF_MAX_VOLUMES)()
        self._sync_volumes()

    def set_ray_certificate(self, on: Optional[bool]):
        """ray-space certificate + four-lanes-per-ray march for the background's raycast (csrc/raycast.cu: k_ray_certify,
        k_raycast_cert): True / False, None = the environment variable EMF_RAY_CERT decides.  Same results either way."""
        check(self._L.emf_engine_set_option(self._e, OPT_RAY_CERTIFICATE, -1 if on is None else int(bool(on))), "emf_engine_set_option")

    def __del__(self):
        e, self._e = getattr(self, "_e", None), None
        if e:
            self._L.emf_engine_destroy(e)

    # ---- volume list -> engine --------------------------------------------------------------------------
    def add_object(self, obj_pose: Affine, voxelSize: float, volumeRes=None) -> Optional[ObjTSDF]:
        obj = super().add_object(obj_pose, voxelSize, volumeRes)
        self._dirty = True
        self._new_ids = getattr(self, "_new_ids", set())
        if obj is not None:
            self._new_ids.add(obj.id)
        return obj

    def _view(self, what, index, shape, typestr):
        im = Image()
        check(self._L.emf_engine_image(self._e, what, index, C.byref(im)), "emf_engine_image")
        return torch.as_tensor(_DevMem(im.ptr, shape, typestr), device=self.device)

    def _sync_volumes(self):
        vols = self.local_volumes()
        arr = (Volume * max(len(vols), 1))()
        for i, v in enumerate(vols):
            arr[i] = v.c_volume(with_grads=True)
        self._keep = vols
        s = torch.cuda.current_stream(self.device).cuda_stream
        check(self._L.emf_engine_set_volumes(self._e, len(vols), arr, 1 if self.background is not None else 0, s),
              "emf_engine_set_volumes")
        h, w = self.h, self.w
        self.points = self._view(IMG_POINTS, 0, (h, w, 3), "<f4")
        self.associationNorm = self._view(IMG_NORM, 0, (h, w), "<f4")
        self.raylengths = self._view(IMG_RAY, 0, (h, w), "<f4")
        self.vertices = self._view(IMG_VERT, 0, (h, w, 3), "<f4")
        self.normals = self._view(IMG_NORMALS, 0, (h, w, 3), "<f4")
        self.modelSegmentation = self._view(IMG_SEG, 0, (h, w), "|u1")
        self.associationWeights, self.obj_raylengths, self.obj_vertices = {}, {}, {}
        self.obj_normals, self.obj_modelSegmentation = {}, {}
        for i, v in enumerate(vols):
            a = self._view(IMG_VOL_ASSOC, i, (h, w), "<f4")
            r = self._view(IMG_VOL_RAY, i, (h, w), "<f4")
            ve = self._view(IMG_VOL_VERT, i, (h, w, 3), "<f4")
            n = self._view(IMG_VOL_NORMALS, i, (h, w, 3), "<f4")
            m = self._view(IMG_VOL_MASK, i, (h, w), "|u1")
            if v.id == 0:
                self.bg_associationWeights, self.bg_raylengths, self.bg_vertices, self.bg_normals, self.bg_mask = a, r, ve, n, m
            else:
                self.associationWeights[v.id], self.obj_raylengths[v.id], self.obj_vertices[v.id] = a, r, ve
                self.obj_normals[v.id], self.obj_modelSegmentation[v.id] = n, m
        ptr = self._L.emf_engine_vis_counts_device(self._e)
        self.vis_count = torch.as_tensor(_DevMem(ptr, (_lib.EMF_MAX_VOLUMES,), "<i4"), device=self.device)
        for i, v in enumerate(vols):
            if v.id in getattr(self, "_new_ids", ()):
                check(self._L.emf_engine_force_integrate(self._e, i), "emf_engine_force_integrate")
        self._new_ids = set()
        self._dirty = False
        self._T = None
        self._gather_bufs = None
        self._pk = None

    # ---- one call = some phases of a frame ----------------------------------------------------------------
    def _frame(self, flags: int, depth: Optional[torch.Tensor] = None, host_depth: Optional[torch.Tensor] = None, download: bool = True):
        if self._dirty:
            self._sync_volumes()
        vols = self._keep
        # relative poses: recomputed whenever any pose differs from the ones they were computed from (the caller may assign
        # eng.pose / obj.pose between stage calls, as the reference does between tracking and the second association pass)
        key = b"".join([self.pose.R.tobytes(), self.pose.t.tobytes()] + [v.pose.R.tobytes() + v.pose.t.tobytes() for v in vols])
        if self._T is None or self._T_key != key:
            self._T = rel_pose_arrays(self.pose, [v.pose for v in vols]) if vols else (np.zeros((1, 12), np.float32),) * 2
            self._T_key = key
        T_co, T_oc = self._T
        if host_depth is not None:
            # host buffers at both ends (emf_engine_submit_host): upload, frame, download of the composite, all queued
            tk = C.c_longlong()
            check(self._L.emf_engine_submit_host(self._e, host_depth.data_ptr(), T_co.ctypes.data_as(C.POINTER(Pose)),
                                                 T_oc.ctypes.data_as(C.POINTER(Pose)), int(flags), 1 if download else 0,
                                                 torch.cuda.current_stream(self.device).cuda_stream, C.byref(tk)), "emf_engine_submit_host")
            self._ticket = int(tk.value)
            dv = self.__dict__.setdefault("_depth_views", {})
            self.depth = dv.get(self._ticket & 1)
            if self.depth is None:      # (two fixed device slots)
                im = Image()
                check(self._L.emf_engine_depth_slot(self._e, self._ticket, C.byref(im)), "emf_engine_depth_slot")
                self.depth = dv[self._ticket & 1] = torch.as_tensor(_DevMem(im.ptr, (self.h, self.w), "<f4"), device=self.device)
        else:
            d = depth if depth is not None else self.depth
            check(self._L.emf_engine_frame(self._e, C.byref(ops.image(d)), T_co.ctypes.data_as(C.POINTER(Pose)),
                                           T_oc.ctypes.data_as(C.POINTER(Pose)), int(flags),
                                           torch.cuda.current_stream(self.device).cuda_stream), "emf_engine_frame")
        # objects created since the last integrate are visible by definition until it has run (EMFusion.cpp:550,918)
        if flags & (F_COMPOSITE | F_COMPOSITE_NOBG):
            self._vis_extra = set(self._created)
        if flags & F_INTEGRATE:
            self._created = set()
        n = len(vols)
        launches = 0
        if flags & F_POINTS: launches += 1
        if flags & (F_ASSOC | F_ASSOC_PARTIAL | F_ASSOC_PARTIAL_NOBG) and n: launches += 1
        if flags & F_NORMALISE and n: launches += 1
        if flags & F_RAYCAST and n:   # (+ k_ray_certify and k_raycast_cert when the opt-in ray-space certificate is on)
            launches += 3 if (os.environ.get("EMF_RAY_CERT") == "1" and self.background is not None and n > 1) else (2 if os.environ.get("EMF_RAY_CERT") == "1" and self.background is not None else 1)
        if flags & (F_COMPOSITE | F_COMPOSITE_NOBG): launches += 1
        if flags & (F_INTEGRATE | F_INTEGRATE_BG | F_INTEGRATE_OBJ) and n:   # depth pyramid + brick classification + integrate (+ brick maps)
            launches += 3 + (2 if any(v.constBits is not None for v in vols) else 0)
        ops.LAUNCHES["engineFrame"] = ops.LAUNCHES.get("engineFrame", 0) + launches

    def set_depth(self, depth: torch.Tensor):
        self.depth = depth
        self._T = None          # a new frame: poses may have been reassigned by the caller
        self._frame(F_POINTS)

    def computeAssociationWeights(self):
        if self.world == 1:
            self._frame(F_ASSOC)
            return
        import torch.distributed as dist
        # (a replica of the background is left out of the partial sum: rank 0 adds the background's weight)
        px = self._peer_exchange()
        if px is not None:
            # the partial normaliser is written straight into this rank's exchange slot; the normalising kernel waits
            # for every rank's flag and sums the partial images out of the peers' memory itself (no all-reduce)
            if self._dirty:
                self._sync_volumes()
            px.begin_normaliser(self)
            self._frame(F_ASSOC_PARTIAL_NOBG if (self.replicate_background and self.rank != 0) else F_ASSOC_PARTIAL)
            px.finish_normaliser(self)
            return
        self._frame(F_ASSOC_PARTIAL_NOBG if (self.replicate_background and self.rank != 0) else F_ASSOC_PARTIAL)
        dist.all_reduce(self.associationNorm, op=dist.ReduceOp.SUM, group=self.group)
        self._frame(F_NORMALISE)

    def raycast(self):
        if self.world == 1:
            self._frame(F_RAYCAST | F_COMPOSITE)
            self._pending_vis = True
            return
        # every rank composites its own objects against an EMPTY background; rank 0 merges
        px = self._peer_exchange()
        if px is not None:
            if self._dirty:
                self._sync_volumes()
            px.begin_composite(self)            # raycast / pre-composite outputs go straight into this rank's exchange slot
            self._frame(F_RAYCAST | F_COMPOSITE_NOBG)
            px.signal_composite(self)
            self._mark("ray_local")
            # the background is gated by no visibility counter: integrate it now, under the other ranks' raycasts and the merge
            if self.background is not None and not self._defer_integrate:
                self._frame(F_INTEGRATE_BG)
                self._bg_integrated = True
            self._mark("int_bg")
            px.finish_composite(self)
            self._mark("merge")
            self._pending_vis = True
            return
        self._frame(F_RAYCAST | F_COMPOSITE_NOBG)
        self._composite_distributed()

    def _packed(self):
        """the pre-composite block [ray | vert | normals | seg] of the engine pool as one uint8 tensor (+ offsets)"""
        if getattr(self, "_pk", None) is not None:
            return self._pk
        p0 = self.raylengths.data_ptr()
        offs = (0, self.vertices.data_ptr() - p0, self.normals.data_ptr() - p0, self.modelSegmentation.data_ptr() - p0)
        size = offs[3] + self.h * self.w
        assert 0 < offs[1] < offs[2] < offs[3]
        self._pk = (torch.as_tensor(_DevMem(p0, (size,), "|u1"), device=self.device), offs, size)
        return self._pk

    def _composite_distributed(self):
        """gather of the per-rank pre-composites to rank 0 (one NCCL call), merge there (one launch, emf_composite_merge),
        visibility counts broadcast to every rank's device counters (the integrate gate); no host synchronisation"""
        import torch.distributed as dist
        h, w, dev = self.h, self.w, self.device
        packed, offs, size = self._packed()
        n_all = len(self.all_ids)
        rows = h // self.world if self.replicate_background else 0
        # send block: [ray | vert | normals | seg] of the pre-composite, then (replicated background) this rank's band of the
        # background's raycast [ray | vert | normals | mask]
        band_off = (0, rows * w * 4, rows * w * 16, rows * w * 28)
        total = size + rows * w * 29
        g = self._gather_bufs
        if g is None or g.get("total") != total or g["n_all"] != n_all:
            g = dict(total=total, n_all=n_all, counts=torch.zeros((max(n_all, 1),), dtype=torch.int32, device=dev),
                     local=torch.tensor([self.all_ids.index(o.id) for o in self.objects], dtype=torch.int64, device=dev),
                     send=torch.empty((total,), dtype=torch.uint8, device=dev) if rows else None)
            if self.rank == 0:
                g["all"] = torch.empty((self.world, total), dtype=torch.uint8, device=dev)
            if rows:
                y0 = self.rank * rows
                u8 = lambda t: t.reshape(-1).view(torch.uint8)
                g["pieces"] = [packed, u8(self.bg_raylengths[y0:y0 + rows]), u8(self.bg_vertices[y0:y0 + rows]),
                               u8(self.bg_normals[y0:y0 + rows]), u8(self.bg_mask[y0:y0 + rows])]
            self._gather_bufs = g
        if rows:
            torch.cat(g["pieces"], out=g["send"])
            send = g["send"]
        else:
            send = packed
        if self.rank == 0:
            if "args" not in g:     # everything the merge call needs is fixed until the volume list changes
                base = g["all"].data_ptr()
                mk = lambda r, o, el: Image(base + r * total + o, w * el, w, h)
                arr = lambda o, el: (Image * self.world)(*[mk(r, o, el) for r in range(self.world)])
                ptrs = lambda o: (C.c_void_p * self.world)(*[base + r * total + size + o for r in range(self.world)])
                g["list"] = list(g["all"].unbind(0))
                g["args"] = (self.world, arr(offs[0], 4), arr(offs[1], 12), arr(offs[2], 12), arr(offs[3], 1), n_all,
                             (C.c_int * max(n_all, 1))(*[int(i) for i in self.all_ids]), ops.image(self.bg_raylengths),
                             ops.image(self.bg_vertices), ops.image(self.bg_normals), ops.image(self.bg_mask),
                             int(self.params.boundary),
                             # the merged composite goes straight into the engine's frame images (rank 0's own pre-composite
                             # has been copied into the gather buffer like everyone else's)
                             ops.image(self.raylengths), ops.image(self.vertices), ops.image(self.normals),
                             ops.image(self.modelSegmentation), g["counts"].data_ptr(), rows,
                             ptrs(band_off[0]) if rows else None, ptrs(band_off[1]) if rows else None,
                             ptrs(band_off[2]) if rows else None, ptrs(band_off[3]) if rows else None)
            dist.gather(send, g["list"], dst=0, group=self.group)
            check(self._L.emf_composite_merge(*g["args"], torch.cuda.current_stream(dev).cuda_stream), "emf_composite_merge")
            ops.LAUNCHES["compositeMerge"] = ops.LAUNCHES.get("compositeMerge", 0) + 1
        else:
            dist.gather(send, None, dst=0, group=self.group)
        dist.broadcast(g["counts"], src=0, group=self.group)
        # this rank's objects, in its local list order, gate the integrate on the device
        if g["local"].numel():
            self.vis_count[:g["local"].numel()] = g["counts"][g["local"]]
        self._global_counts = g["counts"]
        self._pending_vis = True

    def _peer_exchange(self):
        """the NVLink peer-memory exchange of this engine (created on first use; None = NCCL collectives)"""
        if self._px is None and self._peer_wanted is not False and self.world > 1:
            self._px = PeerExchange.create(self)
            if self._px is None:
                if self._peer_wanted:
                    raise _lib.EmfError("peer_exchange=True but the peers' buffers cannot be opened (CUDA IPC)")
                self._peer_wanted = False
        return self._px

    def _resolve_visibility(self):
        """vis_objs for host-side bookkeeping (reads the asynchronously copied counters; not on the frame's path)"""
        if not getattr(self, "_pending_vis", False):
            return
        self._pending_vis = False
        if self.world == 1:
            n = len(self.objects)
            check(self._L.emf_engine_vis_counts(self._e, self._counts, n), "emf_engine_vis_counts")
            self._vis_objs = {o.id for o, c in zip(self.objects, self._counts[:n]) if c > self.params.visibilityThresh}
        else:
            cs = self._global_counts.cpu().numpy()[:len(self.all_ids)]
            self._vis_objs = {i for i, c in zip(self.all_ids, cs) if int(c) > self.params.visibilityThresh}
        self._vis_objs |= getattr(self, "_vis_extra", set())

    @property
    def vis_objs(self):
        self._resolve_visibility()
        return self._vis_objs

    @vis_objs.setter
    def vis_objs(self, v):
        self._vis_objs = set(v)
        self._pending_vis = False

    # ---- host buffers at both ends: the C ABI's emf_engine_submit_host / emf_engine_result_host -----------------------
    def submit_host(self, depth_host: torch.Tensor, cam_pose: Optional[Affine] = None, obj_poses: Optional[dict] = None,
                    download: bool = True) -> int:
        """processFrame with the depth image in page-locked HOST memory (H x W float32): upload, frame and download of the
        composite (segmentation + ray lengths) are queued on three streams inside the library; returns a ticket at once.
        Single GPU (the multi-GPU frame is several engine calls around the exchanges: pipeline.HostFramePipeline)."""
        if self.world != 1:
            raise _lib.EmfError("submit_host drives a single-GPU frame; use pipeline.HostFramePipeline for world_size > 1")
        if not (depth_host.is_pinned() and depth_host.dtype == torch.float32 and depth_host.is_contiguous()
                and tuple(depth_host.shape) == (self.h, self.w)):
            raise _lib.EmfError("depth_host must be a pinned, contiguous H x W float32 tensor")
        if cam_pose is not None:
            self.pose = cam_pose
        if obj_poses:
            for o in self.objects:
                if o.id in obj_poses:
                    o.pose = obj_poses[o.id]
        if self.frameCount == 0:
            self._frame(F_POINTS | F_INTEGRATE | F_INTEGRATE_ALL, host_depth=depth_host, download=download)
        else:
            self._frame(F_ALL, host_depth=depth_host, download=download)
            self._pending_vis = True
        if self.materialize_grads:
            for v in self._keep:
                v._grads_dirty = True
                v.updateGradients()
        self.frameCount += 1
        return self._ticket

    def result_host(self, ticket: int):
        """(segmentation uint8, ray lengths float32) of a submitted frame as tensors over the library's page-locked buffers;
        blocks until they have arrived.  Valid until two further frames have been submitted."""
        ps, pr = C.c_void_p(), C.c_void_p()
        rc = self._L.emf_engine_result_host(self._e, int(ticket), C.byref(ps), C.byref(pr))
        if rc == _lib.EMF_ERR_INVALID:
            raise ValueError("result of this frame is no longer (or not yet) available")
        check(rc, "emf_engine_result_host")
        views = self.__dict__.setdefault("_host_views", {})
        v = views.get(ps.value)
        if v is None:       # (the library double-buffers: two fixed pairs of page-locked buffers)
            n = self.h * self.w
            seg = np.ctypeslib.as_array((C.c_uint8 * n).from_address(ps.value)).reshape(self.h, self.w)
            ray = np.ctypeslib.as_array((C.c_float * n).from_address(pr.value)).reshape(self.h, self.w)
            v = views[ps.value] = (torch.from_numpy(seg), torch.from_numpy(ray))
        return v

    def _mark(self, name):
        """stage events of a timed multi-GPU frame (torch events on the frame's stream)"""
        if getattr(self, "_marks", None) is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(self.device))
            self._marks[name] = ev

    def integrateDepth(self, only_visible: bool = True):
        if getattr(self, "_bg_integrated", False) and only_visible:      # (the staged multi-GPU raycast() already integrated the background)
            self._bg_integrated = False
            self._frame(F_INTEGRATE_OBJ)
        else:
            self._frame(F_INTEGRATE | (0 if only_visible else F_INTEGRATE_ALL))
        for v in self._keep:
            v._grads_dirty = True
            v.updateGradients()

    def processFrame(self, depth: torch.Tensor, cam_pose: Optional[Affine] = None, obj_poses: Optional[dict] = None,
                     timed: bool = False):
        self.depth = depth
        if cam_pose is not None:
            self.pose = cam_pose
        if obj_poses:
            for o in self.objects:
                if o.id in obj_poses:
                    o.pose = obj_poses[o.id]
        self._T = None
        t = F_TIMED if timed else 0
        if self.frameCount == 0:
            self._frame(F_POINTS | F_INTEGRATE | F_INTEGRATE_ALL | t)
        elif self.world == 1:
            self._frame(F_ALL | t)
            self._pending_vis = True
        else:
            px = self._peer_exchange()
            if px is not None:
                px.poll_errors()            # a peer wait that timed out in an earlier frame surfaces here (no stall: asynchronous copy)
            if timed:       # (timed frames are issued back to back; stage_ms() reads all of them at once)
                self._marks = {}
                self._mark_log = getattr(self, "_mark_log", None) or []
                self._mark_log.append(self._marks)
            else:
                self._marks = None
            self._mark("start")
            self._frame(F_POINTS)
            self.computeAssociationWeights()
            self._mark("assoc")
            self._bg_integrated = False
            self._defer_integrate = False
            try:
                self.raycast()
            finally:
                self._defer_integrate = True
            self._frame(F_INTEGRATE_OBJ if self._bg_integrated else F_INTEGRATE)
            self._bg_integrated = False
            self._mark("end")
            if px is not None:
                px.snapshot_errors(self)
        if self.materialize_grads:
            for v in self._keep:
                v._grads_dirty = True
                v.updateGradients()
        self.frameCount += 1

    def stage_ms(self):
        """device ms of (association, raycast + composite, integrate) of the last timed frame.  Multi-GPU: this rank's times
        between the stage marks -- association incl. the normaliser exchange; local raycast + pre-composite, plus the merge /
        the wait for it; background integrate (under the merge) plus the objects' integrate"""
        if self.world > 1:
            log = [m for m in (getattr(self, "_mark_log", None) or []) if "end" in m]
            self._mark_log = []
            if not log:
                raise _lib.EmfError("no timed frame")
            log[-1]["end"].synchronize()
            out = np.zeros(3)
            detail = np.zeros(5)
            for m in log:       # mean over the timed frames issued since the last call
                dt = lambda a, b: m[a].elapsed_time(m[b]) if a in m and b in m else 0.0
                detail += [dt("start", "assoc"), dt("assoc", "ray_local"), dt("ray_local", "int_bg"), dt("int_bg", "merge"), dt("merge", "end")]
                if "ray_local" in m:
                    out += [dt("start", "assoc"), dt("assoc", "ray_local") + dt("int_bg", "merge"), dt("ray_local", "int_bg") + dt("merge", "end")]
                else:
                    out += [dt("start", "assoc"), 0.0, dt("assoc", "end")]
            # this rank's phases one by one: association + normaliser exchange, local raycast + pre-composite, background
            # integrate, merge (rank 0) / wait for it, objects' integrate
            self.stage_detail = dict(zip(("association", "raycast_local", "integrate_background", "merge_or_wait", "integrate_objects"),
                                         (float(x) for x in detail / len(log))))
            return [float(x) for x in out / len(log)]
        check(self._L.emf_engine_stage_ms(self._e, self._stage), "emf_engine_stage_ms")
        return [float(x) for x in self._stage]


class PeerExchange:
    """Per-rank exchange buffer + its peers' buffers (CUDA IPC), and the three exchanges of a sharded frame on top of them:
    the sum of the partial association normalisers, the composite merge reading every rank's pre-composite in place, and
    the visibility counters stored into every rank's buffer (csrc/xchg.cu).  All on the frame's stream, no host sync.

    Buffer layout (bytes): [flags 3 kinds x 16 ranks x u32 | pad to 256] [error word | pad to 256]
                           [counts 2 slots x 128 x i32] [normaliser 2 slots x W*H f32] [pre-composite 2 slots x `cap`]"""
    KIND_NORM, KIND_PRE, KIND_COUNTS = 0, 1, 2
    TIMEOUT_S = 60.0     # a rank may lag at start-up (cold imports); a real dead peer then surfaces as an error word, not a hang

    def __init__(self):
        pass

    @staticmethod
    def create(eng):
        import torch.distributed as dist
        L = _lib.lib()
        self = PeerExchange()
        self.L = L
        w, h, n = eng.w, eng.h, eng.world
        if n > 16:
            return None
        self.n, self.rank = n, eng.rank
        self.off_flags, self.off_err, self.off_counts = 0, 256, 512
        self.off_norm = 512 + 2 * 128 * 4
        self.norm_bytes = (w * h * 4 + 255) // 256 * 256
        self.off_pre = self.off_norm + 2 * self.norm_bytes
        self.cap = 2 * sum((w * h * el + 255) // 256 * 256 for el in (4, 12, 12, 1))    # pre-composite + background raycast, full frames
        total = self.off_pre + 2 * self.cap
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        ok = L.emf_xchg_alloc(total, C.byref(ptr), handle) == 0
        handles = [None] * n
        dist.all_gather_object(handles, handle.raw if ok else None, group=eng.group)
        if any(hd is None for hd in handles):
            if ok:
                L.emf_xchg_free(ptr)
            return None
        self.base = [0] * n
        self.base[self.rank] = ptr.value
        opened = True
        for r in range(n):
            if r == self.rank:
                continue
            p = C.c_void_p()
            if L.emf_xchg_open(handles[r], C.byref(p)) != 0:
                opened = False
                break
            self.base[r] = p.value
        flags = [None] * n
        dist.all_gather_object(flags, opened, group=eng.group)
        if not all(flags):
            return None
        self.local = torch.as_tensor(_DevMem(self.base[self.rank], (total,), "|u1"), device=eng.device)
        self.err = self.base[self.rank] + self.off_err
        self._sig = {}
        self._sum = {}
        self._merge = {}
        self.norm_seq = self.pre_seq = 0
        return self

    def _flag_ptr(self, owner, kind, src):
        return self.base[owner] + self.off_flags + 4 * (kind * 16 + src)

    def _signal(self, eng, kind, owners, epoch):
        key = (kind, tuple(owners))
        arr = self._sig.get(key)
        if arr is None:
            arr = (C.c_void_p * len(owners))(*[self._flag_ptr(o, kind, self.rank) for o in owners])
            self._sig[key] = arr
        check(self.L.emf_xchg_signal(len(owners), arr, epoch & 0xffffffff, torch.cuda.current_stream(eng.device).cuda_stream), "xchg_signal")

    def _wait(self, eng, kind, first, count, epoch):
        check(self.L.emf_xchg_wait(self._flag_ptr(self.rank, kind, first), count, epoch & 0xffffffff, self.err, self.TIMEOUT_S,
                                   torch.cuda.current_stream(eng.device).cuda_stream), "xchg_wait")

    def check_errors(self):
        """0 if no wait has timed out so far (reads one word; call off the frame path)"""
        return int(self.local[self.off_err:self.off_err + 4].view(torch.int32).item())

    def begin_normaliser(self, eng):
        """before EMF_FRAME_ASSOC_PARTIAL: its partial normaliser goes into this call's slot of the exchange buffer.
        Every rank makes the same sequence of calls: the call number is the flag value, its parity the slot (a slot is
        rewritten two calls later, after every peer has signalled -- i.e. finished reading -- the call in between)."""
        self.norm_seq += 1
        slot = self.norm_seq & 1
        img = Image(self.base[self.rank] + self.off_norm + slot * self.norm_bytes, eng.w * 4, eng.w, eng.h)
        check(self.L.emf_engine_set_partial_norm_target(eng._e, C.byref(img)), "set_partial_norm_target")

    def finish_normaliser(self, eng):
        """associationNorm <- sum over ranks (in rank order) of the partial normalisers, every association image divided by
        it: the all-reduce of EMFusion::computeAssociationWeights (src/core/EMFusion.cpp:653-665) inside its consumer"""
        epoch, slot = self.norm_seq, self.norm_seq & 1
        self._signal(eng, self.KIND_NORM, list(range(self.n)), epoch)
        arr = self._sum.get(slot)
        if arr is None:
            arr = (C.c_void_p * self.n)(*[self.base[r] + self.off_norm + slot * self.norm_bytes for r in range(self.n)])
            self._sum[slot] = arr
        check(self.L.emf_engine_normalise_from_parts(eng._e, self.n, arr, self._flag_ptr(self.rank, self.KIND_NORM, 0),
                                                     epoch & 0xffffffff, self.err, self.TIMEOUT_S,
                                                     torch.cuda.current_stream(eng.device).cuda_stream), "normalise_from_parts")
        ops.LAUNCHES["peerExchange"] = ops.LAUNCHES.get("peerExchange", 0) + 2

    def counts(self, eng):
        slot = self.pre_seq & 1
        o = self.off_counts + slot * 512
        return self.local[o:o + 512].view(torch.int32)

    # ---- the composite exchange.  A slot of this rank's buffer holds eight full-frame images:
    #      [pre-composite ray | vert | normals | seg] [background raycast ray | vert | normals | mask]
    def _slot_images(self, r, slot, w, h):
        base = self.base[r] + self.off_pre + slot * self.cap
        px = w * h
        up = lambda b: (b + 255) // 256 * 256
        o, out = 0, []
        for el in (4, 12, 12, 1, 4, 12, 12, 1):
            out.append(Image(base + o, w * el, w, h))
            o += up(px * el)
        return out

    def begin_composite(self, eng):
        """before EMF_FRAME_RAYCAST | EMF_FRAME_COMPOSITE_NOBG: the pre-composite and (replicated background) the background's
        raycast are written by their kernels straight into this call's slot of the exchange buffer"""
        self.pre_seq += 1
        slot = self.pre_seq & 1
        key = ("tgt", slot)
        t = self._merge.get(key)
        if t is None:
            im = self._slot_images(self.rank, slot, eng.w, eng.h)
            t = ((Image * 4)(*im[:4]), (Image * 4)(*im[4:]))
            self._merge[key] = t
        check(self.L.emf_engine_set_composite_target(eng._e, t[0]), "set_composite_target")
        if eng.replicate_background and eng.background is not None:
            check(self.L.emf_engine_set_background_target(eng._e, t[1]), "set_background_target")
        # the integrate is gated by the merged frame's counters, read where rank 0 stores them (this slot of this rank's buffer)
        idx = self._merge.get("gate_idx")
        if idx is None or idx[1] != tuple(v.id for v in eng._keep):
            pos = {i: k for k, i in enumerate(eng.all_ids)}
            arr = (C.c_int * max(len(eng._keep), 1))(*[pos.get(v.id, 0) for v in eng._keep])
            idx = (arr, tuple(v.id for v in eng._keep))
            self._merge["gate_idx"] = idx
        check(self.L.emf_engine_set_gate_source(eng._e, self.base[self.rank] + self.off_counts + slot * 512, idx[0], len(eng._keep)),
              "set_gate_source")

    def signal_composite(self, eng):
        self._signal(eng, self.KIND_PRE, [0], self.pre_seq)

    def finish_composite(self, eng):
        """rank 0 merges straight out of the peers' buffers (emf_composite_merge with peer pointers) into its engine's frame
        images and stores the visibility counters into every rank's buffer; the others wait for the counters"""
        epoch, slot = self.pre_seq, self.pre_seq & 1
        s = torch.cuda.current_stream(eng.device).cuda_stream
        w, h = eng.w, eng.h
        n_all = len(eng.all_ids)
        rows = h // self.n if eng.replicate_background else 0
        launches = 1
        if self.rank == 0:
            self._wait(eng, self.KIND_PRE, 0, self.n, epoch)
            key = ("merge", slot, n_all, tuple(eng.all_ids), rows)
            a = self._merge.get(key)
            if a is None:
                ims = [self._slot_images(r, slot, w, h) for r in range(self.n)]
                arr = lambda k: (Image * self.n)(*[ims[r][k] for r in range(self.n)])
                # band p of the background = rows [p * rows, (p + 1) * rows) of rank p's background images
                band = lambda k, el: (C.c_void_p * self.n)(*[ims[r][k].ptr + r * rows * w * el for r in range(self.n)])
                cnt = self.base[0] + self.off_counts + slot * 512
                if rows:
                    bg = [ops.image(eng.bg_raylengths), ops.image(eng.bg_vertices), ops.image(eng.bg_normals), ops.image(eng.bg_mask)]
                else:
                    bg = [ops.image(eng.bg_raylengths), ops.image(eng.bg_vertices), ops.image(eng.bg_normals), ops.image(eng.bg_mask)]
                args = (self.n, arr(0), arr(1), arr(2), arr(3), n_all, (C.c_int * max(n_all, 1))(*[int(i) for i in eng.all_ids]),
                        bg[0], bg[1], bg[2], bg[3], int(eng.params.boundary),
                        ops.image(eng.raylengths), ops.image(eng.vertices), ops.image(eng.normals), ops.image(eng.modelSegmentation),
                        cnt, rows, band(4, 4) if rows else None, band(5, 12) if rows else None, band(6, 12) if rows else None,
                        band(7, 1) if rows else None)
                others = [self.base[r] + self.off_counts + slot * 512 for r in range(1, self.n)]
                a = (args, cnt, (C.c_void_p * max(len(others), 1))(*others), len(others))
                self._merge[key] = a
            args, cnt, others, n_others = a
            check(self.L.emf_composite_merge(*args, s), "emf_composite_merge")
            if n_others and n_all:
                check(self.L.emf_xchg_scatter_u32(cnt, n_all, n_others, others, s), "xchg_scatter")
            self._signal(eng, self.KIND_COUNTS, list(range(1, self.n)), epoch)
            launches += 4
        else:
            self._wait(eng, self.KIND_COUNTS, 0, 1, epoch)
            launches += 1
        eng._global_counts = self.counts(eng)
        ops.LAUNCHES["peerExchange"] = ops.LAUNCHES.get("peerExchange", 0) + launches

    # ---- a timed-out wait sets the error word; it is copied out asynchronously after every frame and looked at before the next
    def snapshot_errors(self, eng):
        if getattr(self, "_err_host", None) is None:
            self._err_host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._err_event = torch.cuda.Event()
        self._err_host.copy_(self.local[self.off_err:self.off_err + 4].view(torch.int32), non_blocking=True)
        self._err_event.record(torch.cuda.current_stream(eng.device))

    def poll_errors(self):
        if getattr(self, "_err_host", None) is not None and self._err_event.query() and int(self._err_host[0]) != 0:
            raise _lib.EmfError(f"multi-GPU exchange: a wait for peer {int(self._err_host[0]) - 1} timed out; the frame used stale data")
