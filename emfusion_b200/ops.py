"""Level-1 operator API on torch CUDA tensors -- the Python mirror of
emf::cuda::TSDF::* / emf::cuda::ObjTSDF::* / emf::cuda::EMFusion::computePoints
(reference include/EMFusion/core/cuda/{TSDF,ObjTSDF,EMFusion}.cuh).  Every function is one
call through the C ABI on torch's current stream; tensors are only used for their device
pointer, shape and stride.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import Image, Pose, TsdfParams, Volume, check
from .poses import Affine


# kernels launched through this module, by C-ABI entry point (bench.py reports the sum as gpu_launches)
LAUNCHES = {}
_KERNELS_PER_CALL = {"computePoints": 1, "updateTSDF": 1, "computeTSDFGrads": 1, "raycastTSDF": 1, "getVolumeVals": 1,
                     "updateFgBgProbs": 1, "computeFgProbs": 1, "computeAssociation": 1, "assocWeights": 1,
                     "assocNormalise": 1, "marchingCubes": 3, "raycastVolumes": 1, "raycastComposite": 1, "integrateVolumes": 1,
                     "updateBrickMaps": 2, "trackLinearise": 1, "trackNormalisedWeights": 1, "copyValues": 1,
                     "resizeVolume": 1, "preprocessDepth": 1, "trackIterate": 2}


def _count(name: str, kernels: Optional[int] = None) -> None:
    LAUNCHES[name] = LAUNCHES.get(name, 0) + (_KERNELS_PER_CALL[name] if kernels is None else kernels)


def launches_total() -> int:
    return sum(LAUNCHES.values())


def _stream(stream=None) -> int:
    if stream is None:
        return torch.cuda.current_stream().cuda_stream
    if isinstance(stream, torch.cuda.Stream):
        return stream.cuda_stream
    return int(stream)


def image(t: torch.Tensor) -> Image:
    """2-D (H, W) or 3-D (H, W, C) CUDA tensor, rows contiguous -> emf_image."""
    if not t.is_cuda:
        raise _lib.EmfError("emfusion_b200 operates on CUDA tensors only (no CPU fallback)")
    if t.dim() == 3:
        if t.stride(2) != 1 or t.stride(1) != t.shape[2]:
            raise _lib.EmfError("image channels must be interleaved and contiguous")
    elif t.dim() == 2:
        if t.stride(1) != 1:
            raise _lib.EmfError("image rows must be contiguous")
    else:
        raise _lib.EmfError("image must be (H, W) or (H, W, C)")
    return Image(t.data_ptr(), t.stride(0) * t.element_size(), t.shape[1], t.shape[0])


def images(ts: Sequence[torch.Tensor]):
    arr = (Image * len(ts))()
    for i, t in enumerate(ts):
        arr[i] = image(t)
    return arr


def pose(a: Affine) -> Pose:
    p = Pose()
    p.R[:] = a.rotation32().tolist()
    p.t[:] = a.translation32().tolist()
    return p


def poses(aa: Sequence[Affine]):
    arr = (Pose * len(aa))()
    for i, a in enumerate(aa):
        arr[i] = pose(a)
    return arr


def _f9(K) -> C.Array:
    return (C.c_float * 9)(*np.asarray(K, dtype=np.float32).reshape(9).tolist())


def _i3(res) -> C.Array:
    return (C.c_int * 3)(*[int(r) for r in res])


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda or not t.is_contiguous():
        raise _lib.EmfError("volume arrays must be contiguous CUDA tensors")
    return t.data_ptr()


def volume(tsdf, weights, res, voxel_size, truncdist, grads=None, fg_probs=None, vid=0, const_bits=None,
           brick_map=None, fg_box=None) -> Volume:
    v = Volume()
    v.tsdf = _ptr(tsdf)
    v.weights = _ptr(weights)
    v.grads = _ptr(grads)
    v.fg_probs = _ptr(fg_probs)
    v.fg_box = _ptr(fg_box)
    v.const_bits = _ptr(const_bits)
    v.brick_map = _ptr(brick_map)
    v.res[:] = [int(r) for r in res]
    v.voxel_size = float(voxel_size)
    v.truncdist = float(truncdist)
    v.id = int(vid)
    return v


def tsdf_params(max_tsdf_weight=64.0, assoc_sigma=0.02, alpha=0.8, uni_prior=1.0) -> TsdfParams:
    return TsdfParams(max_tsdf_weight, assoc_sigma, alpha, uni_prior)


# ---- level 1 --------------------------------------------------------------------------------
def computePoints(depth, points, intr, stream=None):
    check(_lib.lib().emf_compute_points(image(depth), image(points), _f9(intr), _stream(stream)), "computePoints")
    _count("computePoints")


def updateTSDF(depth, assocWeights, tsdfVol, tsdfWeights, rel_pose_OC: Affine, intr, volumeRes, voxelSize,
               truncdist, maxWeight, stream=None):
    check(_lib.lib().emf_update_tsdf(image(depth), image(assocWeights), _ptr(tsdfVol), _ptr(tsdfWeights),
                                     pose(rel_pose_OC), _f9(intr), _i3(volumeRes), voxelSize, truncdist, maxWeight,
                                     _stream(stream)), "updateTSDF")
    _count("updateTSDF")


def computeTSDFGrads(tsdfVol, tsdfGrads, volumeRes, stream=None):
    check(_lib.lib().emf_compute_tsdf_grads(_ptr(tsdfVol), _ptr(tsdfGrads), _i3(volumeRes), _stream(stream)),
          "computeTSDFGrads")
    _count("computeTSDFGrads")


def raycastTSDF(tsdfVol, tsdfGrads, tsdfWeights, raylengths, vertices, normals, mask, rel_pose_CO: Affine, intr,
                volumeRes, voxelSize, truncdist, fgProbs=None, hit_voxel=None, stream=None):
    check(_lib.lib().emf_raycast_tsdf(_ptr(tsdfVol), _ptr(tsdfGrads), _ptr(tsdfWeights), _ptr(fgProbs),
                                      image(raylengths), image(vertices), image(normals), image(mask),
                                      pose(rel_pose_CO), _f9(intr), _i3(volumeRes), voxelSize, truncdist,
                                      _ptr(hit_voxel), _stream(stream)), "raycastTSDF")
    _count("raycastTSDF")


def getVolumeVals(vol, points, rel_pose_CO: Affine, volumeRes, voxelSize, vals, stream=None):
    check(_lib.lib().emf_get_volume_vals(_ptr(vol), image(points), pose(rel_pose_CO), _i3(volumeRes), voxelSize,
                                         image(vals), _stream(stream)), "getVolumeVals")
    _count("getVolumeVals")


def updateFgBgProbs(mask, occluded_mask, tsdfVol, tsdfWeights, fgBgProbs, rel_pose_OC: Affine, intr, volumeRes,
                    voxelSize, stream=None):
    check(_lib.lib().emf_update_fgbg_probs(image(mask), image(occluded_mask), _ptr(tsdfVol), _ptr(tsdfWeights),
                                           _ptr(fgBgProbs), pose(rel_pose_OC), _f9(intr), _i3(volumeRes), voxelSize,
                                           _stream(stream)), "updateFgBgProbs")
    _count("updateFgBgProbs")


def computeFgProbs(fgBgProbs, fgProbs, fgVolMask=None, stream=None, fgBox=None, volumeRes=None):
    """fgBox (6 x int32 CUDA tensor, needs volumeRes): also the voxel bounds of {fgProb > 0.5} (raycast cull)."""
    if fgBox is None:
        check(_lib.lib().emf_compute_fg_probs(_ptr(fgBgProbs), fgProbs.numel(), _ptr(fgProbs), _ptr(fgVolMask),
                                              _stream(stream)), "computeFgProbs")
        _count("computeFgProbs")
    else:
        check(_lib.lib().emf_compute_fg_probs_box(_ptr(fgBgProbs), _i3(volumeRes), _ptr(fgProbs), _ptr(fgVolMask),
                                                  _ptr(fgBox), _stream(stream)), "computeFgProbs")
        _count("computeFgProbs")
        _count("computeFgProbs")


# ---- level 2 / 3 ----------------------------------------------------------------------------
def computeAssociation(vol: Volume, points, rel_pose_CO: Affine, params: TsdfParams, associationWeights,
                       associationMask=None, stream=None):
    check(_lib.lib().emf_compute_association(C.byref(vol), image(points), pose(rel_pose_CO), C.byref(params),
                                             image(associationWeights),
                                             image(associationMask) if associationMask is not None else None,
                                             _stream(stream)), "computeAssociation")
    _count("computeAssociation")


def _vol_array(vols: Sequence[Volume]):
    arr = (Volume * len(vols))()
    for i, v in enumerate(vols):
        arr[i] = v
    return arr


def assocWeights(vols, rel_poses_CO, points, params: TsdfParams, assoc_out, mode=0, norm=None, stream=None):
    check(_lib.lib().emf_assoc_weights(len(vols), _vol_array(vols), poses(rel_poses_CO), image(points),
                                       C.byref(params), images(assoc_out), mode,
                                       image(norm) if norm is not None else None, _stream(stream)), "assocWeights")
    _count("assocWeights")


def assocNormalise(assoc_io, norm, stream=None):
    check(_lib.lib().emf_assoc_normalise(len(assoc_io), images(assoc_io), image(norm), _stream(stream)),
          "assocNormalise")
    _count("assocNormalise")


def volumeScreenRect(volumeRes, voxelSize, rel_pose_CO: Affine, intr, width, height):
    out = (C.c_int * 4)()
    check(_lib.lib().emf_volume_screen_rect(_i3(volumeRes), voxelSize, pose(rel_pose_CO), _f9(intr), width, height,
                                            out), "volumeScreenRect")
    return list(out)


def raycastWorkspace(width, height, device):
    """device scratch for raycastVolumes(workspace=...): the ray-space certificate table (csrc/raycast.cu, k_ray_certify) and the
    longest-first tile schedule (k_ray_schedule), zeroed; keep it between frames"""
    n = int(_lib.lib().emf_raycast_workspace_bytes(int(width), int(height)))
    return torch.zeros(n, dtype=torch.uint8, device=device)


def raycastVolumes(vols, rel_poses_CO, intr, rects, ray_out, vert_out, norm_out, mask_out, stream=None, stats=None,
                   workspace=None, certificate=True, schedule=True, wide=False):
    """workspace (raycastWorkspace): certificate = the background's rays skip certified free-space samples, schedule = the first
    volume's tiles are marched longest first by the previous call's cost; wide = four lanes per background ray (no workspace
    needed).  Same results whatever the options."""
    if stats is not None and (stats.numel() < 8 or stats.element_size() != 8):
        raise _lib.EmfError("raycast stats must be a tensor of at least 8 64-bit counters")
    flat = (C.c_int * (4 * len(vols)))(*[int(v) for r in rects for v in r]) if rects is not None else None
    check(_lib.lib().emf_raycast_volumes_opt(len(vols), _vol_array(vols), poses(rel_poses_CO), _f9(intr), flat,
                                             images(ray_out), images(vert_out), images(norm_out), images(mask_out),
                                             _ptr(stats), _ptr(workspace), workspace.numel() if workspace is not None else 0,
                                             (1 if certificate else 0) | (2 if schedule else 0) | (4 if wide else 0), _stream(stream)),
          "raycastVolumes")
    _count("raycastVolumes", 1 if workspace is None else 2)


def raycastComposite(ids, rects, obj_ray, obj_vert, obj_norm, obj_mask, bg_ray, bg_vert, bg_norm, bg_mask, boundary,
                     ray, vert, norm, seg, vis_count, stream=None):
    n = len(ids)
    ids_a = (C.c_int * max(n, 1))(*[int(i) for i in ids])
    flat = (C.c_int * max(4 * n, 1))(*[int(v) for r in rects for v in r]) if rects is not None else None
    check(_lib.lib().emf_raycast_composite(n, ids_a, flat, images(obj_ray) if n else None,
                                           images(obj_vert) if n else None, images(obj_norm) if n else None,
                                           images(obj_mask) if n else None, image(bg_ray), image(bg_vert),
                                           image(bg_norm), image(bg_mask), boundary, image(ray), image(vert),
                                           image(norm), image(seg), _ptr(vis_count), _stream(stream)),
          "raycastComposite")
    _count("raycastComposite")


_WORKSPACES = {}


def integrateWorkspace(depth):
    """cached device workspace (depth pyramid) for frames of this size on this device"""
    key = (depth.device, depth.shape[0], depth.shape[1])
    ws = _WORKSPACES.get(key)
    if ws is None:
        n = int(_lib.lib().emf_integrate_workspace_bytes(int(depth.shape[1]), int(depth.shape[0])))
        ws = torch.empty((n,), dtype=torch.uint8, device=depth.device)
        _WORKSPACES[key] = ws
    return ws


def integrateVolumes(vols, rel_poses_OC, intr, depth, assoc, maxWeight, stream=None, gate_counts=None, gates=None,
                     gate_thresh=0, stats=None, workspace="auto"):
    """gate_counts (int32 CUDA tensor) + gates (per-volume index or -1): device-side visibility filter;
    stats: optional uint64/int64 CUDA tensor of 8 counters (accumulated; see include/emf_b200.h);
    workspace: "auto" (cached depth-pyramid workspace -> segment-level kernel), None (per-voxel kernel) or a uint8 tensor."""
    if stats is not None and (stats.numel() < 8 or stats.element_size() != 8):
        raise _lib.EmfError("integrate stats must be a tensor of at least 8 64-bit counters")
    if workspace == "auto":
        workspace = integrateWorkspace(depth)
    if workspace is not None:
        g = (C.c_int * len(vols))(*[int(x) for x in gates]) if gates is not None else None
        check(_lib.lib().emf_integrate_volumes_ws(len(vols), _vol_array(vols), poses(rel_poses_OC), _f9(intr),
                                                  image(depth), images(assoc), maxWeight, _ptr(gate_counts), g,
                                                  int(gate_thresh), _ptr(stats), workspace.data_ptr(), workspace.numel(),
                                                  _stream(stream)), "integrateVolumes")
        _count("integrateVolumes")
        _count("integrateVolumes")
        return
    if gate_counts is None and stats is None:
        check(_lib.lib().emf_integrate_volumes(len(vols), _vol_array(vols), poses(rel_poses_OC), _f9(intr),
                                               image(depth), images(assoc), maxWeight, _stream(stream)),
              "integrateVolumes")
    else:
        g = (C.c_int * len(vols))(*[int(x) for x in gates]) if gates is not None else None
        check(_lib.lib().emf_integrate_volumes_gated(len(vols), _vol_array(vols), poses(rel_poses_OC), _f9(intr),
                                                     image(depth), images(assoc), maxWeight, _ptr(gate_counts), g,
                                                     int(gate_thresh), _ptr(stats), _stream(stream)),
              "integrateVolumes")
    _count("integrateVolumes")


def updateBrickMaps(vols, stream=None):
    """brick_map <- const_bits for every volume that carries both (two launches)."""
    check(_lib.lib().emf_update_brick_maps(len(vols), _vol_array(vols), _stream(stream)), "updateBrickMaps")
    if any(v.const_bits and v.brick_map for v in vols):
        _count("updateBrickMaps")


def brickMapBytes(res) -> int:
    return int(_lib.lib().emf_brick_map_bytes(_i3(res)))


def resetBitmaps(vol: Volume, stream=None):
    check(_lib.lib().emf_reset_bitmaps(C.byref(vol), _stream(stream)), "resetBitmaps")


def bitmapWords(res) -> int:
    """32-bit words of ONE segment bitmap of a volume (emf_bitmap_words_per_row(Rx) * Ry * Rz)."""
    return ((int(res[0]) // 4 + 31) // 32) * int(res[1]) * int(res[2])


def preprocessDepth(depth_raw, depth, points=None, intr=None, kernel_size=7, sigma_depth=0.04, sigma_spatial=4.5, stream=None):
    """EMFusion::preprocessDepth (+ computePoints when `points` is given) in one launch"""
    check(_lib.lib().emf_preprocess_depth(image(depth_raw), image(depth), image(points) if points is not None else None,
                                          _f9(intr) if intr is not None else None, int(kernel_size), float(sigma_depth),
                                          float(sigma_spatial), _stream(stream)), "preprocessDepth")
    _count("preprocessDepth")


# ---- resize -----------------------------------------------------------------------------------
def copyValues(src, dst, offset, srcRes, dstRes, stream=None):
    """emf::cuda::TSDF::copyValues: dst(x - offset) = src(x) where the target exists; channels from the last dimension"""
    ch = 1 if src.dim() == 2 else int(src.shape[-1])
    check(_lib.lib().emf_copy_values(_ptr(src), _ptr(dst), ch, _i3(offset), _i3(srcRes), _i3(dstRes), _stream(stream)),
          "copyValues")
    _count("copyValues")


def resizeVolume(srcTsdf, srcWeights, srcFgBg, srcRes, dstTsdf, dstWeights, dstFgBg, dstRes, offset, stream=None):
    check(_lib.lib().emf_resize_volume(_ptr(srcTsdf), _ptr(srcWeights), _ptr(srcFgBg), _i3(srcRes), _ptr(dstTsdf),
                                       _ptr(dstWeights), _ptr(dstFgBg), _i3(dstRes), _i3(offset), _stream(stream)),
          "resizeVolume")
    _count("resizeVolume")


# ---- tracker ----------------------------------------------------------------------------------
_TRACK_WS = {}


def trackWorkspace(device, n_vol: int):
    """cached, once-zeroed device workspace of emf_track_linearise for up to n_vol volumes"""
    key = device
    ws = _TRACK_WS.get(key)
    need = int(_lib.lib().emf_track_workspace_bytes(int(n_vol)))
    if ws is None or ws.numel() < need:
        ws = torch.empty((int(_lib.lib().emf_track_workspace_bytes(_lib.EMF_MAX_VOLUMES)),), dtype=torch.uint8, device=device)
        check(_lib.lib().emf_track_workspace_init(ws.data_ptr(), ws.numel(), _stream()), "trackWorkspaceInit")
        _TRACK_WS[key] = ws
    return ws


def _opt_images(ts, n):
    """array of n emf_image; entries for None tensors have ptr == NULL"""
    if ts is None:
        return None
    arr = (Image * n)()
    for i, t in enumerate(ts):
        arr[i] = image(t) if t is not None else Image(None, 0, 0, 0)
    return arr


def trackLinearise(vols, rel_poses_CO, modes, points, assoc, huberThresh, maxTSDFWeight, intWeights, records,
                   tsdfVals=None, trackWeights=None, poseGrads=None, stream=None, intr=None):
    """One launch = the device part of one tracker iteration of every volume (emf_track_linearise, include/emf_b200.h).
    modes[i]: 0 skip / 1 linearise / 2 error only; records: (n_vol, 48) float32 CUDA tensor; intr: the camera matrix of
    `points` (optional: lets the kernel skip image tiles that cannot see a volume)."""
    n = len(vols)
    if records.dtype != torch.float32 or records.numel() < n * _lib.EMF_TRACK_RECORD or not records.is_contiguous():
        raise _lib.EmfError("records must be a contiguous float32 CUDA tensor of n_vol x 48")
    ws = trackWorkspace(points.device, n)
    m = (C.c_int * n)(*[int(x) for x in modes])
    pg = None
    if poseGrads is not None:
        pg = (C.c_void_p * n)(*[(_ptr(g) if g is not None else None) for g in poseGrads])
    check(_lib.lib().emf_track_linearise(n, _vol_array(vols), poses(rel_poses_CO), m, image(points),
                                         _f9(intr) if intr is not None else None, _opt_images(assoc, n),
                                         float(huberThresh), float(maxTSDFWeight), _opt_images(intWeights, n),
                                         _opt_images(tsdfVals, n), _opt_images(trackWeights, n), pg, _ptr(records),
                                         ws.data_ptr(), ws.numel(), _stream(stream)), "trackLinearise")
    if any(int(x) for x in modes):
        _count("trackLinearise")


def trackNormalisedWeights(intWeights, record, out, stream=None):
    check(_lib.lib().emf_track_normalised_weights(image(intWeights), _ptr(record), image(out), _stream(stream)),
          "trackNormalisedWeights")
    _count("trackNormalisedWeights")


class TrackPlan:
    """emf_track_linearise with every argument that does not change between the iterations of one frame marshalled
    once (volume table, image headers, workspace); per call only the poses and modes are written."""

    def __init__(self, vols, points, assoc, huberThresh, maxTSDFWeight, intWeights, records, tsdfVals=None,
                 trackWeights=None, intr=None):
        n = self.n = len(vols)
        if records.dtype != torch.float32 or records.numel() < n * _lib.EMF_TRACK_RECORD or not records.is_contiguous():
            raise _lib.EmfError("records must be a contiguous float32 CUDA tensor of n_vol x 48")
        self._keep = (vols, points, assoc, intWeights, records, tsdfVals, trackWeights)
        self._vols = _vol_array(vols)
        self._points = image(points)
        self._K = _f9(intr) if intr is not None else None
        self._assoc = _opt_images(assoc, n)
        self._iw = _opt_images(intWeights, n)
        self._vals = _opt_images(tsdfVals, n)
        self._tw = _opt_images(trackWeights, n)
        self._rec = _ptr(records)
        self._ws = trackWorkspace(points.device, n)
        self._poses = (Pose * n)()
        self._modes = (C.c_int * n)()
        self._huber, self._maxw = float(huberThresh), float(maxTSDFWeight)
        self._fn = _lib.lib().emf_track_linearise

    def launch(self, T_co: np.ndarray, modes, stream=None):
        """T_co: (n, 12) float32 (R row-major, then t), e.g. from poses.pack_poses; modes: n ints"""
        T_co = np.ascontiguousarray(T_co, dtype=np.float32)
        C.memmove(self._poses, T_co.ctypes.data, 48 * self.n)
        self._modes[:] = [int(m) for m in modes]
        check(self._fn(self.n, self._vols, self._poses, self._modes, self._points, self._K, self._assoc, self._huber,
                       self._maxw, self._iw, self._vals, self._tw, None, self._rec, self._ws.data_ptr(), self._ws.numel(),
                       _stream(stream)), "trackLinearise")
        if any(modes):
            _count("trackLinearise")


# emf_track_state (include/emf_b200.h) as a numpy record: what the host initialises and reads back
TRACK_STATE_DTYPE = np.dtype([("R", "<f8", (9,)), ("t", "<f8", (3,)), ("R_old", "<f8", (9,)), ("t_old", "<f8", (3,)),
                              ("mu", "<f8"), ("nu", "<f8"), ("rho", "<f8"), ("A", "<f4", (36,)), ("b", "<f4", (6,)),
                              ("x", "<f4", (6,)), ("err", "<f4"), ("err_new", "<f4"), ("converged", "<i4"),
                              ("first_iteration", "<i4"), ("evaluate_gradient", "<i4"), ("trial_pending", "<i4"),
                              ("iterations", "<i4"), ("linearisations", "<i4")], align=True)


class TrackLoopPlan:
    """emf_track_iterate with its arguments marshalled once: the device-resident Levenberg-Marquardt loop"""

    def __init__(self, vols, states, rel_poses_CO, points, intr, assoc, params, intWeights, records):
        n = self.n = len(vols)
        if states.dtype != torch.uint8 or states.numel() != n * TRACK_STATE_DTYPE.itemsize:
            raise _lib.EmfError("states must be a uint8 CUDA tensor of n_vol emf_track_state records")
        self._keep = (vols, states, points, assoc, intWeights, records)
        self._vols = _vol_array(vols)
        self._states = _ptr(states)
        self._hint = poses(rel_poses_CO)
        self._points = image(points)
        self._K = _f9(intr)
        self._assoc = images(assoc)
        self._iw = images(intWeights)
        self._rec = _ptr(records)
        self._ws = trackWorkspace(points.device, n)
        self._lm = _lib.TrackLMParams(params.tau, params.eps1, params.eps2, params.nu_init, params.huberThresh, params.maxTSDFWeight)
        self._fn = _lib.lib().emf_track_iterate

    def enqueue(self, n_iterations: int, stream=None):
        check(self._fn(self.n, self._vols, self._states, self._hint, self._points, self._K, self._assoc, C.byref(self._lm),
                       self._iw, self._rec, self._ws.data_ptr(), self._ws.numel(), int(n_iterations), _stream(stream)),
              "trackIterate")
        LAUNCHES["trackIterate"] = LAUNCHES.get("trackIterate", 0) + 2 * int(n_iterations)


def marchingCubes(vol, stream=None):
    """emf::TSDF::getMesh / emf::ObjTSDF::getMesh on the device (csrc/mcubes.cu): -> (vertices (n, 3) float32, normals (n, 3)
    float32, triangles (m, 4) int32 = VTK polygons [3, i0, i1, i2]) as CUDA tensors; one device->host read (the two counts)."""
    L = _lib.lib()
    res = (C.c_int * 3)(*[int(r) for r in vol.res])
    dev = torch.device("cuda", torch.cuda.current_device())
    ws = torch.empty(int(L.emf_mesh_workspace_bytes(res)), dtype=torch.uint8, device=dev)
    check(L.emf_mesh_count(C.byref(vol), ws.data_ptr(), ws.numel(), _stream(stream)), "meshCount")
    nv, nt = [int(x) for x in ws[:8].view(torch.int32).cpu().tolist()]
    if nv < 0 or nt < 0:
        raise _lib.EmfError("mesh larger than 2^31 - 1 elements")
    verts = torch.empty((nv, 3), dtype=torch.float32, device=dev)
    norms = torch.empty((nv, 3), dtype=torch.float32, device=dev)
    tris = torch.empty((nt // 4, 4), dtype=torch.int32, device=dev)
    if nv:
        check(L.emf_mesh_extract(C.byref(vol), ws.data_ptr(), ws.numel(), verts.data_ptr(), norms.data_ptr(), tris.data_ptr(),
                                 _stream(stream)), "meshExtract")
    _count("marchingCubes", 3 if nv else 2)
    return verts, norms, tris
