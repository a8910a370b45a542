"""ctypes binding of the C ABI declared in include/emf_b200.h.

The product has exactly one compute path: libemf_b200.so (hand-written sm_100a CUDA).  If the
library is missing this module raises at import of the first op -- there is no CPU or PyTorch
fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# (EMF_B200_LIB: another build of the same library, for A/B experiments -- scripts/ab_build.sh)
LIB_PATH = os.environ.get("EMF_B200_LIB") or os.path.join(_HERE, "lib", "libemf_b200.so")

EMF_OK = 0
EMF_ERR_INVALID = -1
EMF_ERR_CUDA = -2
EMF_ERR_UNSUPPORTED = -3
EMF_MAX_VOLUMES = 96
EMF_TRACK_RECORD = 48


class EmfError(RuntimeError):
    pass


class Image(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("pitch", C.c_size_t), ("width", C.c_int), ("height", C.c_int)]


class Pose(C.Structure):
    _fields_ = [("R", C.c_float * 9), ("t", C.c_float * 3)]


class TsdfParams(C.Structure):
    _fields_ = [("max_tsdf_weight", C.c_float), ("assoc_sigma", C.c_float), ("alpha", C.c_float),
                ("uni_prior", C.c_float)]


class Volume(C.Structure):
    _fields_ = [("tsdf", C.c_void_p), ("weights", C.c_void_p), ("grads", C.c_void_p), ("fg_probs", C.c_void_p), ("fg_box", C.c_void_p),
                ("const_bits", C.c_void_p), ("brick_map", C.c_void_p), ("res", C.c_int * 3), ("voxel_size", C.c_float), ("truncdist", C.c_float), ("id", C.c_int)]


class TrackLMParams(C.Structure):
    _fields_ = [("tau", C.c_float), ("eps1", C.c_float), ("eps2", C.c_float), ("nu_init", C.c_float),
                ("huber_thresh", C.c_float), ("max_tsdf_weight", C.c_float)]


class EngineConfig(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("K", C.c_float * 9), ("params", TsdfParams),
                ("boundary", C.c_int), ("visibility_thresh", C.c_int)]


_P = C.POINTER
_SIGS = {
    "emf_fill_image_f32": [_P(Image), C.c_float, C.c_void_p],
    "emf_engine_set_volumes": [C.c_void_p, C.c_int, _P(Volume), C.c_int, C.c_void_p],
    "emf_engine_frame": [C.c_void_p, _P(Image), _P(Pose), _P(Pose), C.c_uint, C.c_void_p],
    "emf_engine_stage_ms": [C.c_void_p, _P(C.c_float)],
    "emf_engine_image": [C.c_void_p, C.c_int, C.c_int, _P(Image)],
    "emf_engine_vis_counts": [C.c_void_p, _P(C.c_int32), C.c_int],
    "emf_engine_force_integrate": [C.c_void_p, C.c_int],
    "emf_engine_set_background_rows": [C.c_void_p, C.c_int, C.c_int],
    "emf_compute_points": [_P(Image), _P(Image), _P(C.c_float), C.c_void_p],
    "emf_update_tsdf": [_P(Image), _P(Image), C.c_void_p, C.c_void_p, _P(Pose), _P(C.c_float), _P(C.c_int),
                        C.c_float, C.c_float, C.c_float, C.c_void_p],
    "emf_compute_tsdf_grads": [C.c_void_p, C.c_void_p, _P(C.c_int), C.c_void_p],
    "emf_raycast_tsdf": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _P(Image), _P(Image), _P(Image), _P(Image),
                         _P(Pose), _P(C.c_float), _P(C.c_int), C.c_float, C.c_float, C.c_void_p, C.c_void_p],
    "emf_get_volume_vals": [C.c_void_p, _P(Image), _P(Pose), _P(C.c_int), C.c_float, _P(Image), C.c_void_p],
    "emf_update_fgbg_probs": [_P(Image), _P(Image), C.c_void_p, C.c_void_p, C.c_void_p, _P(Pose), _P(C.c_float),
                              _P(C.c_int), C.c_float, C.c_void_p],
    "emf_compute_fg_probs": [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p],
    "emf_compute_fg_probs_box": [C.c_void_p, _P(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "emf_compute_association": [_P(Volume), _P(Image), _P(Pose), _P(TsdfParams), _P(Image), _P(Image), C.c_void_p],
    "emf_assoc_weights": [C.c_int, _P(Volume), _P(Pose), _P(Image), _P(TsdfParams), _P(Image), C.c_int, _P(Image),
                          C.c_void_p],
    "emf_assoc_normalise": [C.c_int, _P(Image), _P(Image), C.c_void_p],
    "emf_raycast_volumes": [C.c_int, _P(Volume), _P(Pose), _P(C.c_float), _P(C.c_int), _P(Image), _P(Image),
                            _P(Image), _P(Image), C.c_void_p, C.c_void_p],
    "emf_raycast_volumes_ws": [C.c_int, _P(Volume), _P(Pose), _P(C.c_float), _P(C.c_int), _P(Image), _P(Image),
                               _P(Image), _P(Image), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p],
    "emf_raycast_volumes_opt": [C.c_int, _P(Volume), _P(Pose), _P(C.c_float), _P(C.c_int), _P(Image), _P(Image),
                                _P(Image), _P(Image), C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p],
    "emf_raycast_schedule_update": [C.c_int, C.c_int, _P(C.c_int), C.c_void_p, C.c_size_t, C.c_void_p],
    "emf_raycast_composite": [C.c_int, _P(C.c_int), _P(C.c_int), _P(Image), _P(Image), _P(Image), _P(Image),
                              _P(Image), _P(Image), _P(Image), _P(Image), C.c_int, _P(Image), _P(Image), _P(Image),
                              _P(Image), C.c_void_p, C.c_void_p],
    "emf_composite_merge": [C.c_int, _P(Image), _P(Image), _P(Image), _P(Image), C.c_int, _P(C.c_int), _P(Image), _P(Image),
                            _P(Image), _P(Image), C.c_int, _P(Image), _P(Image), _P(Image), _P(Image), C.c_void_p, C.c_int,
                            _P(C.c_void_p), _P(C.c_void_p), _P(C.c_void_p), _P(C.c_void_p), C.c_void_p],
    "emf_integrate_volumes": [C.c_int, _P(Volume), _P(Pose), _P(C.c_float), _P(Image), _P(Image), C.c_float,
                              C.c_void_p],
    "emf_integrate_volumes_gated": [C.c_int, _P(Volume), _P(Pose), _P(C.c_float), _P(Image), _P(Image), C.c_float,
                                    C.c_void_p, _P(C.c_int), C.c_int, C.c_void_p, C.c_void_p],
    "emf_integrate_volumes_ws": [C.c_int, _P(Volume), _P(Pose), _P(C.c_float), _P(Image), _P(Image), C.c_float,
                                 C.c_void_p, _P(C.c_int), C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p],
    "emf_integrate_volumes_phase": [C.c_int, _P(Volume), _P(Pose), _P(C.c_float), _P(Image), _P(Image), C.c_float,
                                    C.c_void_p, _P(C.c_int), C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p],
    "emf_mesh_count": [_P(Volume), C.c_void_p, C.c_size_t, C.c_void_p],
    "emf_mesh_extract": [_P(Volume), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "emf_update_brick_maps": [C.c_int, _P(Volume), C.c_void_p],
    "emf_reset_bitmaps": [_P(Volume), C.c_void_p],
    "emf_preprocess_depth": [_P(Image), _P(Image), _P(Image), _P(C.c_float), C.c_int, C.c_float, C.c_float, C.c_void_p],
    "emf_copy_values": [C.c_void_p, C.c_void_p, C.c_int, _P(C.c_int), _P(C.c_int), _P(C.c_int), C.c_void_p],
    "emf_resize_volume": [C.c_void_p, C.c_void_p, C.c_void_p, _P(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p, _P(C.c_int),
                          _P(C.c_int), C.c_void_p],
    "emf_track_workspace_init": [C.c_void_p, C.c_size_t, C.c_void_p],
    "emf_track_linearise": [C.c_int, _P(Volume), _P(Pose), _P(C.c_int), _P(Image), _P(C.c_float), _P(Image), C.c_float, C.c_float,
                            _P(Image), _P(Image), _P(Image), _P(C.c_void_p), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p],
    "emf_track_iterate": [C.c_int, _P(Volume), C.c_void_p, _P(Pose), _P(Image), _P(C.c_float), _P(Image), _P(TrackLMParams),
                          _P(Image), C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p],
    "emf_track_normalised_weights": [_P(Image), C.c_void_p, _P(Image), C.c_void_p],
    "emf_assoc_normalise_parts": [C.c_int, _P(Image), C.c_int, _P(C.c_void_p), _P(Image), C.c_void_p, C.c_uint32, C.c_void_p,
                                  C.c_double, C.c_void_p],
    "emf_engine_set_partial_norm_target": [C.c_void_p, _P(Image)],
    "emf_engine_submit_host": [C.c_void_p, C.c_void_p, _P(Pose), _P(Pose), C.c_uint, C.c_int, C.c_void_p, _P(C.c_longlong)],
    "emf_engine_result_host": [C.c_void_p, C.c_longlong, _P(C.c_void_p), _P(C.c_void_p)],
    "emf_engine_depth_slot": [C.c_void_p, C.c_longlong, _P(Image)],
    "emf_engine_set_composite_target": [C.c_void_p, _P(Image)],
    "emf_engine_set_background_target": [C.c_void_p, _P(Image)],
    "emf_engine_set_option": [C.c_void_p, C.c_int, C.c_int],
    "emf_engine_set_gate_source": [C.c_void_p, C.c_void_p, _P(C.c_int), C.c_int],
    "emf_engine_normalise_from_parts": [C.c_void_p, C.c_int, _P(C.c_void_p), C.c_void_p, C.c_uint32, C.c_void_p, C.c_double,
                                        C.c_void_p],
    "emf_xchg_alloc": [C.c_size_t, _P(C.c_void_p), C.c_char_p],
    "emf_xchg_open": [C.c_char_p, _P(C.c_void_p)],
    "emf_xchg_close": [C.c_void_p],
    "emf_xchg_free": [C.c_void_p],
    "emf_xchg_signal": [C.c_int, _P(C.c_void_p), C.c_uint32, C.c_void_p],
    "emf_xchg_wait": [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_double, C.c_void_p],
    "emf_xchg_sum_images": [C.c_int, _P(C.c_void_p), _P(Image), C.c_void_p],
    "emf_xchg_scatter_u32": [C.c_void_p, C.c_int, C.c_int, _P(C.c_void_p), C.c_void_p],
    "emf_volume_screen_rect": [_P(C.c_int), C.c_float, _P(Pose), _P(C.c_float), C.c_int, C.c_int, _P(C.c_int)],
}
EXPORTED = sorted(list(_SIGS) + ["emf_version", "emf_brick_map_bytes", "emf_track_workspace_bytes", "emf_integrate_workspace_bytes", "emf_raycast_workspace_bytes", "emf_mesh_workspace_bytes", "emf_engine_create", "emf_engine_destroy",
                                "emf_engine_vis_counts_device"])

_lib = None


def lib() -> C.CDLL:
    """Load libemf_b200.so (once).  Raises EmfError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EmfError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(run `python -m emfusion_b200.build` or __graft_entry__.build()); "
                "emfusion_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, args in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int
        L.emf_integrate_workspace_bytes.argtypes = [C.c_int, C.c_int]
        L.emf_integrate_workspace_bytes.restype = C.c_size_t
        L.emf_raycast_workspace_bytes.argtypes = [C.c_int, C.c_int]
        L.emf_raycast_workspace_bytes.restype = C.c_size_t
        L.emf_mesh_workspace_bytes.argtypes = [_P(C.c_int)]
        L.emf_mesh_workspace_bytes.restype = C.c_size_t
        L.emf_track_workspace_bytes.argtypes = [C.c_int]
        L.emf_track_workspace_bytes.restype = C.c_size_t
        L.emf_brick_map_bytes.argtypes = [_P(C.c_int)]
        L.emf_brick_map_bytes.restype = C.c_size_t
        L.emf_engine_create.argtypes = [_P(EngineConfig)]
        L.emf_engine_create.restype = C.c_void_p
        L.emf_engine_destroy.argtypes = [C.c_void_p]
        L.emf_engine_destroy.restype = None
        L.emf_engine_vis_counts_device.argtypes = [C.c_void_p]
        L.emf_engine_vis_counts_device.restype = C.c_void_p
        L.emf_version.restype = C.c_char_p
        L.emf_version.argtypes = []
        _lib = L
    return _lib


_ERR = {EMF_ERR_INVALID: "invalid argument", EMF_ERR_CUDA: "CUDA error", EMF_ERR_UNSUPPORTED: "unsupported"}


def check(rc: int, what: str) -> None:
    if rc != EMF_OK:
        raise EmfError(f"{what} failed: {_ERR.get(rc, rc)}")


def version() -> str:
    return lib().emf_version().decode()
