"""emfusion_b200 -- B200-native multi-volume TSDF engine behind EM-Fusion's TSDF/ObjTSDF surface.

Only the dense hot path of the reference (integrate, raycast + composite, EM association) lives
here; see DESIGN.md.  Compute is hand-written CUDA for sm_100a in csrc/, reached through the
C ABI in include/emf_b200.h; Python is the host mirror of the reference's class surface plus
torch (device memory, streams, torch.distributed) plumbing.
"""
from ._lib import EmfError, version, LIB_PATH  # noqa: F401

__all__ = ["EmfError", "version", "LIB_PATH"]
