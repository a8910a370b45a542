"""Host-facing frame pipeline: depth frames arrive in pinned host memory, the composited model view (segmentation + ray
lengths) is returned in pinned host memory -- what emf::EMFusion::processFrame does between `depth_raw.upload`
(reference src/core/EMFusion.cpp:72) and the host-side consumers of the raycast (`getLastMasks`, the renderers,
src/core/EMFusion.cpp:131-200).

The reference uploads, processes and downloads strictly one after the other, blocking the host on each.  Here the three
are on three streams: the upload of frame n+1 and the download of frame n run under the kernels of the neighbouring frames;
the host only waits when it asks for a result.  Every frame's input still crosses PCIe exactly once and every frame's
result is read back exactly once.
"""
from __future__ import annotations

from typing import Optional

import torch

from .poses import Affine


class HostFramePipeline:
    DEPTH_SLOTS = 2

    def __init__(self, engine, download: bool = True):
        """download=False: a rank that does not hold the composite (multi-GPU, rank != 0) only uploads and computes"""
        self.eng = engine
        self.download = download
        # one GPU: upload / frame / download are queued by the library itself (include/emf_b200.h: emf_engine_submit_host);
        # several GPUs: the frame is several engine calls around the exchanges, the three streams are driven from here
        self._native = getattr(engine, "world", 1) == 1 and hasattr(engine, "submit_host")
        self._count = 0
        if self._native:
            w, h = engine.params.frameSize
            self.h2d_bytes_per_frame = h * w * 4
            self.d2h_bytes_per_frame = h * w * 5
            return
        dev = engine.device
        w, h = engine.params.frameSize
        self._main = torch.cuda.current_stream(dev)
        self._up = torch.cuda.Stream(dev)
        self._down = torch.cuda.Stream(dev)
        n = self.DEPTH_SLOTS
        self._depth = [torch.empty((h, w), dtype=torch.float32, device=dev) for _ in range(n)]
        self._seg_dev = [torch.empty((h, w), dtype=torch.uint8, device=dev) for _ in range(n)]
        self._ray_dev = [torch.empty((h, w), dtype=torch.float32, device=dev) for _ in range(n)]
        self.seg_host = [torch.empty((h, w), dtype=torch.uint8).pin_memory() for _ in range(n)]
        self.ray_host = [torch.empty((h, w), dtype=torch.float32).pin_memory() for _ in range(n)]
        self._uploaded = [torch.cuda.Event() for _ in range(n)]
        self._computed = [torch.cuda.Event() for _ in range(n)]
        self._downloaded = [torch.cuda.Event() for _ in range(n)]
        self._count = 0
        self.h2d_bytes_per_frame = h * w * 4
        self.d2h_bytes_per_frame = h * w * 5

    def submit(self, depth_host: torch.Tensor, cam_pose: Optional[Affine] = None, obj_poses: Optional[dict] = None) -> int:
        """Queue one frame (depth_host: pinned H x W float32).  Returns its ticket; never blocks the host unless the slot's
        previous result has not been read back yet."""
        if self._native:       # single GPU: the whole pipeline is one C-ABI call (emf_engine_submit_host)
            self._count += 1
            return self.eng.submit_host(depth_host, cam_pose, obj_poses, download=self.download)
        k = self._count
        s = k % self.DEPTH_SLOTS
        with torch.cuda.stream(self._up):
            if k >= self.DEPTH_SLOTS:
                self._up.wait_event(self._computed[s])         # the frame that last used this depth slot is done with it
            self._depth[s].copy_(depth_host, non_blocking=True)
            self._uploaded[s].record(self._up)
        self._main.wait_event(self._uploaded[s])
        if k >= self.DEPTH_SLOTS:
            self._main.wait_event(self._downloaded[s])         # ... and its staged result has left the device
        self.eng.processFrame(self._depth[s], cam_pose, obj_poses)
        # stage the composite on the compute stream (3 us) so that the next frame may overwrite the engine's images
        if self.download:
            self._seg_dev[s].copy_(self.eng.modelSegmentation, non_blocking=True)
            self._ray_dev[s].copy_(self.eng.raylengths, non_blocking=True)
        self._computed[s].record(self._main)
        with torch.cuda.stream(self._down):
            self._down.wait_event(self._computed[s])
            if self.download:
                self.seg_host[s].copy_(self._seg_dev[s], non_blocking=True)
                self.ray_host[s].copy_(self._ray_dev[s], non_blocking=True)
            self._downloaded[s].record(self._down)
        self._count += 1
        return k

    def result(self, ticket: int):
        """(segmentation, ray lengths) of a submitted frame in pinned host memory; blocks until they have arrived.  Valid
        until DEPTH_SLOTS further frames have been submitted."""
        if self._native:
            return self.eng.result_host(ticket)
        if ticket < self._count - self.DEPTH_SLOTS or ticket >= self._count:
            raise ValueError("result of this frame is no longer (or not yet) available")
        s = ticket % self.DEPTH_SLOTS
        self._downloaded[s].synchronize()
        return self.seg_host[s], self.ray_host[s]
