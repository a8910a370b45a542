"""Host side of the SDF tracker -- the Levenberg-Marquardt loop of emf::TSDF (reference src/core/TSDF.cpp:170-338)
as emf::EMFusion::performTracking drives it (src/core/EMFusion.cpp:672-722), for a whole list of volumes at once.

The reference runs, per volume and iteration, nine methods (computeGradients ... computePoseUpdate) that issue ~20
launches and block on three downloads.  Here the device part of an iteration of EVERY volume is one call of
emf_track_linearise (csrc/track.cu) and one device->host read of n_vol x 48 floats; the trial poses of
computePoseUpdate are a second call (error only) and a second read.  The control flow per volume -- damping, gain
ratio, accept / reject, the three convergence tests, what is recomputed when -- is the reference's, statement by statement.

Pose algebra: the reference keeps rel_pose_CO as a float Sophus::SE3f (unit quaternion + translation); here it is a
float64 (R, t) pair rounded to float32 once per launch, with Sophus' closed forms for exp / log restated below.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops
from .poses import Affine, pack_poses


# ---- SE(3) exponential / logarithm, tangent = (upsilon, omega) as in Sophus::SE3::exp / log --------------------
def _hat(w):
    return np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])


def so3_exp(omega):
    th = float(np.linalg.norm(omega))
    O = _hat(omega)
    if th < 1e-10:
        return np.eye(3) + O + 0.5 * (O @ O)
    return np.eye(3) + (np.sin(th) / th) * O + ((1.0 - np.cos(th)) / (th * th)) * (O @ O)


def so3_log(R):
    c = min(1.0, max(-1.0, (np.trace(R) - 1.0) / 2.0))
    th = float(np.arccos(c))
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    if th < 1e-10:
        return 0.5 * v
    if np.pi - th < 1e-6:   # near pi: axis from the symmetric part
        A = (R + np.eye(3)) / 2.0
        ax = np.sqrt(np.maximum(np.diag(A), 0.0))
        k = int(np.argmax(ax))
        ax = A[:, k] / ax[k]
        ax = ax / np.linalg.norm(ax)
        if np.dot(ax, v) < 0:
            ax = -ax
        return th * ax
    return (th / (2.0 * np.sin(th))) * v


def se3_exp(x) -> Affine:
    ups, om = np.asarray(x[:3], dtype=np.float64), np.asarray(x[3:], dtype=np.float64)
    th = float(np.linalg.norm(om))
    O = _hat(om)
    R = so3_exp(om)
    if th < 1e-10:
        V = np.eye(3) + 0.5 * O
    else:
        V = np.eye(3) + ((1.0 - np.cos(th)) / (th * th)) * O + ((th - np.sin(th)) / (th ** 3)) * (O @ O)
    return Affine(R, V @ ups)


def se3_log(T: Affine) -> np.ndarray:
    om = so3_log(T.R)
    th = float(np.linalg.norm(om))
    O = _hat(om)
    if th < 1e-10:
        Vi = np.eye(3) - 0.5 * O + (1.0 / 12.0) * (O @ O)
    else:
        half = 0.5 * th
        Vi = np.eye(3) - 0.5 * O + ((1.0 - th * np.cos(half) / (2.0 * np.sin(half))) / (th * th)) * (O @ O)
    return np.concatenate([Vi @ T.t, om])


def orthonormalise(T: Affine) -> Affine:
    """TSDF::prepareTracking (src/core/TSDF.cpp:174-181): Q of the Householder QR of the rotation block, columns
    flipped where the diagonal of R is negative."""
    Q, Rr = np.linalg.qr(T.R)
    for i in range(3):
        if Rr[i, i] < 0:
            Q[:, i] *= -1
    return Affine(Q, T.t)


class TrackState:
    """the tracking members of one emf::TSDF (include/EMFusion/core/TSDF.h:302-326)"""

    def __init__(self, vol, h: int, w: int, device, keep_images: bool):
        self.vol = vol
        self.rel_pose_CO = Affine.identity()
        self.mu = 0.0
        self.nu = 2.0
        self.rho = 0.0
        self.trackingConverged = False
        self.firstIteration = True
        self.evaluateGradient = True
        self.A = np.zeros((6, 6), dtype=np.float32)
        self.b = np.zeros(6, dtype=np.float32)
        self.x = np.zeros(6, dtype=np.float32)
        self.iterations = 0
        self.linearisations = 0
        self.intWeights = torch.zeros((h, w), dtype=torch.float32, device=device)   # before the NORM_INF scale
        self.tsdfVals = torch.zeros((h, w), dtype=torch.float32, device=device) if keep_images else None
        self.trackWeights = torch.zeros((h, w), dtype=torch.float32, device=device) if keep_images else None

    # TSDF::prepareTracking, src/core/TSDF.cpp:170-192
    def prepareTracking(self, cam_pose: Affine):
        self.rel_pose_CO = orthonormalise(self.vol.pose.inv() * cam_pose)
        self.nu = float(self.vol.params.nu_init)
        self.trackingConverged = False
        self.firstIteration = True
        self.evaluateGradient = True
        self.iterations = 0
        self.linearisations = 0


class Tracker:
    """performTracking for a list of volumes that are tracked together (the background alone, then all objects:
    src/core/EMFusion.cpp:673-721)."""

    def __init__(self, volumes: Sequence, frameSize, device, keep_images: bool = False, intr=None):
        """intr: the camera matrix the points are un-projected with (optional; lets the kernel skip the image tiles
        that cannot see a volume -- most of the frame for an object)."""
        w, h = frameSize
        self.intr = intr
        self.device = torch.device(device)
        self.states: List[TrackState] = [TrackState(v, h, w, self.device, keep_images) for v in volumes]
        n = len(self.states)
        self.records = torch.zeros((n, _lib.EMF_TRACK_RECORD), dtype=torch.float32, device=self.device)
        self._rec_host = torch.empty((n, _lib.EMF_TRACK_RECORD), dtype=torch.float32).pin_memory()
        self.keep_images = keep_images
        self.device_reads = 0

    def _plan(self, points, assoc):
        S = self.states
        cv = [s.vol.c_volume(with_grads=True) for s in S]   # the materialised float3 gradients when they are up to date
        return ops.TrackPlan(cv, points, assoc, S[0].vol.params.huberThresh, S[0].vol.params.maxTSDFWeight,
                             [s.intWeights for s in S], self.records,
                             tsdfVals=[s.tsdfVals for s in S] if self.keep_images else None,
                             trackWeights=[s.trackWeights for s in S] if self.keep_images else None, intr=self.intr)

    def _launch(self, plan, modes, poses):
        plan.launch(pack_poses(poses), modes)
        self._rec_host.copy_(self.records, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.device_reads += 1
        return self._rec_host.numpy()

    def track(self, points: torch.Tensor, assoc: Sequence[torch.Tensor], cam_pose: Affine, maxTrackingIter: int = 100):
        """prepareTracking + up to maxTrackingIter iterations for every volume; returns the states (rel_pose_CO updated).
        assoc[i]: association weights of volume i (W x H f32)."""
        S = self.states
        for s in S:
            s.prepareTracking(cam_pose)
        plan = self._plan(points, assoc)
        for _ in range(maxTrackingIter):
            if all(s.trackingConverged for s in S):
                break
            # ---- computeGradients ... reduceHessians (+ the err of computePoseUpdate) at the current poses
            modes = [0 if s.trackingConverged else (1 if s.evaluateGradient else 2) for s in S]
            rec = self._launch(plan, modes, [s.rel_pose_CO for s in S])
            trial_modes = [0] * len(S)
            trial_poses = [s.rel_pose_CO for s in S]
            errs = [0.0] * len(S)
            olds: List[Optional[Affine]] = [None] * len(S)
            for i, s in enumerate(S):
                if modes[i] == 0:
                    continue
                s.iterations += 1
                if modes[i] == 1:
                    s.linearisations += 1
                    s.A = rec[i, :36].reshape(6, 6).copy()
                    s.b = rec[i, 36:42].copy()
                    # reduceHessians, src/core/TSDF.cpp:278-282
                    if float(np.abs(s.b).max()) < s.vol.params.eps1:
                        s.trackingConverged = True
                        continue
                errs[i] = float(rec[i, 42])
                # ---- computePoseUpdate, src/core/TSDF.cpp:285-338 (up to the trial pose)
                if s.firstIteration:
                    s.mu = float(s.vol.params.tau) * float(np.diag(s.A).max())
                    s.firstIteration = False
                try:
                    s.x = np.linalg.solve(s.A + np.float32(s.mu) * np.eye(6, dtype=np.float32), s.b).astype(np.float32)
                except np.linalg.LinAlgError:   # cv::solve returns false and zeroes x: the step test below then ends the run
                    s.x = np.zeros(6, dtype=np.float32)
                if float(np.linalg.norm(s.x)) < s.vol.params.eps2 * (float(np.linalg.norm(se3_log(s.rel_pose_CO))) + s.vol.params.eps2):
                    s.trackingConverged = True
                    continue
                olds[i] = s.rel_pose_CO
                s.rel_pose_CO = se3_exp(-s.x.astype(np.float64)) * s.rel_pose_CO
                trial_modes[i] = 2
                trial_poses[i] = s.rel_pose_CO
            if not any(trial_modes):
                continue
            # ---- computeTSDFVals at the trial poses + computeError
            rec = self._launch(plan, trial_modes, trial_poses)
            for i, s in enumerate(S):
                if trial_modes[i] == 0:
                    continue
                err_new = float(rec[i, 42])
                x = s.x.astype(np.float32)
                gain = np.float32(0.5) * np.dot(-x, np.float32(s.mu) * -x - s.b)
                s.rho = (errs[i] - err_new) / float(gain) if gain != 0 else -1.0
                if s.rho > 0:
                    c = 2.0 * s.rho - 1.0
                    s.mu *= max(1.0 / 3.0, 1.0 - c * c * c)
                    s.nu = float(s.vol.params.nu_init)
                    s.evaluateGradient = True
                else:
                    s.rel_pose_CO = olds[i]
                    s.mu *= s.nu
                    s.nu *= float(s.vol.params.nu_init)
                    s.evaluateGradient = False
        return S

    def track_device(self, points: torch.Tensor, assoc: Sequence[torch.Tensor], cam_pose: Affine, maxTrackingIter: int = 100,
                     chunk: int = 4):
        """The same loop with its host part on the device (emf_track_iterate): iterations are enqueued `chunk` at a time and
        nothing is waited for in between -- a copy of the states follows every chunk and is looked at only once it has
        arrived (never blocking unless four chunks are in flight), to stop enqueuing when every volume has converged.
        Needs the camera matrix (Tracker(..., intr=K))."""
        if self.intr is None:
            raise _lib.EmfError("track_device needs the camera matrix: Tracker(..., intr=K)")
        S = self.states
        n = len(S)
        for s in S:
            s.prepareTracking(cam_pose)
        prm = S[0].vol.params
        st = np.zeros(n, dtype=ops.TRACK_STATE_DTYPE)
        for i, s in enumerate(S):
            st["R"][i] = s.rel_pose_CO.R.reshape(9); st["t"][i] = s.rel_pose_CO.t
            st["nu"][i] = float(prm.nu_init)
            st["first_iteration"][i] = 1; st["evaluate_gradient"][i] = 1
        nbytes = st.nbytes
        dev = torch.from_numpy(st.view(np.uint8).copy()).to(self.device)
        cv = [s.vol.c_volume(with_grads=True) for s in S]
        plan = ops.TrackLoopPlan(cv, dev, [s.rel_pose_CO for s in S], points, self.intr, assoc, prm,
                                 [s.intWeights for s in S], self.records)
        ring = [(torch.empty((nbytes,), dtype=torch.uint8).pin_memory(), torch.cuda.Event()) for _ in range(4)]
        pending = []          # indices into the ring, oldest first
        issued, k_ring, all_done = 0, 0, False

        def converged(buf) -> bool:
            return bool(np.all(buf.numpy().view(ops.TRACK_STATE_DTYPE)["converged"] != 0))
        while issued < maxTrackingIter and not all_done:
            k = min(chunk, maxTrackingIter - issued)
            plan.enqueue(k)
            issued += k
            if len(pending) == len(ring):                 # bound the run-ahead: wait for the oldest snapshot
                j = pending.pop(0)
                ring[j][1].synchronize()
                self.device_reads += 1
                all_done = converged(ring[j][0])
            j = k_ring % len(ring); k_ring += 1
            ring[j][0].copy_(dev, non_blocking=True)
            ring[j][1].record()
            pending.append(j)
            while pending and ring[pending[0]][1].query():   # snapshots that have arrived (no waiting)
                j0 = pending.pop(0)
                all_done = all_done or converged(ring[j0][0])
        final = torch.empty((nbytes,), dtype=torch.uint8).pin_memory()
        final.copy_(dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.device_reads += 1
        out = final.numpy().view(ops.TRACK_STATE_DTYPE)
        for i, s in enumerate(S):
            s.rel_pose_CO = Affine(out["R"][i].reshape(3, 3), out["t"][i])
            s.mu, s.nu, s.rho = float(out["mu"][i]), float(out["nu"][i]), float(out["rho"][i])
            s.A = out["A"][i].reshape(6, 6).copy(); s.b = out["b"][i].copy(); s.x = out["x"][i].copy()
            s.trackingConverged = bool(out["converged"][i]); s.evaluateGradient = bool(out["evaluate_gradient"][i])
            s.firstIteration = bool(out["first_iteration"][i])
            s.iterations, s.linearisations = int(out["iterations"][i]), int(out["linearisations"][i])
        self.iterations_enqueued = issued
        return S

    # TSDF::syncTrack (src/core/TSDF.cpp:333-338): cam_pose = pose * rel_pose_CO
    def syncTrackCamera(self, i: int = 0) -> Affine:
        s = self.states[i]
        return s.vol.pose * s.rel_pose_CO

    # ObjTSDF::syncTrack (src/core/ObjTSDF.cpp:228-235): pose = cam_pose * rel_pose_CO^-1
    def syncTrackObjects(self, cam_pose: Affine):
        for s in self.states:
            s.vol.pose = cam_pose * s.rel_pose_CO.inv()

    # TSDF::getTrackingWeights' image (src/core/TSDF.cpp:340-343): the normalised, combined weights
    def intWeightsNormalised(self, i: int) -> torch.Tensor:
        out = torch.empty_like(self.states[i].intWeights)
        ops.trackNormalisedWeights(self.states[i].intWeights, self.records[i], out)
        return out
