#!/usr/bin/env python
"""Tracker row: device time of ONE Levenberg-Marquardt iteration of every volume of config 4
(512^3 background + 32 objects @128^3, 640x480) -- emf_track_linearise (one launch + one 6 KB read) against the
reference's launch chain (oracle/_ref: computePoseGradients, getVolumeVals x 2, computeAb, multSingletonCol x 2 compiled
unchanged + the restated OpenCV-CUDA ops, per volume, with its blocking downloads).  Prints one JSON line.

  python scripts/bench_tracker.py [--config 4] [--iters 20]
Not the headline metric (bench.py); a measured number for the next row of the scope table."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=4)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--no-reference", action="store_true")
    args = ap.parse_args()
    import torch
    import bench
    from emfusion_b200 import ops
    from emfusion_b200.native import NativeEngine
    from emfusion_b200.poses import rel_pose_CO
    from emfusion_b200.synth import Scene
    from emfusion_b200.tracking import Tracker, se3_exp, se3_log
    from emfusion_b200.volume import ObjTSDF, Params

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    name, bg, k, ob, w, h = bench.CONFIGS[args.config]
    scene = Scene(n_objects=k, width=w, height=h, seed=0)
    prm = Params(frameSize=(w, h), intr=scene.K, globalVolumeDims=(bg,) * 3, globalVoxelSize=5.12 / bg, objVolumeDims=(ob,) * 3)
    ObjTSDF.nextID = 0
    eng = NativeEngine(prm, dev)
    for i in range(k):
        eng.add_object(scene.object_pose(i, 0), scene.object_voxel_size(i, ob))
    frames = [scene.render(f) for f in range(6)]
    d_dev = [torch.from_numpy(d).to(dev) for d, _ in frames]
    for f in range(5):   # a model to track against
        eng.processFrame(d_dev[f], scene.cam_pose(f), {o.id: scene.object_pose(o.id - 1, f) for o in eng.objects})
    f = 5
    eng.processFrame(d_dev[f], scene.cam_pose(f), {o.id: scene.object_pose(o.id - 1, f) for o in eng.objects})
    torch.cuda.synchronize()
    vols = eng.local_volumes()
    n = len(vols)
    points = eng.points
    assoc = eng._assoc_images(vols)
    cam = scene.cam_pose(f) * se3_exp(np.array([0.004, -0.003, 0.005, 0.003, -0.002, 0.004]))
    T = [rel_pose_CO(cam, v.pose) for v in vols]
    cv = [v.c_volume() for v in vols]
    iw = [torch.zeros((h, w), device=dev) for _ in vols]
    rec = torch.zeros((n, 48), device=dev)
    rec_host = torch.empty((n, 48)).pin_memory()

    from emfusion_b200.poses import pack_poses
    plan = ops.TrackPlan(cv, points, assoc, 0.2, 64.0, iw, rec, intr=scene.K)
    T_arr = pack_poses(T)

    def ours(mode):
        plan.launch(T_arr, [mode] * n)
        rec_host.copy_(rec, non_blocking=True)

    for _ in range(3):
        ours(1); ours(2)
    torch.cuda.synchronize()
    res = {}
    for mode, key in ((1, "linearise_ms"), (2, "error_only_ms")):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            ours(mode)
        e1.record()
        torch.cuda.synchronize()
        res[key] = e0.elapsed_time(e1) / args.iters
    # wall time of an iteration as the host loop sees it (launch + read + sync)
    t0 = time.perf_counter()
    for _ in range(args.iters):
        ours(1)
        torch.cuda.current_stream().synchronize()
    res["linearise_wall_ms"] = (time.perf_counter() - t0) * 1e3 / args.iters
    # the kernel alone: launches issued back to back without the read (host issue is then the only other cost)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        plan.launch(T_arr, [1] * n)
    e1.record()
    torch.cuda.synchronize()
    res["linearise_back_to_back_ms"] = e0.elapsed_time(e1) / 200

    # algorithmic bytes: 20 B per VISITED pixel and volume (point 12, association 4, combined weight 4) + 160 B per pixel
    # that gathers (tsdf 32 + weight 32 + float3 gradient 96, the reference's own footprint)
    vals = [torch.zeros((h, w), device=dev) for _ in vols]
    ops.trackLinearise(cv, T, [1] * n, points, assoc, 0.2, 64.0, iw, rec, tsdfVals=vals, intr=scene.K)
    torch.cuda.synchronize()
    inb = int(sum(int((v != 0).sum()) for v in vals))
    # pixels the launch really visits: the tile rectangles recorded by the kernel (records[:, 45..46])
    rb = rec.cpu().numpy().view(np.uint32)
    visited = int(sum(((int(a) >> 16) - (int(a) & 0xffff)) * ((int(b) >> 16) - (int(b) & 0xffff)) * 256 for a, b in zip(rb[:, 45], rb[:, 46])))
    alg = 20.0 * visited + 160.0 * inb
    peak, src = bench.peaks()
    res.update({"in_bounds_pixel_volume_pairs": inb, "visited_pixel_volume_pairs": visited, "algorithmic_bytes": alg,
                "roofline": {"bound": "hbm", "kernel": "k_track", "achieved": alg / (res["linearise_ms"] * 1e-3) / 1e9, "peak": peak,
                             "unit": "GB/s", "frac": alg / (res["linearise_ms"] * 1e-3) / 1e9 / peak, "peak_source": src,
                             "note": "gathers are L2-resident; the kernel is latency/issue bound, the fraction is reported for completeness"}})

    # whole LM loop on the background from a perturbed camera pose
    tr = Tracker([eng.background], (w, h), dev, intr=scene.K)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = tr.track(points, [eng.bg_associationWeights], cam, maxTrackingIter=100)[0]
    torch.cuda.synchronize()
    res["background_lm"] = {"wall_ms": (time.perf_counter() - t0) * 1e3, "iterations": st.iterations, "linearisations": st.linearisations,
                            "device_reads": tr.device_reads,
                            "pose_error_before": float(np.linalg.norm(se3_log(scene.cam_pose(f).inv() * cam))),
                            "pose_error_after": float(np.linalg.norm(se3_log(scene.cam_pose(f).inv() * tr.syncTrackCamera(0))))}

    # the same run with the host part of the loop on the device (emf_track_iterate), and all 33 volumes together
    tr2 = Tracker([eng.background], (w, h), dev, intr=scene.K)
    tr2.track_device(points, [eng.bg_associationWeights], cam, maxTrackingIter=100)      # warm-up (allocations)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st2 = tr2.track_device(points, [eng.bg_associationWeights], cam, maxTrackingIter=100)[0]
    torch.cuda.synchronize()
    res["background_lm_device_loop"] = {"wall_ms": (time.perf_counter() - t0) * 1e3, "iterations": st2.iterations,
                                        "linearisations": st2.linearisations, "iterations_enqueued": tr2.iterations_enqueued,
                                        "pose_error_after": float(np.linalg.norm(se3_log(scene.cam_pose(f).inv() * tr2.syncTrackCamera(0))))}
    tr3 = Tracker(vols, (w, h), dev, intr=scene.K)
    tr3.track_device(points, assoc, cam, maxTrackingIter=100)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st3 = tr3.track_device(points, assoc, cam, maxTrackingIter=100)
    torch.cuda.synchronize()
    res["all_volumes_lm_device_loop"] = {"wall_ms": (time.perf_counter() - t0) * 1e3, "iterations_enqueued": tr3.iterations_enqueued,
                                         "iterations_max": max(s.iterations for s in st3), "converged": sum(int(s.trackingConverged) for s in st3)}
    tr4 = Tracker(vols, (w, h), dev, intr=scene.K)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st4 = tr4.track(points, assoc, cam, maxTrackingIter=100)
    torch.cuda.synchronize()
    res["all_volumes_lm_host_loop"] = {"wall_ms": (time.perf_counter() - t0) * 1e3, "iterations_max": max(s.iterations for s in st4),
                                       "device_reads": tr4.device_reads}

    out = {"what": "one tracker iteration of every volume", "config": name, "n_volumes": n, "ours": res}
    if not args.no_reference:
        from tests import ref_gpu
        if ref_gpu.available():
            rt = ref_gpu.RefTracker(w, h)
            gv = []
            for v in vols:
                g = torch.zeros((v.numVoxels(), 3), device=dev)
                ref_gpu.update_gradients(v.tsdfVol, g, v.volumeRes)
                gv.append(g)
            torch.cuda.synchronize()

            def ref_iter():
                for v, g, a, t in zip(vols, gv, assoc, T):
                    rt.linearise(v.tsdfVol, g, v.tsdfWeights, points, a, t.rotation32(), t.translation32(), v.volumeRes, v.voxelSize)
            ref_iter()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                ref_iter()
            torch.cuda.synchronize()
            out["reference_chain"] = {"linearise_wall_ms": (time.perf_counter() - t0) * 1e3 / 3,
                                      "what": "reference kernels + restated OpenCV launches and blocking downloads, per volume "
                                              "(its own wall time: every volume's iteration ends in host barriers)"}
            out["speedup_wall"] = out["reference_chain"]["linearise_wall_ms"] / res["linearise_wall_ms"]
            rt.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
