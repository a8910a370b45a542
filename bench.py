#!/usr/bin/env python
"""bench.py -- Mvoxels/s of the EM-Fusion dense hot path (association + raycast/composite + integrate).

Metric (BASELINE.json / SURVEY.md section 8d): Mvoxels/s = sum over volumes of Rx*Ry*Rz, times frames,
divided by the device time of {one association pass + raycast incl. composite + integrate}, / 1e6.
Workload at every N: config 4 of BASELINE.json -- 512^3 background + 32 objects @128^3, synthetic
640x480 stream (a "step" = one frame).  N > 1 shards the objects across ranks (background on rank 0,
one all-reduce for the association normaliser, one gather for the composite): strong scaling.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1..5]

`--impl reference` times the reference's own CUDA kernels (oracle/_ref/libemf_ref.so: src/core/cuda/*.cu
compiled unchanged against a type shim) driven with the reference's launch structure, on ONE B200, same
workload and metric.  (The reference has no CPU implementation of this path -- it is CUDA-only; the scalar
CPU restatement is reported separately as `cpu_baseline`.)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (bg_res, n_obj, obj_res, width, height)
    1: ("cfg1 64^3 bg, 0 obj, 640x480", 64, 0, 64, 640, 480),
    2: ("cfg2 512^3 bg, 0 obj, 640x480", 512, 0, 64, 640, 480),
    3: ("cfg3 512^3 bg + 8 obj @64^3, 640x480", 512, 8, 64, 640, 480),
    4: ("cfg4 512^3 bg + 32 obj @128^3, 640x480", 512, 32, 128, 640, 480),
    5: ("cfg5 1024^3 bg + 64 obj @128^3, 1280x960", 1024, 64, 128, 1280, 960),
}
N_STREAM_FRAMES = 12   # distinct synthetic frames, cycled


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "unchanged-caller"],
                    help="unchanged-caller: the reference's launch structure (what an unmodified src/core/EMFusion.cpp does) over this "
                         "repo's level-1 operators -- the operator-level drop-in of INTEGRATION.md section 1")
    ap.add_argument("--config", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--brick-maps", action="store_true",
                    help="maintain constant-brick maps and let the raycast jump through them (bit-identical results)")
    ap.add_argument("--background-on-rank0", action="store_true",
                    help="multi-GPU: keep the background on GPU 0 only (BASELINE.json's layout) instead of replicating it "
                         "and sharding its raycast by image rows")
    ap.add_argument("--nccl-exchange", action="store_true",
                    help="multi-GPU: NCCL collectives (all-reduce, gather, broadcast) instead of the NVLink peer-memory exchange")
    ap.add_argument("--ray-certificate", type=int, default=-1,
                    help="1 / 0: ray-space certificate + four-lanes-per-ray march for the background's raycast on / off "
                         "(default: off -- measured slower at every N on the bench scene, DESIGN.md section 4.3)")
    ap.add_argument("--materialize-grads", action="store_true",
                    help="also materialise the float3 gradient volumes every frame (reference behaviour)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line): NVML polled every
    millisecond from a thread (the timed region is tens of milliseconds -- `nvidia-smi -lms` does not get a sample in);
    nvidia-smi is the fall-back when the NVML binding is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index=0):
        self.p = None
        self.lines = []
        self.sm, self.bits, self.mx = [], 0, None
        self.run = True
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = None
            try:        # the device this process computes on, by UUID (CUDA_VISIBLE_DEVICES may renumber or name devices by UUID)
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
                try:
                    self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
                except TypeError:
                    self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = None
            if self.h is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
                self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _poll(self):
        n = self.nvml
        while self.run:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                self.bits |= int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                pass
            time.sleep(0.001)

    def _read(self):
        for ln in self.p.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.nvml is not None:
            self.run = False
            self.t.join(timeout=1.0)
            reasons = sorted(name for bit, name in self.REASONS.items() if self.bits & bit)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx, "reasons": reasons,
                    "samples": len(self.sm), "source": "nvml"}
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def render_stream(scene, n):
    return [scene.render(f) for f in range(n)]


def total_voxels(cfg):
    _, bg, k, ob, _, _ = CONFIGS[cfg]
    return bg ** 3 + k * ob ** 3


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def cpu_baseline(cfg, seed=0, budget_s=20.0):
    """Scalar C oracle (OpenMP over z-slabs / image rows, all host cores) on a bounded sample of the same workload: the
    configuration's own volumes and frame size, as many stream frames as fit in ~budget_s seconds (at least one)."""
    from tests import oracle_c
    from tests import scenario as S
    from emfusion_b200.poses import rel_pose_CO, rel_pose_OC
    o = oracle_c.load()
    name, bg, k, ob, w, h = CONFIGS[cfg]
    k_s = k                                            # the full configuration: every object volume
    sc = S.make("cpu", o, w, h, (bg,) * 3, k_s, (ob,) * 3, n_frames=8, integrate_frames=1, seed=seed)
    nvox = bg ** 3 + k_s * ob ** 3
    t_assoc = t_ray = t_int = 0.0
    frames = 0
    t_start = time.time()
    for f in range(1, 8):
        cam = sc.cam(f)
        t0 = time.time()
        pts = o.compute_points(sc.depths[f], sc.K)
        imgs = []
        for v in sc.vols():
            T = rel_pose_CO(cam, v.pose)
            a, _ = o.assoc_volume(v.tsdf, v.fg_probs, pts, S.R9(T), S.T3(T), v.res, v.voxel, v.trunc)
            imgs.append(a)
        o.normalise(imgs)
        t_assoc += time.time() - t0
        t0 = time.time()
        rc = []
        for v in sc.vols():
            T = rel_pose_CO(cam, v.pose)
            g = o.compute_grads(v.tsdf, v.res)
            wts = v.weights if v.fg_probs is None else o.raycast_weights(v.weights, (v.fg_probs > 0.5).astype(np.uint8) * 255)
            rc.append(o.raycast(v.tsdf, g, wts, S.R9(T), S.T3(T), sc.K, v.res, v.voxel, v.trunc, w, h))
            del g
        if k_s:
            o.composite([v.vid for v in sc.objs], [r["ray"] for r in rc[1:]], [r["vert"] for r in rc[1:]],
                        [r["norm"] for r in rc[1:]], [r["mask"] for r in rc[1:]], rc[0]["ray"], rc[0]["vert"],
                        rc[0]["norm"], rc[0]["mask"], 20)
        t_ray += time.time() - t0
        t0 = time.time()
        for v, a in zip(sc.vols(), imgs):
            T = rel_pose_OC(cam, v.pose)
            o.update_tsdf(sc.depths[f], a, v.tsdf, v.weights, S.R9(T), S.T3(T), sc.K, v.res, v.voxel, v.trunc, 64.0)
        t_int += time.time() - t0
        frames += 1
        if time.time() - t_start > budget_s:
            break
    tot = t_assoc + t_ray + t_int
    return {"value": nvox * frames / tot / 1e6, "unit": "Mvoxels/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{frames} frame(s) {w}x{h} of the full configuration: bg {bg}^3 + {k_s} obj @{ob}^3 ({nvox} voxels/frame); assoc {t_assoc:.2f}s "
                      f"raycast+gradients {t_ray:.2f}s integrate {t_int:.2f}s (scalar C oracle, OpenMP, all host cores)"}


# ------------------------------------------------------------------------------------------------
def parity_vs_one_gpu(args, scene, frames, cams, prm, dev, rank, world, n_frames=3):
    """Untimed: the first frames of the stream through a fresh N-rank engine and, on rank 0 alone, through a fresh one-GPU
    engine; rank 0 compares the merged frame with the single-GPU one -- segmentation, ray lengths, visibility bit for bit,
    association within 1e-5 (the normaliser's summation order differs), the background volume within 1e-4."""
    import torch
    import torch.distributed as dist
    from emfusion_b200.native import NativeEngine
    from emfusion_b200.volume import ObjTSDF
    name, bg, k, ob, w, h = CONFIGS[args.config]

    def run(world_size, rk):
        ObjTSDF.nextID = 0
        e = NativeEngine(prm, dev, rank=rk, world_size=world_size,
                         replicate_background=False if args.background_on_rank0 else None,
                         peer_exchange=False if args.nccl_exchange else None)
        for i in range(k):
            e.add_object(scene.object_pose(i, 0), scene.object_voxel_size(i, ob))
        out = []
        for f in range(n_frames + 1):
            d = torch.from_numpy(frames[f][0]).to(dev)
            e.processFrame(d, cams[f], {o.id: scene.object_pose(o.id - 1, f) for o in e.objects})
            if f == 0:
                zeros = torch.zeros((h, w), dtype=torch.uint8, device=dev)
                inst0 = torch.from_numpy(frames[0][1]).to(dev)
                for o in e.objects:
                    o.integrateMask((inst0 == o.id).to(torch.uint8), zeros, e.pose, prm.intr)
            elif rk == 0:
                out.append((e.modelSegmentation.clone(), e.raylengths.clone(), e.bg_associationWeights.clone(), sorted(e.vis_objs)))
            else:
                _ = e.vis_objs
        torch.cuda.synchronize()
        return e, out

    eN, outN = run(world, rank)
    res = None
    if rank == 0:
        e1, out1 = run(1, 0)
        # frame 1 raycasts volumes that were integrated with association == 1 everywhere: bit for bit.  Later frames see
        # volumes integrated with association weights that differ in the last bits (the normaliser is summed rank by rank
        # instead of object by object): within BASELINE.json's 1e-4
        seg1 = bool((outN[0][0] == out1[0][0]).all())
        ray1 = bool((outN[0][1].view(torch.int32) == out1[0][1].view(torch.int32)).all())
        seg_frac = min(float((a[0] == b[0]).float().mean()) for a, b in zip(outN, out1))
        ray_linf = 0.0
        for a, b in zip(outN, out1):
            both = (a[0] == b[0]) & (a[1] > 0) & (b[1] > 0)
            if bool(both.any()):
                ray_linf = max(ray_linf, float((a[1] - b[1]).abs()[both].max()))
        vis_ok = all(a[3] == b[3] for a, b in zip(outN, out1))
        assoc = max(float((a[2] - b[2]).abs().max()) for a, b in zip(outN, out1))
        vol = float((eN.background.tsdfVol - e1.background.tsdfVol).abs().max()) if eN.background is not None else None
        res = {"frames": n_frames, "frame1_segmentation_bit_exact": seg1, "frame1_raylengths_bit_exact": ray1,
               "segmentation_equal_fraction_min": seg_frac, "raylengths_linf": ray_linf, "visibility_equal": vis_ok,
               "association_linf": assoc, "background_tsdf_linf": vol,
               "ok": bool(seg1 and ray1 and seg_frac > 0.9995 and ray_linf <= 1e-4 and vis_ok and assoc <= 1e-5 and (vol is None or vol <= 1e-4))}
        del e1
    del eN
    torch.cuda.synchronize()
    dist.barrier()
    return res


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from emfusion_b200 import ops
    from emfusion_b200.native import NativeEngine
    from emfusion_b200.poses import rel_pose_OC
    from emfusion_b200.synth import Scene
    from emfusion_b200.volume import ObjTSDF, Params

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    name, bg, k, ob, w, h = CONFIGS[args.config]
    scene = Scene(n_objects=k, width=w, height=h, seed=0)
    prm = Params(frameSize=(w, h), intr=scene.K, globalVolumeDims=(bg,) * 3, globalVoxelSize=5.12 / bg,
                 objVolumeDims=(ob,) * 3)
    ObjTSDF.nextID = 0
    eng = NativeEngine(prm, dev, rank=rank, world_size=world, materialize_grads=args.materialize_grads,
                       accelerate=args.brick_maps, replicate_background=False if args.background_on_rank0 else None,
                       peer_exchange=False if args.nccl_exchange else None)
    use_cert = args.ray_certificate if args.ray_certificate >= 0 else 0
    eng.set_ray_certificate(bool(use_cert))
    for i in range(k):
        eng.add_object(scene.object_pose(i, 0), scene.object_voxel_size(i, ob))
    frames = render_stream(scene, N_STREAM_FRAMES)
    d_dev = [torch.from_numpy(d).to(dev) for d, _ in frames]
    d_pin = [torch.from_numpy(d).pin_memory() for d, _ in frames]
    cams = [scene.cam_pose(f) for f in range(N_STREAM_FRAMES)]
    oposes = [{o.id: scene.object_pose(o.id - 1, f) for o in eng.objects} for f in range(N_STREAM_FRAMES)]

    # ---- set-up (untimed): frame 0 with association == 1, analytic masks -> fg probabilities
    eng.processFrame(d_dev[0], cams[0], oposes[0])
    zeros = torch.zeros((h, w), dtype=torch.uint8, device=dev)
    inst0 = torch.from_numpy(frames[0][1]).to(dev)
    for o in eng.objects:
        o.integrateMask((inst0 == o.id).to(torch.uint8), zeros, eng.pose, prm.intr)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step(f, depth=None, timed=False):
        """one frame through the public API: points, association, raycast + composite, integrate"""
        i = f % N_STREAM_FRAMES
        eng.processFrame(d_dev[i] if depth is None else depth, cams[i], oposes[i], timed=timed)

    f = 1
    for _ in range(max(args.warmup, 3)):
        step(f); f += 1
    # ---- timed: EXACTLY K frames back to back, device time between two events, max over ranks
    sampler = ClockSampler(local) if rank == 0 else None
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ops.launches_total()
    barrier()
    start.record()
    t_host = time.perf_counter()
    for s in range(args.steps):
        step(f); f += 1
    t_host = (time.perf_counter() - t_host) * 1e3 / args.steps     # host time to issue one frame (not a device time)
    stop.record()
    barrier()
    launches = ops.launches_total() - l0
    clocks = sampler.stop() if sampler else None
    ms_step = start.elapsed_time(stop) / args.steps

    # ---- stage breakdown (separate pass, one device sync per frame to read the engine's stage events)
    n_stage = min(args.steps, 10)
    stage = np.zeros((n_stage, 3))
    if world == 1:
        for s in range(n_stage):
            step(f, timed=True); f += 1
            stage[s] = eng.stage_ms()
        ms_stage = stage.mean(0)
    else:               # back to back (a sync per frame would put the ranks' start skew into the first phase), skipping the first two
        barrier()
        for s in range(2):
            step(f); f += 1
        for s in range(n_stage):
            step(f, timed=True); f += 1
        ms_stage = np.array(eng.stage_ms())
    stage_detail = None
    if world > 1:       # per phase, the slowest rank (the phases of different ranks overlap: they need not add up to the frame)
        ts = torch.tensor(ms_stage, dtype=torch.float64, device=dev)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        ms_stage = ts.cpu().numpy()
        det = getattr(eng, "stage_detail", None)
        if det:         # every rank's own phases (rank 0 merges, the others wait for it)
            keys = list(det)
            mine = torch.tensor([det[k] for k in keys], dtype=torch.float64, device=dev)
            allr = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            stage_detail = {"phases": keys, "per_rank_ms": [[round(float(x), 4) for x in t.cpu().numpy()] for t in allr]}

    # ---- e2e: host depth in (pinned), composited result out (pinned), every step, through the public host-facing API
    #      (HostFramePipeline: upload of frame n+1 and download of frame n overlap the kernels; every frame's depth
    #      crosses PCIe once and every frame's result is read back once, all inside the timed region)
    from emfusion_b200.pipeline import HostFramePipeline
    pipe = HostFramePipeline(eng, download=(rank == 0))
    prev = None
    for s in range(max(args.warmup, 3)):     # W untimed frames through the same entry (first use of its slots, streams and caches)
        i = f % N_STREAM_FRAMES
        tk = pipe.submit(d_pin[i], cams[i], oposes[i]); f += 1
        if prev is not None:
            pipe.result(prev)
        prev = tk
    pipe.result(prev)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    prev = None
    for s in range(args.steps):
        i = f % N_STREAM_FRAMES
        tk = pipe.submit(d_pin[i], cams[i], oposes[i]); f += 1
        if prev is not None:
            pipe.result(prev)          # the host consumes frame n-1 while frame n is on the device
        prev = tk
    pipe.result(prev)
    e1.record()
    barrier()
    if os.environ.get("EMF_BENCH_E2E_DIAG") == "1" and world == 1:      # diagnostics (stderr): where the e2e loop's time goes
        def loop(p, wait):
            nonlocal f
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            pv = None
            for _ in range(args.steps):
                i = f % N_STREAM_FRAMES
                k = p.submit(d_pin[i], cams[i], oposes[i]); f += 1
                if wait and pv is not None:
                    p.result(pv)
                pv = k
            p.result(pv)
            b.record()
            barrier()
            return a.elapsed_time(b) / args.steps
        print("e2e diag: e2e again %.4f, host never waits inside the loop %.4f, no download %.4f, no download + no wait %.4f ms" % (
            loop(pipe, True), loop(pipe, False), loop(HostFramePipeline(eng, download=False), True),
            loop(HostFramePipeline(eng, download=False), False)), file=sys.stderr)
    t = torch.tensor([ms_step, e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, ms_e2e = float(t[0]), float(t[1])

    parity = parity_vs_one_gpu(args, scene, frames, cams, prm, dev, rank, world) if world > 1 else None
    nvox = total_voxels(args.config)
    if rank == 0:
        peak, peak_src = peaks()
        roof = {"bound": "hbm", "kernel": "k_integrate_bricks", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None,
                "traffic": None, "peak_source": peak_src}
        ray_roof = None
        try:   # per-launch DRAM bytes / warp instructions of the same kernels on the same workload, from the committed ncu
               # capture -- only if it was taken from the kernel sources as they are now (stamped with their git blob hashes)
            with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as fh:
                tr_all = json.load(fh)
            cur = {f: subprocess.run(["git", "hash-object", os.path.join(ROOT, "emfusion_b200", "csrc", f)], capture_output=True,
                                     text=True).stdout.strip() for f in ("integrate.cu", "raycast.cu")}
            if args.config == 4 and world == 1:
                tr = tr_all.get("k_integrate_bricks")
                if tr and tr_all.get("source_hashes", {}).get("integrate.cu") == cur["integrate.cu"]:
                    roof["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
                    roof["traffic_source"] = tr_all.get("source")
                else:
                    roof["traffic_source"] = "stale: profiles/r2_traffic.json was captured from another integrate.cu"
                rr = tr_all.get("k_raycast")
                if rr and tr_all.get("source_hashes", {}).get("raycast.cu") == cur["raycast.cu"]:
                    ray_roof = rr
        except Exception:
            pass
        if world == 1:
            # algorithmic bytes of the integrate launch: 16 B/voxel x the voxels of the volumes it integrated (upper-bound
            # convention); next to it the exact bytes from the kernel's own counters (separate untimed launch)
            vis = eng.vis_objs
            vols = [v for v in eng.local_volumes() if v.id == 0 or v.id in vis]
            int_vox = sum(v.numVoxels() for v in vols)
            st = torch.zeros(8, dtype=torch.int64, device=dev)
            i = f % N_STREAM_FRAMES
            scratch = [(v.tsdfVol.clone(), v.tsdfWeights.clone()) for v in vols]
            cv = [ops.volume(t_, w_, v.volumeRes, v.voxelSize, v.truncdist, vid=v.id) for v, (t_, w_) in zip(vols, scratch)]
            ops.integrateVolumes(cv, [rel_pose_OC(cams[i], v.pose) for v in vols], prm.intr, d_dev[i],
                                 eng._assoc_images(vols), 64.0, stats=st)
            c = st.cpu().numpy()
            exact = 16.0 * c[0] + 8.0 * c[1] + 4.0 * (c[2] + c[3])
            sec = ms_stage[2] * 1e-3
            roof.update({"achieved": 16.0 * int_vox / sec / 1e9, "frac": 16.0 * int_vox / sec / 1e9 / peak,
                         "algorithmic_bytes": 16.0 * int_vox, "kernel_ms": float(ms_stage[2]),
                         "exact_bytes": float(exact), "achieved_exact": float(exact) / sec / 1e9,
                         "frac_exact": float(exact) / sec / 1e9 / peak,
                         "convention": "achieved = 16 B/voxel x every voxel of the integrated volumes (upper bound: voxels outside "
                                       "the frustum move 0 B, so frac can exceed 1); *_exact = bytes the kernel really has to move "
                                       "(16 B updated, 8 B newly-occluded, 4 B weight-only), from its own counters",
                         "counters": {"updated": int(c[0]), "marked_occluded": int(c[1]), "occluded_seen": int(c[2]),
                                      "check_only": int(c[3]), "outside_image_in_interval": int(c[4]),
                                      "voxels_on_exact_path": int(c[5]), "segments_decided_free": int(c[6]),
                                      "segments_decided_occluded": int(c[7])}})
            del scratch
            if ray_roof and clocks and clocks.get("sm_mhz"):
                # the raycast is bound by issue slots, not bytes: warp instructions of one launch (ncu) against what the SMs can
                # issue in the kernel's share of the frame (148 SMs x 4 schedulers x clock)
                peak_issue = 148 * 4 * clocks["sm_mhz"] * 1e6
                t_ray = max(float(ms_stage[1]) - 0.016, 1e-6) * 1e-3        # (the composite launch is ~16 us of the stage)
                roof["raycast"] = {"bound": "issue", "kernel": "k_raycast", "warp_instructions": ray_roof["warp_instructions"],
                                   "achieved": ray_roof["warp_instructions"] / t_ray / 1e9, "peak": peak_issue / 1e9,
                                   "unit": "G warp-instructions/s", "frac": ray_roof["warp_instructions"] / t_ray / peak_issue,
                                   "kernel_ms": t_ray * 1e3, "samples": ray_roof.get("samples"),
                                   "note": "frac = issue slots used; the rest is the dependent chain of the longest rays (DESIGN.md section 4.2)"}
        out = {
            "metric": "Mvoxels/s (integrate+raycast+assoc)", "value": nvox / (ms_step * 1e-3) / 1e6, "unit": "Mvoxels/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "voxels_per_frame": nvox, "l2_policy": "working set (>1 GB/frame) > L2 (126 MB)",
                       "step": "one frame = computePoints + association + raycast/composite + integrate, K frames back to back",
                       "parallelism": (f"objects sharded over {world} GPU(s); background " +
                                       ("replicated, its raycast sharded by image rows" if eng.replicate_background else "on GPU 0") +
                                       ("" if world == 1 else ("; exchanges over NVLink peer memory" if eng._px is not None else "; exchanges via NCCL"))),
                       "gradients": "materialised per frame" if args.materialize_grads else "on the fly (no float3 volume)",
                       "brick_maps": bool(args.brick_maps), "ray_certificate": bool(use_cert), "visible_objects": len(eng.vis_objs)},
            "stages_ms": {"association": float(ms_stage[0]), "raycast+composite": float(ms_stage[1]),
                          "integrate": float(ms_stage[2]),
                          "note": None if world == 1 else "max over ranks per phase; association incl. the normaliser exchange, "
                                  "raycast+composite = local raycast + pre-composite + merge (or the wait for it), integrate = background "
                                  "(issued under the merge) + objects"},
            "e2e": {"value": nvox / (ms_e2e * 1e-3) / 1e6, "unit": "Mvoxels/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h * w * 4, "d2h_bytes_per_step": h * w * 5},
            "gpu_launches": launches, "host_issue_ms_per_step": t_host,
            # a wait of the peer-memory exchange that timed out would invalidate the run: 0 = none
            "exchange_errors": (eng._px.check_errors() if getattr(eng, "_px", None) is not None else 0),
            "parity_vs_n1": parity,
            "clocks": clocks,
            "stages_per_rank": stage_detail,
            "roofline": roof,
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                out["cpu_baseline"] = cpu_baseline(args.config)
            except Exception as e:  # the checker is optional for the measurement itself
                out["cpu_baseline"] = {"error": str(e)}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own kernels + launch structure on one B200 (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from tests import ref_gpu
    unchanged = args.impl == "unchanged-caller"
    if unchanged:
        if not os.path.exists(ref_gpu.B200OPS_PATH):
            emit({"impl": "unchanged-caller", "unavailable": "oracle/_ref/libemf_ref_b200ops.so not built (make -C oracle b200ops)"})
            return
        ref_gpu.use_library(ref_gpu.B200OPS_PATH)
    if not ref_gpu.available():
        emit({"impl": "reference", "unavailable": "oracle/_ref/libemf_ref.so not built (needs /root/reference at build time)"})
        return
    import torch
    from emfusion_b200.poses import Affine, rel_pose_CO, rel_pose_OC
    from emfusion_b200.synth import Scene
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    name, bg, k, ob, w, h = CONFIGS[args.config]
    scene = Scene(n_objects=k, width=w, height=h, seed=0)
    K = scene.K
    frames = render_stream(scene, N_STREAM_FRAMES)
    d_dev = [torch.from_numpy(d).to(dev) for d, _ in frames]
    d_pin = [torch.from_numpy(d).pin_memory() for d, _ in frames]
    vols = []
    bg_voxel = float(np.float32(5.12 / bg))
    specs = [(0, (bg,) * 3, bg_voxel, lambda f: Affine.translation([0, 0, 2.56]))]
    for i in range(k):
        specs.append((i + 1, (ob,) * 3, scene.object_voxel_size(i, ob), (lambda f, i=i: scene.object_pose(i, f))))
    for vid, res, voxel, _ in specs:
        n = int(np.prod(res))
        vols.append(dict(tsdf=torch.zeros(n, device=dev), weights=torch.zeros(n, device=dev),
                         grads=torch.zeros(3 * n, device=dev), fg_probs=torch.zeros(n, device=dev) if vid else None,
                         fgbg=torch.zeros(2 * n, device=dev), res=res, voxel=voxel,
                         trunc=float(np.float32(10.0) * np.float32(voxel)), id=vid))
    ref = ref_gpu.RefFrame(w, h, vols)
    for i in range(len(vols)):
        ref.fill_assoc(i, 1.0)
    points = torch.zeros((h, w, 3), device=dev)

    def poses(f):
        cam = scene.cam_pose(f % N_STREAM_FRAMES)
        ps = [s[3](f % N_STREAM_FRAMES) for s in specs]
        co = [rel_pose_CO(cam, p) for p in ps]
        oc = [rel_pose_OC(cam, p) for p in ps]
        return (np.concatenate([t.rotation32() for t in co]), np.concatenate([t.translation32() for t in co]),
                np.concatenate([t.rotation32() for t in oc]), np.concatenate([t.translation32() for t in oc]), oc)

    # set-up: frame 0 (association == 1) and fg probabilities from the analytic masks -- reference kernels only
    Rco, tco, Roc, toc, oc = poses(0)
    ref.integrate(d_dev[0], Roc, toc, K, 64.0, use_vis=False)
    zeros = torch.zeros((h, w), dtype=torch.uint8, device=dev)
    inst0 = torch.from_numpy(frames[0][1]).to(dev)
    for i, v in enumerate(vols):
        if v["id"] == 0:
            continue
        ref_gpu.update_fgbg((inst0 == v["id"]).to(torch.uint8), zeros, v["tsdf"], v["weights"], v["fgbg"],
                            oc[i].rotation32(), oc[i].translation32(), K, v["res"], v["voxel"])
        fb = v["fgbg"].view(-1, 2)
        s = fb[:, 0] + fb[:, 1]
        v["fg_probs"].copy_(torch.where(s != 0, fb[:, 0] / s, torch.zeros_like(s)))
    ref.update_fg_masks()
    cache = [poses(f) for f in range(N_STREAM_FRAMES)]

    def step(f, ev=None, depth=None):
        i = f % N_STREAM_FRAMES
        Rco, tco, Roc, toc, _ = cache[i]
        d = d_dev[i] if depth is None else depth
        ref_gpu.compute_points(d, points, K)
        if ev: ev[0].record()
        ref.assoc(points, Rco, tco)
        if ev: ev[1].record()
        ref.raycast(Rco, tco, K, 20, 1600)
        if ev: ev[2].record()
        ref.integrate(d, Roc, toc, K, 64.0, use_vis=True)
        if ev: ev[3].record()

    f = 1
    for _ in range(max(args.warmup, 3)):
        step(f); f += 1
    sampler = ClockSampler(0)
    # timed: EXACTLY K frames back to back between two events (every reference frame ends in a host barrier, as in
    # src/core/EMFusion.cpp:886-888, so this is also its wall time)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    for s in range(args.steps):
        step(f); f += 1
    stop.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_hot = start.elapsed_time(stop) / args.steps
    # stage breakdown, separate pass
    n_stage = min(args.steps, 10)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(n_stage)]
    for s in range(n_stage):
        step(f, ev[s]); f += 1
    torch.cuda.synchronize()
    stage = np.array([[e[j].elapsed_time(e[j + 1]) for j in range(3)] for e in ev])
    ms_stage = stage.mean(0)
    # e2e: host depth in, composited segmentation + raylengths out
    seg_host = torch.empty((h, w), dtype=torch.uint8).pin_memory()
    ray_host = torch.empty((h, w), dtype=torch.float32).pin_memory()
    d_in = torch.empty((h, w), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(args.steps):
        d_in.copy_(d_pin[f % N_STREAM_FRAMES], non_blocking=True)
        torch.cuda.synchronize()
        step(f, depth=d_in); f += 1
        ref.download_composite(seg_host, ray_host)
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
    nvox = total_voxels(args.config)
    n_vis = sum(1 for i in range(1, len(vols)) if ref.visible(i))
    out = {"impl": args.impl, "metric": "Mvoxels/s (integrate+raycast+assoc)", "value": nvox / (ms_hot * 1e-3) / 1e6,
           "unit": "Mvoxels/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_hot,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": name, "voxels_per_frame": nvox, "visible_objects": n_vis,
                      "step": "one frame = computePoints + association + raycast/composite + integrate (+ gradients), K frames back to back",
                      "what": ("this repo's level-1 operators (bindings/emf_b200_opencv_binding.cpp on libemf_b200.so) driven with the "
                               "reference's launch structure: what an UNCHANGED src/core/EMFusion.cpp gets" if unchanged else
                               "reference CUDA kernels (src/core/cuda/*.cu compiled unchanged, sm_100a)") + " + restated OpenCV-CUDA "
                              "element-wise launches, per-volume streams and host barriers as in src/core/EMFusion.cpp"},
           "stages_ms": {"association": float(ms_stage[0]), "raycast+composite": float(ms_stage[1]), "integrate+gradients": float(ms_stage[2])},
           "e2e": {"value": nvox / (ms_e2e * 1e-3) / 1e6, "unit": "Mvoxels/s", "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": h * w * 4, "d2h_bytes_per_step": h * w * 5},
           "cpu_baseline": {"value": nvox / (ms_hot * 1e-3) / 1e6, "unit": "Mvoxels/s", "cores": 0, "kind": "reference",
                            "sample": "full workload on one B200 -- the reference path is CUDA-only; no host cores are used"},
           "clocks": clocks}
    emit(out)
    ref.close()


_STDOUT_FD = None


def quiet_stdout():
    """stdout carries exactly one JSON line: everything else a library writes there (NCCL's version banner, ...) is sent to
    stderr until emit() restores the descriptor"""
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(obj):
    sys.stdout.flush()
    if _STDOUT_FD is not None:
        os.dup2(_STDOUT_FD, 1)
    print(json.dumps(obj), flush=True)
    if _STDOUT_FD is not None:
        os.dup2(2, 1)


if __name__ == "__main__":
    a = parse()
    quiet_stdout()
    if a.impl in ("reference", "unchanged-caller"):
        run_reference(a)
    else:
        run_ours(a)
