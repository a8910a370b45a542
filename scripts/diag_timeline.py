"""GPU diagnostic: per-warp timeline of the whole frame's raycast launch (bench scene): who runs when, for how long."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("EMF_RAY_TIMELINE") != "1":
    os.environ["EMF_RAY_HIST"] = "2"
from emfusion_b200 import ops
from emfusion_b200.engine import EMFusionEngine
from emfusion_b200.poses import rel_pose_CO
from emfusion_b200.synth import Scene
from emfusion_b200.volume import ObjTSDF, Params
bg, k, ob, w, h = 512, 32, 128, 640, 480
dev = torch.device("cuda:0")
scene = Scene(n_objects=k, width=w, height=h, seed=0)
prm = Params(frameSize=(w, h), intr=scene.K, globalVolumeDims=(bg,) * 3, globalVoxelSize=5.12 / bg, objVolumeDims=(ob,) * 3)
ObjTSDF.nextID = 0
eng = EMFusionEngine(prm, dev)
for i in range(k):
    eng.add_object(scene.object_pose(i, 0), scene.object_voxel_size(i, ob))
frames = [scene.render(f) for f in range(12)]
d_dev = [torch.from_numpy(d).to(dev) for d, _ in frames]
eng.processFrame(d_dev[0], scene.cam_pose(0), {o.id: scene.object_pose(o.id - 1, 0) for o in eng.objects})
zeros = torch.zeros((h, w), dtype=torch.uint8, device=dev)
inst0 = torch.from_numpy(frames[0][1]).to(dev)
for o in eng.objects:
    o.integrateMask((inst0 == o.id).to(torch.uint8), zeros, eng.pose, prm.intr)
for f in range(1, 10):
    i = f % 12
    eng.processFrame(d_dev[i], scene.cam_pose(i), {o.id: scene.object_pose(o.id - 1, i) for o in eng.objects})
i = 10
eng.pose = scene.cam_pose(i)
for o in eng.objects:
    o.pose = scene.object_pose(o.id - 1, i)
eng.set_depth(d_dev[i])
vols = eng.local_volumes()
rects = eng._rects(vols)
T = [rel_pose_CO(eng.pose, v.pose) for v in vols]
ray = [eng.bg_raylengths] + [eng.obj_raylengths[o.id] for o in eng.objects]
vert = [eng.bg_vertices] + [eng.obj_vertices[o.id] for o in eng.objects]
norm = [eng.bg_normals] + [eng.obj_normals[o.id] for o in eng.objects]
mask = [eng.bg_mask] + [eng.obj_modelSegmentation[o.id] for o in eng.objects]
cv = [v.c_volume(with_grads=True) for v in vols]
ws = ops.raycastWorkspace(w, h, dev)
nblk = sum(((r[2] - r[0] + 15) // 16) * ((r[3] - r[1] + 7) // 8) for r in rects)
TL = os.environ.get("EMF_RAY_TIMELINE") == "1"
for name, wk in (("plain", None), ("sched", ws)) + (() if TL else (("cert", ws),)):
    for rep in range(3):
        st = torch.zeros(32 + 4 * 4 * nblk, dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        ops.raycastVolumes(cv, T, prm.intr, rects, ray, vert, norm, mask, stats=st, workspace=wk, certificate=name == "cert")
        torch.cuda.synchronize()
    a = st.cpu().numpy()
    rec = a[32:].reshape(-1, 4)
    rec = rec[rec[:, 0] > 0]
    t0 = rec[:, 0].min()
    beg, end, it, vol = (rec[:, 0] - t0) / 1e3, (rec[:, 1] - t0) / 1e3, rec[:, 2], rec[:, 3] & 0xffffffff
    print(f"== {name}: warps {len(rec)} kernel span {end.max():.1f} us; counters {a[:8].tolist()}")
    isbg = vol == 0
    for lbl, m in (("bg", isbg), ("obj", ~isbg)):
        if m.any():
            d = end[m] - beg[m]
            print(f"  {lbl}: warps {m.sum()} start [{beg[m].min():.1f}, {beg[m].max():.1f}] end max {end[m].max():.1f} dur mean {d.mean():.1f} p50 {np.percentile(d,50):.1f} p90 {np.percentile(d,90):.1f} p99 {np.percentile(d,99):.1f} max {d.max():.1f}; iters mean {it[m].mean():.1f} max {it[m].max()}")
    # the warps that end last
    order = np.argsort(-end)[:12]
    for o_ in order:
        print(f"    last: vol {vol[o_]} start {beg[o_]:.1f} end {end[o_]:.1f} dur {end[o_]-beg[o_]:.1f} iters {it[o_]} us/iter {(end[o_]-beg[o_])/max(it[o_],1):.2f}")
    # busy warps over time
    ts = np.linspace(0, end.max(), 21)
    print("  active warps at t:", [(round(t), int(((beg <= t) & (end > t)).sum())) for t in ts])
