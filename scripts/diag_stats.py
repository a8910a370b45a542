"""GPU diagnostic (not part of the product or the tests): bench scene (config 4), per-part timings and the
work counters of the raycast and integrate kernels.  Output feeds DESIGN.md's byte models."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emfusion_b200 import ops
from emfusion_b200.engine import EMFusionEngine
from emfusion_b200.poses import rel_pose_CO, rel_pose_OC
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
torch.cuda.synchronize()


def timeit(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


i = 10
eng.pose = scene.cam_pose(i)
for o in eng.objects:
    o.pose = scene.object_pose(o.id - 1, i)
eng.set_depth(d_dev[i])
eng.computeAssociationWeights()
vols = eng.local_volumes()
rects = eng._rects(vols)
T = [rel_pose_CO(eng.pose, v.pose) for v in vols]
ray = [eng.bg_raylengths] + [eng.obj_raylengths[o.id] for o in eng.objects]
vert = [eng.bg_vertices] + [eng.obj_vertices[o.id] for o in eng.objects]
norm = [eng.bg_normals] + [eng.obj_normals[o.id] for o in eng.objects]
mask = [eng.bg_mask] + [eng.obj_modelSegmentation[o.id] for o in eng.objects]
cv = [v.c_volume(with_grads=True) for v in vols]


def rc(sl=slice(None), stats=None):
    ops.raycastVolumes(cv[sl], T[sl], prm.intr, rects[sl], ray[sl], vert[sl], norm[sl], mask[sl], stats=stats)


for name, sl in (("bg", slice(0, 1)), ("objs", slice(1, None))):
    st = torch.zeros(8, dtype=torch.int64, device=dev)
    rc(sl, st)
    px = sum((r[2] - r[0]) * (r[3] - r[1]) for r in rects[sl])
    hits = sum(int(m.sum()) for m in mask[sl])
    print(f"raycast {name}: rect px {px} hits {hits} stats {st.cpu().numpy().tolist()}")
cv_nb = [ops.volume(v.tsdfVol, v.tsdfWeights, v.volumeRes, v.voxelSize, v.truncdist, fg_probs=v._fg(), vid=v.id) for v in vols]
rc()
keep = [x.clone() for x in ray + vert + norm + mask]
ops.raycastVolumes(cv_nb, T, prm.intr, rects, ray, vert, norm, mask)
same = all(bool((a.view(torch.int32) == b.view(torch.int32)).all()) if a.dtype == torch.float32 else bool((a == b).all())
           for a, b in zip(keep, ray + vert + norm + mask))
print("raycast with brick maps == without (bitwise, all volumes):", same)
print("raycast all, no brick maps ms", timeit(lambda: ops.raycastVolumes(cv_nb, T, prm.intr, rects, ray, vert, norm, mask)))
print("brick maps update ms", timeit(lambda: ops.updateBrickMaps(cv)))
print("raycast all ms", timeit(rc))
print("raycast bg  ms", timeit(lambda: rc(slice(0, 1))))
print("raycast obj ms", timeit(lambda: rc(slice(1, None))))
eng.raycast()
print("raycast+composite (engine) ms", timeit(eng.raycast))
print("visible", len(eng.vis_objs))
Toc = [rel_pose_OC(eng.pose, v.pose) for v in vols]
assoc = eng._assoc_images(vols)
cvi = [v.c_volume() for v in vols]
for name, sl in (("all", slice(None)), ("bg", slice(0, 1)), ("objs", slice(1, None))):
    stats = torch.zeros(8, dtype=torch.int64, device=dev)
    ops.integrateVolumes(cvi[sl], Toc[sl], prm.intr, eng.depth, assoc[sl], 64.0, stats=stats)
    print(f"integrate {name} stats {stats.cpu().numpy().tolist()} of {sum(v.numVoxels() for v in vols[sl])} voxels")
    print(f"integrate {name} ms", timeit(lambda: ops.integrateVolumes(cvi[sl], Toc[sl], prm.intr, eng.depth, assoc[sl], 64.0)))
print("assoc ms", timeit(eng.computeAssociationWeights))
print("points ms", timeit(lambda: eng.set_depth(d_dev[i])))

# ---- brick map coverage vs ground truth (background)
b = eng.background
if b.brickMap is not None:
    R = b.volumeRes[0]; nb = R // 8
    t = b.tsdfVol.reshape(nb, 8, nb, 8, nb, 8)
    m = b.brickMap[: nb ** 3].reshape(nb, nb, nb).to(torch.int64)
    code, Pf, D = m >> 4, (m >> 3) & 1, m & 7
    for k, c in enumerate((1.0, 0.0, -1.0)):
        truth = (t == c).all(dim=5).all(dim=3).all(dim=1)
        print(f"bg bricks constant {c:+.0f}: truth {float(truth.float().mean()):.4f} flagged {float((code == k + 1).float().mean()):.4f} "
              f"P {float(((code == k + 1) & (Pf == 1)).float().mean()):.4f} D>=2 {float(((code == k + 1) & (D >= 2)).float().mean()):.4f}")
    print("bg voxels == 1:", float((b.tsdfVol == 1).float().mean()), " == 0:", float((b.tsdfVol == 0).float().mean()), " == -1:", float((b.tsdfVol == -1).float().mean()))
