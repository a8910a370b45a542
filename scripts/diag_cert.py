"""GPU diagnostic: bench scene, background raycast with / without the ray-space certificate: samples taken / skipped, ms."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emfusion_b200 import ops
from emfusion_b200.native import NativeEngine
from emfusion_b200.poses import rel_pose_CO
from emfusion_b200.synth import Scene
from emfusion_b200.volume import ObjTSDF, Params
bg, k, ob, w, h = 512, 32, 128, 640, 480
dev = torch.device("cuda:0")
scene = Scene(n_objects=k, width=w, height=h, seed=0)
prm = Params(frameSize=(w, h), intr=scene.K, globalVolumeDims=(bg,) * 3, globalVoxelSize=5.12 / bg, objVolumeDims=(ob,) * 3)
ObjTSDF.nextID = 0
eng = NativeEngine(prm, dev)
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
b = eng.background
v = b.c_volume()
T = rel_pose_CO(scene.cam_pose(10), b.pose)
z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=dev)
ray, vert, norm, mask = [z(h, w)], [z(h, w, 3)], [z(h, w, 3)], [z(h, w, dt=torch.uint8)]
ws = ops.raycastWorkspace(w, h, dev)
def timeit(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, wk in (("plain", None), ("cert", ws)):
    st = torch.zeros(32, dtype=torch.int64, device=dev)
    ops.raycastVolumes([v], [T], prm.intr, [[0, 0, w, h]], ray, vert, norm, mask, stats=st, workspace=wk)
    ms = timeit(lambda: ops.raycastVolumes([v], [T], prm.intr, [[0, 0, w, h]], ray, vert, norm, mask, workspace=wk))
    print(name, "bg raycast ms", round(ms, 4), "stats [taken, skipped, jumps, weight samples, cert warp-iters, cert lane-iters, plain warp-iters, plain lane-iters]", st.cpu().numpy().tolist())
for rect in ([64, 48, 576, 432], [0, 0, 640, 48]) if os.environ.get("EMF_RAY_CERT") != "2" else ():
    for name, wk in (("plain", None), ("cert", ws)):
        st = torch.zeros(8, dtype=torch.int64, device=dev)
        ops.raycastVolumes([v], [T], prm.intr, [rect], ray, vert, norm, mask, stats=st, workspace=wk)
        ms = timeit(lambda: ops.raycastVolumes([v], [T], prm.intr, [rect], ray, vert, norm, mask, workspace=wk))
        print(rect, name, "ms", round(ms, 4), st.cpu().numpy().tolist())
ops.raycastVolumes([v], [T], prm.intr, [[0, 0, w, h]], ray, vert, norm, mask, workspace=ws)
tb = ws[: 256 * 300 * 4].view(torch.int32).reshape(256, 300)
bits = ((tb[..., None] >> torch.arange(32, device=dev)) & 1).to(torch.float32)
print("certified (tile, slab) fraction", float(bits.mean()), "per slab (every 16th):", [round(float(bits[j].mean()), 2) for j in range(0, 256, 16)])
