"""GPU diagnostic (not part of the product or the tests): bench scenario, acceleration statistics and stage timings."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emfusion_b200 import ops
from emfusion_b200.engine import EMFusionEngine
from emfusion_b200.synth import Scene
from emfusion_b200.volume import ObjTSDF, Params
from emfusion_b200.poses import rel_pose_CO, rel_pose_OC

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
b = eng.background
def popfrac(bits, res):
    words = bits.numel() // 3
    m = bits.reshape(3, words).to(torch.int64) & 0xFFFFFFFF
    cnt = torch.zeros(3)
    for k in range(3):
        x = m[k]
        c = 0
        for sft in range(32):
            c += int(((x >> sft) & 1).sum())
        cnt[k] = c / (res[0] // 4 * res[1] * res[2])
    return cnt.numpy()
print("bg const frac (+1, 0, -1)", popfrac(b.constBits, b.volumeRes))
print("bg safe  frac (+1, 0, -1)", popfrac(b.safeBits, b.volumeRes))
o1 = eng.objects[0]
print("obj const frac", popfrac(o1.constBits, o1.volumeRes), "safe", popfrac(o1.safeBits, o1.volumeRes))

def timeit(fn, n=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

i = 10
eng.pose = scene.cam_pose(i)
eng.set_depth(d_dev[i])
vols = eng.local_volumes()
rects = eng._rects(vols)
T = [rel_pose_CO(eng.pose, v.pose) for v in vols]
ray = [eng.bg_raylengths] + [eng.obj_raylengths[o.id] for o in eng.objects]
vert = [eng.bg_vertices] + [eng.obj_vertices[o.id] for o in eng.objects]
norm = [eng.bg_normals] + [eng.obj_normals[o.id] for o in eng.objects]
mask = [eng.bg_mask] + [eng.obj_modelSegmentation[o.id] for o in eng.objects]
def rc(vs, sl=slice(None)):
    ops.raycastVolumes(vs[sl], T[sl], prm.intr, rects[sl], ray[sl], vert[sl], norm[sl], mask[sl])
cv = [v.c_volume(with_grads=True) for v in vols]
cv_nb = [ops.volume(v.tsdfVol, v.tsdfWeights, v.volumeRes, v.voxelSize, v.truncdist, fg_probs=v._fg(), vid=v.id) for v in vols]
for name, vs in (("bricks", cv), ("nobricks", cv_nb)):
    st = torch.zeros(4, dtype=torch.int64, device=dev)
    ops.raycastVolumes(vs[:1], T[:1], prm.intr, rects[:1], ray[:1], vert[:1], norm[:1], mask[:1], stats=st)
    print("bg raycast stats", name, "(samples, skipped, crawl attempts, weight samples):", st.cpu().numpy())
    st.zero_()
    ops.raycastVolumes(vs[1:], T[1:], prm.intr, rects[1:], ray[1:], vert[1:], norm[1:], mask[1:], stats=st)
    print("obj raycast stats", name, st.cpu().numpy(), "rect px", sum((r[2]-r[0])*(r[3]-r[1]) for r in rects[1:]))
print("raycast all, bricks   ms", timeit(lambda: rc(cv)))
r1 = eng.bg_raylengths.clone(); m1 = eng.bg_mask.clone()
print("raycast all, nobricks ms", timeit(lambda: rc(cv_nb)))
print("same result:", bool((r1 == eng.bg_raylengths).all()), bool((m1 == eng.bg_mask).all()), "hits", int(m1.sum()))
print("raycast bg only, bricks   ms", timeit(lambda: rc(cv, slice(0, 1))))
print("raycast bg only, nobricks ms", timeit(lambda: rc(cv_nb, slice(0, 1))))
print("raycast objs only, bricks   ms", timeit(lambda: rc(cv, slice(1, None))))
print("raycast objs only, nobricks ms", timeit(lambda: rc(cv_nb, slice(1, None))))
# integrate
Toc = [rel_pose_OC(eng.pose, v.pose) for v in vols]
assoc = eng._assoc_images(vols)
stats = torch.zeros(5, dtype=torch.int64, device=dev)
cvi = [v.c_volume() for v in vols]
ops.integrateVolumes(cvi, Toc, prm.intr, eng.depth, assoc, 64.0, stats=stats)
print("integrate stats (updated, marked, occl-seen, check, skipped-in-interval):", stats.cpu().numpy(), "of", sum(v.numVoxels() for v in vols))
stats.zero_()
ops.integrateVolumes(cvi[:1], Toc[:1], prm.intr, eng.depth, assoc[:1], 64.0, stats=stats)
print("bg only stats:", stats.cpu().numpy(), "of", vols[0].numVoxels())
print("integrate all ms", timeit(lambda: ops.integrateVolumes(cvi, Toc, prm.intr, eng.depth, assoc, 64.0)))
print("integrate bg  ms", timeit(lambda: ops.integrateVolumes(cvi[:1], Toc[:1], prm.intr, eng.depth, assoc[:1], 64.0)))
print("integrate objs ms", timeit(lambda: ops.integrateVolumes(cvi[1:], Toc[1:], prm.intr, eng.depth, assoc[1:], 64.0)))
cvi_ns = [ops.volume(v.tsdfVol, v.tsdfWeights, v.volumeRes, v.voxelSize, v.truncdist, fg_probs=v._fg(), vid=v.id) for v in vols]
print("integrate all (no seg status) ms", timeit(lambda: ops.integrateVolumes(cvi_ns, Toc, prm.intr, eng.depth, assoc, 64.0)))
print("safe bits ms", timeit(lambda: ops.updateSafeBits(cvi)))
print("assoc ms", timeit(lambda: eng.computeAssociationWeights()))

# ---- structure of the maps along the viewing direction (bg): strings of bits along z for a few (x, y) columns
def unpack(bits, res):
    rx, ry, rz = res
    wpr = (rx // 4 + 31) // 32
    w = bits.reshape(3, rz, ry, wpr).to(torch.int64) & 0xFFFFFFFF
    sh = torch.arange(32, device=bits.device, dtype=torch.int64)
    return ((w[..., None] >> sh) & 1).to(torch.bool).reshape(3, rz, ry, wpr * 32)[..., : rx // 4]
cb = unpack(b.constBits, b.volumeRes)[0]   # ones map (rz, ry, rx/4)
sb = unpack(b.safeBits, b.volumeRes)[0]
t = b.tsdfVol.reshape(512, 512, 512)
for (x, y) in ((256, 256), (200, 300), (300, 200), (256, 140)):
    print("column x=%d y=%d  (z = 0..511, every voxel)" % (x, y))
    print(" tsdf==1:", "".join("1" if v else "." for v in (t[:, y, x] == 1.0).cpu().numpy()))
    print(" const  :", "".join("1" if v else "." for v in cb[:, y, x // 4].cpu().numpy()))
    print(" safe   :", "".join("1" if v else "." for v in sb[:, y, x // 4].cpu().numpy()))
z = 150
print("slice z=150, rows y=200..215, segments 32..96")
for y in range(200, 216):
    print(" c:", "".join("1" if v else "." for v in cb[z, y, 32:96].cpu().numpy()), " s:", "".join("1" if v else "." for v in sb[z, y, 32:96].cpu().numpy()))
