"""CPU diagnostic (C oracle; not part of the product or the tests): where the samples of a background raycast go on the
bench scene -- per-ray counts by march step size (truncation distance / one voxel / half a voxel) and the share of the
volume that holds the constants +1 / 0 / -1.  Feeds DESIGN.md sections 4.3 and 9.  Takes about a minute and 3 GB."""
import sys, time, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from tests import oracle_c, scenario as S
from emfusion_b200.poses import rel_pose_CO, rel_pose_OC
from emfusion_b200.synth import Scene
o = oracle_c.load()
bg, w, h, k = 512, 640, 480, 32
scene = Scene(n_objects=k, width=w, height=h, seed=0)
vs = float(np.float32(5.12 / bg)); tr = float(np.float32(10.0) * np.float32(vs))
from emfusion_b200.poses import Affine
pose = Affine.translation([0, 0, 2.56])
n = bg ** 3
tsdf = np.zeros(n, np.float32); wts = np.zeros(n, np.float32)
ones = np.ones((h, w), np.float32)
t0 = time.time()
for f in range(8):
    d, _ = scene.render(f)
    T = rel_pose_OC(scene.cam_pose(f), pose)
    o.update_tsdf(d, ones, tsdf, wts, S.R9(T), S.T3(T), scene.K, (bg,) * 3, vs, tr, 64.0)
print("integrated", time.time() - t0)
g = o.compute_grads(tsdf, (bg,) * 3)
T = rel_pose_CO(scene.cam_pose(8), pose)
r = o.raycast(tsdf, g, wts, S.R9(T), S.T3(T), scene.K, (bg,) * 3, vs, tr, w, h, step_stats=True)
si = r["step_img"]
tot = si.sum(-1).astype(np.int64)
print("samples", int(r["steps"][0]), "| per ray: step = truncdist %.1f, one voxel %.1f, half a voxel %.1f" % (si[..., 0].mean(), si[..., 1].mean(), si[..., 2].mean()))
print("percentiles of samples / ray (5, 25, 50, 75, 95):", np.percentile(tot, [5, 25, 50, 75, 95]))
# where along the ray do samples happen? classify voxels: fraction of all voxels exactly 1, 0, -1, other
print("volume: ==1 %.3f ==0 %.3f ==-1 %.3f other %.3f" % ((tsdf == 1).mean(), (tsdf == 0).mean(), (tsdf == -1).mean(), ((tsdf != 1) & (tsdf != 0) & (tsdf != -1)).mean()))
