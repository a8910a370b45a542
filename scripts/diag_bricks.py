"""GPU diagnostic: brick-level split of the bench frame's integrate (how many bricks are decided whole / mixed, per volume kind)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "diag_cert.py")).read().split("b = eng.background")[0])
from emfusion_b200.poses import rel_pose_OC
from emfusion_b200 import _lib
i = 10
eng.pose = scene.cam_pose(i)
for o in eng.objects:
    o.pose = scene.object_pose(o.id - 1, i)
eng.set_depth(d_dev[i]); eng.computeAssociationWeights(); eng.raycast()
vis = eng.vis_objs
vols = [v for v in eng.local_volumes() if v.id == 0 or v.id in vis]
print("visible objects", len(vis))
for name, sel in (("all", vols), ("bg", vols[:1]), ("objs", vols[1:])):
    scratch = [(v.tsdfVol.clone(), v.tsdfWeights.clone()) for v in sel]
    cv = [ops.volume(t_, w_, v.volumeRes, v.voxelSize, v.truncdist, vid=v.id) for v, (t_, w_) in zip(sel, scratch)]
    Toc = [rel_pose_OC(eng.pose, v.pose) for v in sel]
    assoc = eng._assoc_images(sel)
    ws = ops.integrateWorkspace(eng.depth)
    st = torch.zeros(8, dtype=torch.int64, device=dev)
    ops.integrateVolumes(cv, Toc, prm.intr, eng.depth, assoc, 64.0, stats=st, workspace=ws)
    torch.cuda.synchronize()
    total = int(_lib.lib().emf_integrate_workspace_bytes(w, h))
    need = total - 256 - 2 * 8 * (3 << 20)
    cnt = ws[need:need + 12].view(torch.int32).cpu().numpy()
    nb = sum(v.numVoxels() for v in sel) // 512
    def timeit(fn, n=10):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    ms = timeit(lambda: ops.integrateVolumes(cv, Toc, prm.intr, eng.depth, assoc, 64.0, workspace=ws))
    print(f"{name}: bricks {nb} mixed {cnt[0]} ({cnt[0]/nb:.3f}) whole {cnt[1]} ({cnt[1]/nb:.3f}) out {nb-cnt[0]-cnt[1]}; ms {ms:.4f}; stats {st.cpu().numpy().tolist()}")
    del scratch
