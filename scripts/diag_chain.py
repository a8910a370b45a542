"""GPU diagnostic: time of ONE CTA tile of the background raycast (the dependent chain of its longest ray)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "diag_cert.py")).read().split("b = eng.background")[0])
b = eng.background
v = b.c_volume()
T = rel_pose_CO(scene.cam_pose(10), b.pose)
z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=dev)
ray, vert, norm, mask = [z(h, w)], [z(h, w, 3)], [z(h, w, 3)], [z(h, w, dt=torch.uint8)]
ws = ops.raycastWorkspace(w, h, dev)
def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for rect in ([0, 0, 16, 8], [624, 0, 640, 8], [0, 472, 16, 480], [320, 240, 336, 248], [0, 0, 640, 8]):
    for name, wk in (("plain", None), ("cert", ws)):
        st = torch.zeros(32, dtype=torch.int64, device=dev)
        ops.raycastVolumes([v], [T], prm.intr, [rect], ray, vert, norm, mask, stats=st, workspace=wk)
        ms = timeit(lambda: ops.raycastVolumes([v], [T], prm.intr, [rect], ray, vert, norm, mask, workspace=wk))
        s = st.cpu().numpy().tolist()
        print(rect, name, "ms", round(ms, 4), "taken", s[0], "skipped", s[1], "warp-iters", s[4], "lane-iters", s[5], "us/warp-iter(4 warps)", round(ms * 1e3 / max(s[4] / 4, 1), 3))
