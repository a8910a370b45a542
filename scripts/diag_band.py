"""GPU diagnostic: the background's raycast restricted to a band of image rows (what one of N GPUs traces), timed alone.
Measured on a young model (10 frames): rows 0-60 0.255 ms, 180-240 0.241, 420-480 0.381, 0-240 0.360, 0-480 0.772 -- a band is as
long as its longest ray: ~600 pair iterations x ~0.4 us of DEPENDENT INSTRUCTIONS (220 per pair at ~2 in flight), not of cache
misses: prefetch hints 4 iterations ahead (tsdf lines, with or without the weights' lines; 112 registers, no spills) made every
band slower (0.289 / 0.280 / 0.419 / 0.489 / 1.017 ms) and were removed again."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "diag_timeline.py")).read().split("nblk = sum(")[0].replace('os.environ["EMF_RAY_HIST"] = "2"', "pass"))
def timeit(fn, n=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
bgv = [vols[0].c_volume(with_grads=False)]
for band in ((0, 60), (180, 240), (420, 480), (0, 240), (0, 480)):
    r = [[0, band[0], w, band[1]]]
    ms = timeit(lambda: ops.raycastVolumes(bgv, T[:1], prm.intr, r, ray[:1], vert[:1], norm[:1], mask[:1]))
    print(f"rows {band}: {ms:.4f} ms")
