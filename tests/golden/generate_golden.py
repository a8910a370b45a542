"""Generates tests/golden/ref_kernels_small.npz by running the REFERENCE'S OWN CUDA kernels
(oracle/_ref/libemf_ref.so = /root/reference/src/core/cuda/{TSDF,ObjTSDF}.cu compiled unchanged against
oracle/shim) on a B200:

    gpurun -- 'python tests/golden/generate_golden.py gpurun_out/ref_kernels_small.npz'

and then copied into tests/golden/.  The reference ships no golden vectors of its own;
these are the fixtures that pin the C oracle on CPU (tests/test_oracle_golden.py) and the product on GPU.
Inputs are stored next to the outputs so the fixture is self-contained.
"""
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/tests/", 1)[0])
from emfusion_b200.poses import Affine, rel_pose_CO, rel_pose_OC  # noqa: E402
from emfusion_b200.synth import Scene  # noqa: E402
from tests import ref_gpu  # noqa: E402

W, H = 160, 120
BG_RES, OBJ_RES = (40, 36, 44), (24, 24, 24)


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def main(out_path):
    scene = Scene(n_objects=1, width=W, height=H, seed=3, dropout=0.02)
    K = scene.K
    out = dict(K=K, W=W, H=H, bg_res=np.array(BG_RES, np.int32), obj_res=np.array(OBJ_RES, np.int32))
    bg_voxel = np.float32(5.12 / 44)
    obj_voxel = np.float32(scene.object_voxel_size(0, OBJ_RES[0]))
    vols = [dict(res=BG_RES, voxel=float(bg_voxel), trunc=float(np.float32(10) * bg_voxel),
                 pose=lambda f: Affine.translation([0, 0, 2.56])),
            dict(res=OBJ_RES, voxel=float(obj_voxel), trunc=float(np.float32(10) * obj_voxel),
                 pose=lambda f: scene.object_pose(0, f))]
    out["voxel"] = np.array([v["voxel"] for v in vols], np.float32)
    out["trunc"] = np.array([v["trunc"] for v in vols], np.float32)
    rng = np.random.default_rng(11)
    state = [dict(tsdf=torch.zeros(int(np.prod(v["res"])), device="cuda"),
                  weights=torch.zeros(int(np.prod(v["res"])), device="cuda")) for v in vols]
    fgbg = torch.zeros(2 * int(np.prod(OBJ_RES)), device="cuda")
    for f in range(3):
        depth, inst = scene.render(f)
        assoc = np.ones((H, W), np.float32) if f == 0 else rng.random((H, W), dtype=np.float32)
        out[f"depth{f}"] = depth
        out[f"inst{f}"] = inst
        out[f"assoc{f}"] = assoc
        cam = scene.cam_pose(f)
        for i, v in enumerate(vols):
            T = rel_pose_OC(cam, v["pose"](f))
            out[f"R_oc{f}_{i}"], out[f"t_oc{f}_{i}"] = T.rotation32(), T.translation32()
            ref_gpu.update_tsdf(cu(depth), cu(assoc), state[i]["tsdf"], state[i]["weights"], T.rotation32(),
                                T.translation32(), K, v["res"], v["voxel"], v["trunc"], 64.0)
            torch.cuda.synchronize()
            out[f"tsdf{f}_{i}"] = state[i]["tsdf"].cpu().numpy()
            out[f"weights{f}_{i}"] = state[i]["weights"].cpu().numpy()
        T = rel_pose_OC(cam, vols[1]["pose"](f))
        occl = (rng.random((H, W)) < 0.1).astype(np.uint8)
        out[f"occl{f}"] = occl
        ref_gpu.update_fgbg(cu((inst == 1).astype(np.uint8)), cu(occl), state[1]["tsdf"], state[1]["weights"], fgbg,
                            T.rotation32(), T.translation32(), K, OBJ_RES, vols[1]["voxel"])
        torch.cuda.synchronize()
        out[f"fgbg{f}"] = fgbg.cpu().numpy()
    # gradient, raycast and gather from the last state at the frame-2 pose
    cam = scene.cam_pose(2)
    pts_depth = out["depth2"]
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    pts = np.stack([(xs - K[0, 2]) * pts_depth / K[0, 0], (ys - K[1, 2]) * pts_depth / K[1, 1], pts_depth], -1).astype(np.float32)
    out["points2"] = pts
    for i, v in enumerate(vols):
        n = int(np.prod(v["res"]))
        grads = torch.full((n, 3), 5.0, device="cuda")
        ref_gpu.update_gradients(state[i]["tsdf"], grads, v["res"])
        T = rel_pose_CO(cam, v["pose"](2))
        out[f"R_co_{i}"], out[f"t_co_{i}"] = T.rotation32(), T.translation32()
        ray = torch.zeros((H, W), device="cuda"); vert = torch.zeros((H, W, 3), device="cuda")
        norm = torch.zeros((H, W, 3), device="cuda"); mask = torch.zeros((H, W), dtype=torch.uint8, device="cuda")
        ref_gpu.raycast(state[i]["tsdf"], grads, state[i]["weights"], ray, vert, norm, mask, T.rotation32(),
                        T.translation32(), K, v["res"], v["voxel"], v["trunc"])
        vals = torch.zeros((H, W), device="cuda")
        ref_gpu.get_volume_vals(state[i]["tsdf"], cu(pts), T.rotation32(), T.translation32(), v["res"], v["voxel"], vals)
        torch.cuda.synchronize()
        out[f"grads_{i}"] = grads.cpu().numpy()
        out[f"ray_{i}"], out[f"vert_{i}"] = ray.cpu().numpy(), vert.cpu().numpy()
        out[f"norm_{i}"], out[f"mask_{i}"] = norm.cpu().numpy(), mask.cpu().numpy()
        out[f"gather_{i}"] = vals.cpu().numpy()
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, "hits:", int(out["mask_0"].sum()), int(out["mask_1"].sum()))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_kernels_small.npz")
