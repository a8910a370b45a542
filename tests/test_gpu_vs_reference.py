"""-m gpu: three-way check on a B200 -- the reference's own kernels (oracle/_ref/libemf_ref.so, compiled
unchanged from /root/reference against the type shim) vs the C oracle vs the product kernels.
This is what pins the oracle ("parity unpinned" otherwise: the reference has no tests or fixtures) and what
BASELINE.json's "outputs match the reference's own CUDA path" means."""
import numpy as np
import pytest
import torch

from emfusion_b200 import ops
from emfusion_b200.poses import rel_pose_CO, rel_pose_OC
from tests import ref_gpu, scenario as S
from tests.test_gpu_parity import ASSOC_TOL, DEV, SCENARIOS, assert_bits, cu

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_gpu.available(), reason="oracle/_ref not built")]


@pytest.fixture(scope="module", params=SCENARIOS, ids=[s[0] for s in SCENARIOS])
def scn(request, oracle):
    name, w, h, bg_res, n_obj, obj_res, kw = request.param
    return S.make(name, oracle, w, h, bg_res, n_obj, obj_res, n_frames=3, integrate_frames=2, **kw)


def test_integrate_three_way(scn, oracle, cuda_dev):
    f = 2
    rng = np.random.default_rng(5)
    assoc = rng.random((scn.h, scn.w), dtype=np.float32)
    assoc[rng.random((scn.h, scn.w)) < 0.1] = 0.0
    for v in scn.vols():
        pose = v.pose if v.vid == 0 else scn.scene.object_pose(v.vid - 1, f)
        T = rel_pose_OC(scn.cam(f), pose)
        t_o, w_o = v.tsdf.copy(), v.weights.copy()
        oracle.update_tsdf(scn.depths[f], assoc, t_o, w_o, S.R9(T), S.T3(T), scn.K, v.res, v.voxel, v.trunc, 64.0)
        t_r, w_r = cu(v.tsdf), cu(v.weights)
        ref_gpu.update_tsdf(cu(scn.depths[f]), cu(assoc), t_r, w_r, S.R9(T), S.T3(T), scn.K, v.res, v.voxel, v.trunc, 64.0)
        t_g, w_g = cu(v.tsdf), cu(v.weights)
        ops.updateTSDF(cu(scn.depths[f]), cu(assoc), t_g, w_g, T, scn.K, v.res, v.voxel, v.trunc, 64.0)
        torch.cuda.synchronize()
        assert_bits(t_r, t_o, f"reference vs oracle tsdf vol {v.vid}")
        assert_bits(w_r, w_o, f"reference vs oracle weights vol {v.vid}")
        assert_bits(t_g, t_r.cpu().numpy(), f"product vs reference tsdf vol {v.vid}")
        assert_bits(w_g, w_r.cpu().numpy(), f"product vs reference weights vol {v.vid}")


def test_gradients_three_way(scn, oracle, cuda_dev):
    for v in scn.vols():
        g_r = torch.full((v.n, 3), 7.0, device=DEV)
        ref_gpu.update_gradients(cu(v.tsdf), g_r, v.res)
        g_g = torch.full((v.n, 3), 7.0, device=DEV)
        ops.computeTSDFGrads(cu(v.tsdf), g_g, v.res)
        torch.cuda.synchronize()
        assert_bits(g_r, oracle.compute_grads(v.tsdf, v.res), "reference vs oracle grads")
        assert_bits(g_g, g_r.cpu().numpy(), "product vs reference grads")


def test_raycast_three_way(scn, oracle, cuda_dev):
    for v in scn.vols():
        T = rel_pose_CO(scn.cam(2), v.pose)
        grads = oracle.compute_grads(v.tsdf, v.res)
        w_eff = v.weights
        if v.fg_probs is not None:
            w_eff = oracle.raycast_weights(v.weights, (v.fg_probs > 0.5).astype(np.uint8) * 255)
        o = oracle.raycast(v.tsdf, grads, w_eff, S.R9(T), S.T3(T), scn.K, v.res, v.voxel, v.trunc, scn.w, scn.h)
        z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=DEV)
        r = dict(ray=z(scn.h, scn.w), vert=z(scn.h, scn.w, 3), norm=z(scn.h, scn.w, 3), mask=z(scn.h, scn.w, dt=torch.uint8))
        ref_gpu.raycast(cu(v.tsdf), cu(grads), cu(w_eff), r["ray"], r["vert"], r["norm"], r["mask"], S.R9(T), S.T3(T),
                        scn.K, v.res, v.voxel, v.trunc)
        g = dict(ray=z(scn.h, scn.w), vert=z(scn.h, scn.w, 3), norm=z(scn.h, scn.w, 3), mask=z(scn.h, scn.w, dt=torch.uint8))
        hit = torch.full((scn.h, scn.w, 3), -1, dtype=torch.int32, device=DEV)
        ops.raycastTSDF(cu(v.tsdf), None, cu(v.weights), g["ray"], g["vert"], g["norm"], g["mask"], T, scn.K, v.res,
                        v.voxel, v.trunc, fgProbs=cu(v.fg_probs) if v.fg_probs is not None else None, hit_voxel=hit)
        torch.cuda.synchronize()
        for k in ("mask", "ray", "vert", "norm"):
            assert_bits(r[k], o[k], f"reference vs oracle {k} vol {v.vid}")
            assert_bits(g[k], r[k].cpu().numpy(), f"product vs reference {k} vol {v.vid}")
        assert_bits(hit, o["hit"], f"product voxel index vs oracle vol {v.vid}")


def test_gather_three_way(scn, oracle, cuda_dev):
    pts_np = oracle.compute_points(scn.depths[2], scn.K)
    pts = cu(pts_np)
    for v in scn.vols():
        T = rel_pose_CO(scn.cam(2), v.pose)
        ref, _ = oracle.get_volume_vals(v.tsdf, pts_np, S.R9(T), S.T3(T), v.res, v.voxel)
        v_r = torch.full((scn.h, scn.w), 3.0, device=DEV)
        ref_gpu.get_volume_vals(cu(v.tsdf), pts, S.R9(T), S.T3(T), v.res, v.voxel, v_r)
        v_g = torch.full((scn.h, scn.w), 3.0, device=DEV)
        ops.getVolumeVals(cu(v.tsdf), pts, T, v.res, v.voxel, v_g)
        torch.cuda.synchronize()
        assert_bits(v_r, ref, "reference vs oracle gather")
        assert_bits(v_g, v_r.cpu().numpy(), "product vs reference gather")


def test_fgbg_three_way(scn, oracle, cuda_dev):
    occl = (np.random.default_rng(3).random((scn.h, scn.w)) < 0.2).astype(np.uint8)
    for v in scn.objs:
        T = rel_pose_OC(scn.cam(2), v.pose)
        m = (scn.insts[2] == v.vid).astype(np.uint8)
        fo = v.fgbg.copy()
        oracle.update_fgbg(m, occl, v.tsdf, v.weights, fo, S.R9(T), S.T3(T), scn.K, v.res, v.voxel)
        fr = cu(v.fgbg)
        ref_gpu.update_fgbg(cu(m), cu(occl), cu(v.tsdf), cu(v.weights), fr, S.R9(T), S.T3(T), scn.K, v.res, v.voxel)
        fg = cu(v.fgbg)
        ops.updateFgBgProbs(cu(m), cu(occl), cu(v.tsdf), cu(v.weights), fg, T, scn.K, v.res, v.voxel)
        torch.cuda.synchronize()
        assert_bits(fr, fo, "reference vs oracle fgbg")
        assert_bits(fg, fr.cpu().numpy(), "product vs reference fgbg")
