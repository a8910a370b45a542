"""-m gpu parity tests: the CUDA path, called through the C ABI (emfusion_b200.ops), against
 (a) the C oracle (oracle/emf_oracle.c) -- bit-exact on integer/index outputs AND on every float that
     does not go through expf; 1e-5 on association weights (expf differs by <= 2 ulp between libm and CUDA);
 (b) the reference's own kernels (oracle/_ref, compiled unchanged) -- same bars (tests/test_gpu_vs_reference.py).
BASELINE.json's bar is: raycast voxel indices bit-exact, TSDF/association floats within 1e-4."""
import numpy as np
import pytest
import torch

from emfusion_b200 import ops
from emfusion_b200.poses import Affine, rel_pose_CO, rel_pose_OC
from tests import scenario as S

pytestmark = pytest.mark.gpu

ASSOC_TOL = 1e-5   # un-normalised weights are <= 20.2; normalised ones <= 1
DEV = "cuda"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def assert_bits(a: torch.Tensor, b: np.ndarray, what: str):
    a = a.detach().cpu().numpy().reshape(-1)
    b = np.asarray(b).reshape(-1)
    if a.dtype == np.float32:
        same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b)) | ((a == 0) & (b == 0))
    else:
        same = a == b
    assert same.all(), f"{what}: {int((~same).sum())} of {a.size} differ; first at {np.flatnonzero(~same)[:5]}"


SCENARIOS = [
    # name, w, h, bg_res, n_obj, obj_res, kwargs
    ("cfg1_64", 640, 480, (64, 64, 64), 0, (32, 32, 32), {}),
    ("small_objs", 160, 120, (64, 64, 64), 3, (32, 32, 32), {}),
    ("ragged", 160, 120, (50, 42, 46), 2, (22, 26, 30), dict(dropout=0.05)),   # rx % 4 != 0 -> scalar path
    ("noisy", 320, 240, (96, 96, 96), 2, (48, 48, 48), dict(noise=0.002, dropout=0.02)),
]


@pytest.fixture(scope="module", params=SCENARIOS, ids=[s[0] for s in SCENARIOS])
def scn(request, oracle):
    name, w, h, bg_res, n_obj, obj_res, kw = request.param
    return S.make(name, oracle, w, h, bg_res, n_obj, obj_res, n_frames=3, integrate_frames=2, **kw)


def test_compute_points(scn, oracle, cuda_dev):
    d = cu(scn.depths[2])
    pts = torch.zeros((scn.h, scn.w, 3), device=DEV)
    ops.computePoints(d, pts, scn.K)
    assert_bits(pts, oracle.compute_points(scn.depths[2], scn.K), "points")


def test_integrate_bit_exact(scn, oracle, cuda_dev):
    """one more frame on top of the oracle's state, with a non-trivial association image"""
    f = 2
    rng = np.random.default_rng(5)
    assoc = rng.random((scn.h, scn.w), dtype=np.float32)
    assoc[rng.random((scn.h, scn.w)) < 0.1] = 0.0
    for v in scn.vols():
        pose = v.pose if v.vid == 0 else scn.scene.object_pose(v.vid - 1, f)
        T = rel_pose_OC(scn.cam(f), pose)
        t_o, w_o = v.tsdf.copy(), v.weights.copy()
        oracle.update_tsdf(scn.depths[f], assoc, t_o, w_o, S.R9(T), S.T3(T), scn.K, v.res, v.voxel, v.trunc, 64.0)
        t_g, w_g = cu(v.tsdf), cu(v.weights)
        ops.updateTSDF(cu(scn.depths[f]), cu(assoc), t_g, w_g, T, scn.K, v.res, v.voxel, v.trunc, 64.0)
        assert_bits(t_g, t_o, f"tsdf vol {v.vid}")
        assert_bits(w_g, w_o, f"weights vol {v.vid}")


def test_integrate_behind_camera(scn, oracle, cuda_dev):
    """camera turned around: every voxel takes the pc.z <= 0 / invalid-depth branch (tsdf reset where unseen)"""
    v = scn.bg
    cam = Affine.from_rvec([0, np.pi * 0.9, 0], [0.3, 0.1, 2.0])
    T = rel_pose_OC(cam, v.pose)
    assoc = np.ones((scn.h, scn.w), np.float32)
    t_o, w_o = v.tsdf.copy(), v.weights.copy()
    oracle.update_tsdf(scn.depths[0], assoc, t_o, w_o, S.R9(T), S.T3(T), scn.K, v.res, v.voxel, v.trunc, 64.0)
    t_g, w_g = cu(v.tsdf), cu(v.weights)
    ops.updateTSDF(cu(scn.depths[0]), cu(assoc), t_g, w_g, T, scn.K, v.res, v.voxel, v.trunc, 64.0)
    assert_bits(t_g, t_o, "tsdf")
    assert_bits(w_g, w_o, "weights")


def test_gradients_bit_exact(scn, oracle, cuda_dev):
    for v in scn.vols():
        g = torch.full((v.n, 3), 7.0, device=DEV)
        ops.computeTSDFGrads(cu(v.tsdf), g, v.res)
        assert_bits(g, oracle.compute_grads(v.tsdf, v.res), f"grads vol {v.vid}")


def _raycast_both(scn, oracle, v, f, use_grad_volume, far_clip=None):
    pose = v.pose
    T = rel_pose_CO(scn.cam(f), pose)
    grads = oracle.compute_grads(v.tsdf, v.res)
    w_eff = v.weights
    if v.fg_probs is not None:
        w_eff = oracle.raycast_weights(v.weights, (v.fg_probs > 0.5).astype(np.uint8) * 255)
    o = oracle.raycast(v.tsdf, grads, w_eff, S.R9(T), S.T3(T), scn.K, v.res, v.voxel, v.trunc, scn.w, scn.h,
                       raylengths=far_clip)
    ray = cu(far_clip) if far_clip is not None else torch.zeros((scn.h, scn.w), device=DEV)
    vert = torch.zeros((scn.h, scn.w, 3), device=DEV)
    norm = torch.zeros((scn.h, scn.w, 3), device=DEV)
    mask = torch.zeros((scn.h, scn.w), dtype=torch.uint8, device=DEV)
    hit = torch.full((scn.h, scn.w, 3), -1, dtype=torch.int32, device=DEV)
    ops.raycastTSDF(cu(v.tsdf), cu(grads) if use_grad_volume else None, cu(v.weights), ray, vert, norm, mask, T, scn.K,
                    v.res, v.voxel, v.trunc, fgProbs=cu(v.fg_probs) if v.fg_probs is not None else None, hit_voxel=hit)
    return o, dict(ray=ray, vert=vert, norm=norm, mask=mask, hit=hit)


@pytest.mark.parametrize("use_grad_volume", [False, True], ids=["grad_on_the_fly", "grad_volume"])
def test_raycast_bit_exact(scn, oracle, cuda_dev, use_grad_volume):
    for v in scn.vols():
        o, g = _raycast_both(scn, oracle, v, 2, use_grad_volume)
        assert o["mask"].sum() > 0 or v.vid > 0
        assert_bits(g["mask"], o["mask"], f"hit mask vol {v.vid}")
        assert_bits(g["hit"], o["hit"], f"voxel index vol {v.vid}")
        assert_bits(g["ray"], o["ray"], f"raylength vol {v.vid}")
        assert_bits(g["vert"], o["vert"], f"vertex vol {v.vid}")
        assert_bits(g["norm"], o["norm"], f"normal vol {v.vid}")


def test_raycast_far_clip(scn, oracle, cuda_dev):
    """non-zero incoming raylengths act as a far clip (reference TSDF.cu:496-500)"""
    clip = np.full((scn.h, scn.w), 2.0, np.float32)
    clip[:, : scn.w // 2] = 0.0
    o, g = _raycast_both(scn, oracle, scn.bg, 2, False, far_clip=clip)
    assert_bits(g["mask"], o["mask"], "mask")
    assert_bits(g["ray"], o["ray"], "ray")
    assert_bits(g["hit"], o["hit"], "hit")


def test_gather_bit_exact(scn, oracle, cuda_dev):
    pts_np = oracle.compute_points(scn.depths[2], scn.K)
    pts = cu(pts_np)
    for v in scn.vols():
        T = rel_pose_CO(scn.cam(2), v.pose)
        vals = torch.full((scn.h, scn.w), 3.0, device=DEV)
        ops.getVolumeVals(cu(v.tsdf), pts, T, v.res, v.voxel, vals)
        ref, nin = oracle.get_volume_vals(v.tsdf, pts_np, S.R9(T), S.T3(T), v.res, v.voxel)
        assert_bits(vals, ref, f"gather vol {v.vid}")


def test_association_single_volume(scn, oracle, cuda_dev):
    pts_np = oracle.compute_points(scn.depths[2], scn.K)
    pts = cu(pts_np)
    prm = ops.tsdf_params()
    for v in scn.vols():
        T = rel_pose_CO(scn.cam(2), v.pose)
        ref, rmask = oracle.assoc_volume(v.tsdf, v.fg_probs, pts_np, S.R9(T), S.T3(T), v.res, v.voxel, v.trunc)
        out = torch.full((scn.h, scn.w), -1.0, device=DEV)
        m = torch.zeros((scn.h, scn.w), dtype=torch.uint8, device=DEV)
        vol = ops.volume(cu(v.tsdf), cu(v.weights), v.res, v.voxel, v.trunc,
                         fg_probs=cu(v.fg_probs) if v.fg_probs is not None else None, vid=v.vid)
        ops.computeAssociation(vol, pts, T, prm, out, m)
        assert_bits(m, rmask, f"associationMask vol {v.vid}")
        err = np.abs(out.cpu().numpy() - ref).max()
        assert err <= ASSOC_TOL * 20.2, f"assoc vol {v.vid}: L_inf {err}"


def test_fg_probs_bit_exact(scn, oracle, cuda_dev):
    f = 2
    zeros = np.zeros((scn.h, scn.w), np.uint8)
    occl = (np.random.default_rng(3).random((scn.h, scn.w)) < 0.2).astype(np.uint8)
    for v in scn.objs:
        T = rel_pose_OC(scn.cam(f), v.pose)
        m = (scn.insts[f] == v.vid).astype(np.uint8)
        fgbg_o = v.fgbg.copy()
        oracle.update_fgbg(m, occl, v.tsdf, v.weights, fgbg_o, S.R9(T), S.T3(T), scn.K, v.res, v.voxel)
        p_o, vm_o = oracle.compute_fg_probs(fgbg_o)
        fgbg_g = cu(v.fgbg)
        ops.updateFgBgProbs(cu(m), cu(occl), cu(v.tsdf), cu(v.weights), fgbg_g, T, scn.K, v.res, v.voxel)
        p_g = torch.zeros(v.n, device=DEV)
        vm_g = torch.zeros(v.n, dtype=torch.uint8, device=DEV)
        ops.computeFgProbs(fgbg_g, p_g, vm_g)
        assert_bits(fgbg_g, fgbg_o, "fgbg")
        assert_bits(p_g, p_o, "fgProbs")
        assert_bits(vm_g, vm_o, "fgVolMask")
