"""-m gpu: the ray-space certificate (csrc/raycast.cu: k_ray_certify + march_cert) must never change a result.
 * raycast with the certificate == raycast without it, bit for bit (hit mask, t*, vertices, normals), for cameras
   looking along every volume axis in both directions, rolled, oblique, inside the volume, and for ragged
   resolutions, sub-rectangles and an in/out far clip is not involved (batched semantics);
 * the certificate is not vacuous: most of the background's march samples are skipped;
 * against the C oracle on a GPU-integrated volume."""
import numpy as np
import pytest
import torch

from emfusion_b200 import ops
from emfusion_b200.poses import Affine, rel_pose_CO, rel_pose_OC
from emfusion_b200.synth import Scene
from tests import scenario as S
from tests.test_gpu_parity import DEV, assert_bits, cu

pytestmark = pytest.mark.gpu


def integrated_volume(res, scene, frames, w, h, pose, cams=None):
    n = int(np.prod(res))
    voxel = float(np.float32(5.12 / max(res)))
    trunc = float(np.float32(10.0) * np.float32(voxel))
    t_g, w_g = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    v = ops.volume(t_g, w_g, res, voxel, trunc)
    ones = torch.ones((h, w), device=DEV)
    for f in range(frames):
        depth, _ = scene.render(f)
        cam = scene.cam_pose(f) if cams is None else cams[f]
        ops.integrateVolumes([v], [rel_pose_OC(cam, pose)], scene.K, cu(depth), [ones], 64.0)
    return v, t_g, w_g, voxel, trunc


def cast(v, T, K, w, h, rect, ws, stats=None):
    z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=DEV)
    ray, vert, norm, mask = [z(h, w)], [z(h, w, 3)], [z(h, w, 3)], [z(h, w, dt=torch.uint8)]
    ops.raycastVolumes([v], [T], K, [rect], ray, vert, norm, mask, stats=stats, workspace=ws)
    return ray[0], vert[0], norm[0], mask[0]


def view_poses():
    """cameras looking at the volume centre (0, 0, 2.56) along +-z, +-x, +-y, diagonally, rolled, from inside"""
    c = np.array([0.0, 0.0, 2.56])
    out = [("+z", Affine.identity()),
           ("-z", Affine.from_rvec([0, np.pi, 0], [0, 0, 5.2])),
           ("+x", Affine.from_rvec([0, np.pi / 2, 0], [-2.7, 0, 2.56])),
           ("-x", Affine.from_rvec([0, -np.pi / 2, 0], [2.7, 0.1, 2.5])),
           ("+y", Affine.from_rvec([-np.pi / 2, 0, 0], [0, -2.7, 2.56])),
           ("-y", Affine.from_rvec([np.pi / 2, 0, 0], [0.1, 2.7, 2.6])),
           ("diag", Affine.from_rvec([0, np.pi / 4, 0], [-2.0, 0, 0.5])),
           ("diag3", Affine.from_rvec(np.array([-0.6, 0.7, 0.2]), [-1.6, -1.4, 0.6])),
           ("rolled", Affine.from_rvec([0, 0, np.pi / 3], [0.2, -0.1, -0.4])),
           ("inside", Affine.from_rvec([0.1, -0.2, 0.05], [0.3, -0.2, 1.5])),
           ("close", Affine.from_rvec([0.02, 0.01, 0.0], [-0.05, 0.03, 0.2]))]
    return out


@pytest.mark.parametrize("res", [(128, 128, 128), (96, 64, 112), (70, 50, 66)], ids=["128", "96x64x112", "ragged"])
def test_certificate_changes_nothing(cuda_dev, res):
    w, h = 320, 240
    scene = Scene(n_objects=3, width=w, height=h, seed=11, dropout=0.01)
    pose = Affine.translation([0, 0, 2.56])
    # integrate from several sides so that every view direction below sees free space
    names, cams = zip(*view_poses())
    v, t_g, w_g, voxel, trunc = integrated_volume(res, scene, len(cams), w, h, pose, cams=list(cams))
    ws = ops.raycastWorkspace(w, h, DEV)
    skipped_total = 0
    for name, cam in view_poses() + [("stream%d" % f, scene.cam_pose(f)) for f in (3, 17)]:
        T = rel_pose_CO(cam, pose)
        for rect in ([0, 0, w, h], [37, 21, 251, 199]):
            st0 = torch.zeros(8, dtype=torch.int64, device=DEV)
            st1 = torch.zeros(8, dtype=torch.int64, device=DEV)
            a = cast(v, T, scene.K, w, h, rect, None, st0)
            b = cast(v, T, scene.K, w, h, rect, ws, st1)
            for x, y, what in zip(a, b, ("ray", "vert", "norm", "mask")):
                assert_bits(y, x.cpu().numpy(), f"{name} rect {rect} {what}")
            s0, s1 = st0.cpu().numpy(), st1.cpu().numpy()
            # every sample is either taken or skipped
            assert s1[0] + s1[1] == s0[0], (name, s0, s1)
            skipped_total += int(s1[1])
    assert skipped_total > 0


def test_certificate_skips_most_free_space(oracle, cuda_dev):
    """a room-sized 256^3 background seen from its face (the bench geometry): the certificate removes most samples and
    the result equals the C oracle's"""
    w, h = 320, 240
    res = (256, 256, 256)
    scene = Scene(n_objects=6, width=w, height=h, seed=0)
    pose = Affine.translation([0, 0, 2.56])
    v, t_g, w_g, voxel, trunc = integrated_volume(res, scene, 6, w, h, pose)
    ws = ops.raycastWorkspace(w, h, DEV)
    T = rel_pose_CO(scene.cam_pose(6), pose)
    st = torch.zeros(8, dtype=torch.int64, device=DEV)
    ray, vert, norm, mask = cast(v, T, scene.K, w, h, [0, 0, w, h], ws, st)
    s = st.cpu().numpy()
    assert s[1] > s[0], s              # more than half of the samples are skipped
    t_np, w_np = t_g.cpu().numpy(), w_g.cpu().numpy()
    o = oracle.raycast(t_np, oracle.compute_grads(t_np, res), w_np, S.R9(T), S.T3(T), scene.K, res, voxel, trunc, w, h)
    assert_bits(mask, o["mask"], "mask")
    assert_bits(ray, o["ray"], "ray")
    assert_bits(vert, o["vert"], "vert")
    assert_bits(norm, o["norm"], "norm")
    assert int(o["steps"][0]) == int(s[0] + s[1])
    assert int(o["mask"].sum()) > 0.5 * w * h


def test_schedule_changes_nothing(cuda_dev):
    """longest-first tile order (k_ray_schedule): the second and third call run in the order the previous one recorded; every
    output of every volume stays bit-identical to the row-major launch"""
    w, h = 320, 240
    scene = Scene(n_objects=3, width=w, height=h, seed=5)
    pose = Affine.translation([0, 0, 2.56])
    v, t_g, w_g, voxel, trunc = integrated_volume((128, 128, 128), scene, 4, w, h, pose)
    v2, *_ = integrated_volume((96, 64, 80), scene, 3, w, h, pose)
    ws = ops.raycastWorkspace(w, h, DEV)
    z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=DEV)
    for f in (4, 5, 6):
        T = rel_pose_CO(scene.cam_pose(f), pose)
        rects = [[0, 0, w, h], [40, 30, 290, 200]]
        outs = []
        for wk in (None, ws):
            ray, vert, norm, mask = [z(h, w), z(h, w)], [z(h, w, 3), z(h, w, 3)], [z(h, w, 3), z(h, w, 3)], [z(h, w, dt=torch.uint8), z(h, w, dt=torch.uint8)]
            ops.raycastVolumes([v, v2], [T, T], scene.K, rects, ray, vert, norm, mask, workspace=wk, certificate=False)
            outs.append(ray + vert + norm + mask)
        for a, b in zip(*outs):
            assert_bits(b, a.cpu().numpy(), f"frame {f}")
        hdr = ws[-(4 + 2 * 20 * 30) * 4:].view(torch.int32)[:4].cpu().numpy()
        assert hdr[0] == 20 * 30       # the order for the next call is in place
        order = ws[-(4 + 2 * 20 * 30) * 4:].view(torch.int32)[4:4 + 600].cpu().numpy()
        assert np.array_equal(np.sort(order), np.arange(600))


def test_wide_march_changes_nothing(cuda_dev):
    """EMF_RAY_WIDE: four lanes per background ray, no certificate -- same bits as one ray per lane"""
    w, h = 320, 240
    scene = Scene(n_objects=3, width=w, height=h, seed=7)
    pose = Affine.translation([0, 0, 2.56])
    v, *_ = integrated_volume((128, 128, 128), scene, 5, w, h, pose)
    z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=DEV)
    for f in (5, 9):
        T = rel_pose_CO(scene.cam_pose(f), pose)
        for rect in ([0, 0, w, h], [33, 17, 300, 201]):
            a = cast(v, T, scene.K, w, h, rect, None)
            ray, vert, norm, mask = [z(h, w)], [z(h, w, 3)], [z(h, w, 3)], [z(h, w, dt=torch.uint8)]
            ops.raycastVolumes([v], [T], scene.K, [rect], ray, vert, norm, mask, certificate=False, schedule=False, wide=True)
            for x, y, what in zip(a, (ray[0], vert[0], norm[0], mask[0]), ("ray", "vert", "norm", "mask")):
                assert_bits(y, x.cpu().numpy(), f"frame {f} rect {rect} {what}")
