"""CPU: host-side logic of the mirror -- pose algebra (the reference composes relative poses on the host,
src/core/TSDF.cpp:112,141,162), parameter defaults (include/EMFusion/core/data.h), the synthetic stream, and
the C oracle's own invariants on config 1 of BASELINE.json (64^3 background, one 640x480 frame, CPU only)."""
import numpy as np
import pytest

from emfusion_b200.poses import Affine, rel_pose_CO, rel_pose_OC
from emfusion_b200.synth import Scene
from emfusion_b200.volume import Params, TSDFParams
from tests import scenario as S


def test_affine_algebra():
    rng = np.random.default_rng(0)
    a = Affine.from_rvec(rng.normal(size=3), rng.normal(size=3))
    b = Affine.from_rvec(rng.normal(size=3), rng.normal(size=3))
    i = a * a.inv()
    assert np.allclose(i.R, np.eye(3), atol=1e-12) and np.allclose(i.t, 0, atol=1e-12)
    assert np.allclose(np.linalg.det(a.R), 1.0)
    x = rng.normal(size=3)
    assert np.allclose((a * b).R @ x + (a * b).t, a.R @ (b.R @ x + b.t) + a.t)
    # T_OC and T_CO are mutual inverses
    oc, co = rel_pose_OC(a, b), rel_pose_CO(a, b)
    j = oc * co
    assert np.allclose(j.R, np.eye(3), atol=1e-12) and np.allclose(j.t, 0, atol=1e-12)
    # volume -> camera: a point at the volume origin lands at cam^-1 * pose.t
    assert np.allclose(oc.t, a.inv().R @ b.t + a.inv().t)
    assert a.rotation32().dtype == np.float32 and a.rotation32().shape == (9,)
    # Rodrigues: rotation by pi/2 about z maps x to y
    r = Affine.from_rvec([0, 0, np.pi / 2])
    assert np.allclose(r.R @ [1, 0, 0], [0, 1, 0])
    assert np.allclose(Affine.from_rvec([0, 0, 0]).R, np.eye(3))


def test_param_defaults_follow_the_reference():
    p = Params()
    t = TSDFParams()
    # include/EMFusion/core/data.h:37-46 and :81-122
    assert (t.maxTSDFWeight, t.assocSigma, t.alpha, t.uniPrior) == (64.0, 0.02, 0.8, 1.0)
    assert p.frameSize == (640, 480) and p.visibilityThresh == 1600 and p.boundary == 20 and p.volPad == 2.0
    assert np.allclose(p.intr, [[525, 0, 319.5], [0, 525, 239.5], [0, 0, 1]])
    assert p.globalRelTruncDist == 10.0 and p.objRelTruncDist == 10.0
    assert np.allclose(p.volumePose.t, [0, 0, 2.56])


def test_scene_is_deterministic_and_sane():
    a = Scene(n_objects=5, width=160, height=120, seed=4, dropout=0.02, noise_sigma=0.002)
    b = Scene(n_objects=5, width=160, height=120, seed=4, dropout=0.02, noise_sigma=0.002)
    da, ia = a.render(3)
    db, ib = b.render(3)
    assert np.array_equal(da, db) and np.array_equal(ia, ib)
    assert da.dtype == np.float32 and ia.dtype == np.uint8
    assert 0.005 < (da == 0).mean() < 0.05                 # dropouts
    assert da.max() < 6.0 and da[da > 0].min() > 0.5
    assert set(np.unique(ia)) <= set(range(6)) and len(np.unique(ia)) >= 3
    # a sphere pixel is nearer than the room behind it
    clean = Scene(n_objects=5, width=160, height=120, seed=4)
    room = Scene(n_objects=0, width=160, height=120, seed=4)
    dc, ic = clean.render(0)
    dr, _ = room.render(0)
    assert (dc[ic > 0] < dr[ic > 0]).all() and np.array_equal(dc[ic == 0], dr[ic == 0])
    # object volumes: metric edge = volPad * diameter (src/core/EMFusion.cpp:537-547)
    assert np.isclose(clean.object_voxel_size(0, 64) * 64, 2.0 * 2 * clean.spheres[0].radius, rtol=1e-6)


@pytest.fixture(scope="module")
def cfg1(oracle):
    """BASELINE.json configs[0]: single 64^3 background, one synthetic 640x480 frame, integrate + raycast on CPU"""
    return S.make("cfg1", oracle, 640, 480, (64, 64, 64), 0, (32, 32, 32), n_frames=2, integrate_frames=1)


def test_config1_integrate_counts(cfg1, oracle):
    v = cfg1.bg
    t, w = np.zeros_like(v.tsdf), np.zeros_like(v.weights)
    T = rel_pose_OC(cfg1.cam(0), v.pose)
    ones = np.ones((480, 640), np.float32)
    c = oracle.update_tsdf(cfg1.depths[0], ones, t, w, S.R9(T), S.T3(T), cfg1.K, v.res, v.voxel, v.trunc, 64.0,
                           counts=True)
    assert int(c.sum()) == v.n
    frac = c / v.n
    # SURVEY Appendix C probe: ~35 % of the cube projects into the image, ~20 % updated, ~15 % marked occluded
    assert 0.25 < frac[0] + frac[1] + frac[2] + frac[3] < 0.45
    assert 0.12 < frac[0] < 0.40 and 0.03 < frac[1] < 0.25
    assert np.array_equal(t, v.tsdf) and np.array_equal(w, v.weights)
    assert set(np.unique(w)) == {0.0, 1.0}
    assert t.min() == -1.0 and t.max() == 1.0
    # idempotent branch: integrating an all-invalid depth frame changes nothing that was seen and only
    # un-marks never-seen voxels (TSDF.cu:369-374)
    t2, w2 = t.copy(), w.copy()
    oracle.update_tsdf(np.zeros((480, 640), np.float32), ones, t2, w2, S.R9(T), S.T3(T), cfg1.K, v.res, v.voxel,
                       v.trunc, 64.0)
    assert np.array_equal(w2, w)
    seen = w > 0
    assert np.array_equal(t2[seen], t[seen])
    assert (t2[~seen] <= 0).all()


def test_config1_raycast_reproduces_the_depth_with_the_reference_bias(cfg1, oracle):
    """The raycast of the just-integrated frame recovers the input depth up to the reference's one-step overshoot
    (t* uses the already-advanced sample: SURVEY A.4 -- median error about half a voxel, not zero)."""
    v = cfg1.bg
    T = rel_pose_CO(cfg1.cam(0), v.pose)
    g = oracle.compute_grads(v.tsdf, v.res)
    r = oracle.raycast(v.tsdf, g, v.weights, S.R9(T), S.T3(T), cfg1.K, v.res, v.voxel, v.trunc, 640, 480)
    m = r["mask"].astype(bool)
    assert m.mean() > 0.9
    dz = np.abs(r["vert"][..., 2][m] - cfg1.depths[0][m])
    assert 0.2 * v.voxel < np.median(dz) < 0.8 * v.voxel
    n = np.linalg.norm(r["norm"][m], axis=-1)
    assert np.allclose(n[np.isfinite(n)], 1.0, atol=1e-4)
    assert (r["ray"][~m] == 0).all() and (r["ray"][m] > 0).all()
    assert r["steps"][0] > 640 * 480 * 10          # march steps are counted (roofline numerator)
    hit = r["hit"][m]
    assert (hit >= 0).all() and (hit[:, 0] < 63).all() and (hit[:, 1] < 63).all() and (hit[:, 2] < 63).all()


def test_association_normaliser_properties(oracle):
    sc = S.make("assoc", oracle, 160, 120, (64, 64, 64), 3, (32, 32, 32), n_frames=2, integrate_frames=1, seed=2)
    pts = oracle.compute_points(sc.depths[1], sc.K)
    imgs = []
    for v in sc.vols():
        T = rel_pose_CO(sc.cam(1), v.pose)
        a, m = oracle.assoc_volume(v.tsdf, v.fg_probs, pts, S.R9(T), S.T3(T), v.res, v.voxel, v.trunc)
        assert ((a == 0) == (m != 0)).all()                   # w = 0 exactly where the gather returned 0
        nz = a[a != 0]
        assert nz.min() >= 0.2 - 1e-6 and nz.max() <= 0.8 / 0.04 + 0.2 + 1e-4   # (1-a)*uni .. a/(2 sigma) + (1-a)*uni
        imgs.append(a)
    raw = [i.copy() for i in imgs]
    norm = oracle.normalise(imgs)
    s = np.zeros_like(norm)
    for r in raw:
        s = s + r                                             # bg first, then ascending id, fp32
    assert np.array_equal(s, norm)
    tot = sum(imgs)
    assert np.allclose(tot[norm > 0], 1.0, atol=1e-5) and (tot[norm == 0] == 0).all()
    # x / 0 -> 0 (guarded cv::cuda::divide), IEEE NaN when the switch is off
    z = [np.zeros((2, 2), np.float32), np.zeros((2, 2), np.float32)]
    oracle.normalise(z)
    assert (z[0] == 0).all()
    z = [np.zeros((2, 2), np.float32), np.zeros((2, 2), np.float32)]
    oracle.normalise(z, div0_is_zero=False)
    assert np.isnan(z[0]).all()


def test_composite_rules(oracle):
    """src/core/EMFusion.cpp:760-794 on hand-made 1x4 images: strictly-nearer-or-first wins, background wins when
    more than 5 cm nearer, untouched pixels take the background."""
    h, w = 1, 4
    f = lambda *v: np.array([v], np.float32)
    ray1, ray2 = f(1.0, 2.0, 0.0, 1.5), f(1.0, 1.0, 3.0, 0.0)
    m1, m2 = np.array([[1, 1, 0, 1]], np.uint8), np.array([[1, 1, 1, 0]], np.uint8)
    v3 = lambda r, k: np.repeat(r[..., None], 3, -1) * k
    bg_ray, bg_mask = f(0.9, 0.96, 2.5, 0.0), np.array([[1, 1, 1, 0]], np.uint8)
    c = oracle.composite([5, 9], [ray1, ray2], [v3(ray1, 1), v3(ray2, 2)], [v3(ray1, -1), v3(ray2, -2)], [m1, m2],
                         bg_ray, v3(bg_ray, 7), v3(bg_ray, -7), bg_mask, 0)
    # px0: tie 1.0 vs 1.0 -> first in list (id 5); bg 0.9 is 0.1 nearer -> bg wins (seg 0)
    # px1: obj2 (1.0) nearer than obj1 (2.0) -> id 9; bg 0.96: 1.0 - 0.96 = 0.04 <= 0.05 -> stays 9
    # px2: only obj2 at 3.0; bg 2.5 nearer by 0.5 -> 0
    # px3: only obj1 at 1.5; no bg hit -> 5
    assert c["seg"].tolist() == [[0, 9, 0, 5]]
    assert c["ray"].tolist() == [[1.0, 1.0, 3.0, 1.5]]       # raylengths keep the object value (reference quirk)
    assert np.allclose(c["vert"][0, 1], 2.0) and np.allclose(c["vert"][0, 0], 0.9 * 7) and np.allclose(c["vert"][0, 3], 1.5)
    assert c["vis"].tolist() == [1, 1]
