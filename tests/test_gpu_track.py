"""-m gpu: the tracker row.  emf_track_linearise through the C ABI against
 (a) the C oracle's op-by-op restatement of one tracker iteration (per-pixel images bit-exact, sums to 1e-5 relative),
 (b) the reference's own kernels and launch chain (oracle/_ref: computePoseGradients, getVolumeVals, computeAb,
     multSingletonCol compiled unchanged + the OpenCV ops between them) -- `grads` bit-exact, A / b / error to 2e-4 of the
     largest entry (the reference sums 307 200 floats per entry in float, in an order OpenCV does not specify),
 (c) the whole Levenberg-Marquardt loop: a perturbed camera pose is pulled back onto the volume."""
import numpy as np
import pytest
import torch

from emfusion_b200 import ops
from emfusion_b200.poses import Affine, rel_pose_CO
from emfusion_b200.tracking import Tracker, se3_exp, se3_log
from emfusion_b200.volume import TSDF, ObjTSDF, TSDFParams
from tests import ref_gpu, scenario as S
from tests.test_gpu_parity import DEV, SCENARIOS, assert_bits, cu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=SCENARIOS, ids=[s[0] for s in SCENARIOS])
def scn(request, oracle):
    name, w, h, bg_res, n_obj, obj_res, kw = request.param
    return S.make(name, oracle, w, h, bg_res, n_obj, obj_res, n_frames=3, integrate_frames=2, **kw)


def _perturbed(T: Affine, k: int) -> Affine:
    """a relative pose a few millimetres / tenths of a degree off (what the tracker sees between frames)"""
    d = np.array([0.004, -0.003, 0.005, 0.003, -0.002, 0.004]) * (1 + 0.3 * k)
    return se3_exp(d) * T


def _run_ours(v, T, pts, assoc, grads_vol=None, mode=1, rec=None, iw=None, intr=None):
    n_pix = pts.shape[0] * pts.shape[1]
    t_d, w_d = cu(v.tsdf), cu(v.weights)     # (kept alive below: the descriptor only holds raw pointers)
    cv = ops.volume(t_d, w_d, v.res, v.voxel, v.trunc, grads=grads_vol, vid=v.vid)
    rec = torch.zeros((1, 48), device=DEV) if rec is None else rec
    iw = torch.zeros(pts.shape[:2], device=DEV) if iw is None else iw
    vals = torch.full(pts.shape[:2], 7.0, device=DEV)
    tw = torch.full(pts.shape[:2], 7.0, device=DEV)
    g6 = torch.full((n_pix, 6), 7.0, device=DEV)
    ops.trackLinearise([cv], [T], [mode], pts, [assoc], 0.2, 64.0, [iw], rec, tsdfVals=[vals], trackWeights=[tw],
                       poseGrads=[g6], intr=intr)
    torch.cuda.synchronize()
    return dict(rec=rec, intWeights=iw, tsdfVals=vals, trackWeights=tw, grads=g6, keep=(cv, t_d, w_d, grads_vol))


def test_linearise_vs_oracle(scn, oracle, cuda_dev):
    rng = np.random.default_rng(11)
    pts_h = oracle.compute_points(scn.depths[2], scn.K)
    pts = cu(pts_h)
    for k, v in enumerate(scn.vols()):
        T = _perturbed(rel_pose_CO(scn.cam(2), v.pose), k)
        assoc = rng.random((scn.h, scn.w), dtype=np.float32)
        assoc[rng.random((scn.h, scn.w)) < 0.1] = 0.0
        gvol = oracle.compute_grads(v.tsdf, v.res)
        o = oracle.track_linearise(v.tsdf, gvol, v.weights, pts_h, assoc, S.R9(T), S.T3(T), v.res, v.voxel)
        for use_vol in (False, True):   # gradients on the fly / from the materialised float3 volume
            g = _run_ours(v, T, pts, cu(assoc), grads_vol=cu(gvol) if use_vol else None)
            assert_bits(g["grads"], o["grads"], f"vol {v.vid} grads (volume={use_vol})")
            assert_bits(g["tsdfVals"], o["tsdfVals"], f"vol {v.vid} tsdfVals")
            assert_bits(g["trackWeights"], o["trackWeights"], f"vol {v.vid} trackWeights")
            rec = g["rec"].cpu().numpy()[0].astype(np.float64)
            # intWeights: ours are held before the NORM_INF scale
            iw = ops_normalised(g)
            np.testing.assert_allclose(iw, o["intWeights"], rtol=6e-7, atol=1e-12)
            sA = max(np.abs(o["A"]).max(), 1e-30)
            assert np.abs(rec[:36].reshape(6, 6) - o["A"]).max() <= 1e-5 * sA, f"vol {v.vid} A"
            assert np.abs(rec[36:42] - o["b"]).max() <= 1e-5 * max(np.abs(o["b"]).max(), 1e-30), f"vol {v.vid} b"
            assert abs(rec[42] - o["err"]) <= 1e-5 * max(abs(o["err"]), 1e-30), f"vol {v.vid} error"
            assert rec[43] == np.float32(o["wmax"])
            A = rec[:36].reshape(6, 6)
            assert np.array_equal(A, A.T)


def ops_normalised(g):
    out = torch.empty_like(g["intWeights"])
    ops.trackNormalisedWeights(g["intWeights"], g["rec"][0], out)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_error_only_mode(scn, oracle, cuda_dev):
    """mode 2 = computeTSDFVals at a trial pose + computeError with the weights of the last linearisation"""
    pts_h = oracle.compute_points(scn.depths[2], scn.K)
    pts = cu(pts_h)
    v = scn.bg
    T = _perturbed(rel_pose_CO(scn.cam(2), v.pose), 0)
    assoc = np.ones((scn.h, scn.w), np.float32)
    g = _run_ours(v, T, pts, cu(assoc))
    A0 = g["rec"].cpu().numpy()[0, :42].copy()
    T2 = _perturbed(T, 1)
    g2 = _run_ours(v, T2, pts, cu(assoc), mode=2, rec=g["rec"], iw=g["intWeights"])
    vals_o, _ = oracle.get_volume_vals(v.tsdf, pts_h, S.R9(T2), S.T3(T2), v.res, v.voxel)
    assert_bits(g2["tsdfVals"], vals_o, "trial tsdfVals")
    w_n = ops_normalised(g).astype(np.float64)
    want = float(((vals_o.astype(np.float32) * vals_o).astype(np.float64) * w_n).sum())
    rec = g2["rec"].cpu().numpy()[0]
    assert abs(rec[42] - want) <= 1e-5 * max(abs(want), 1e-30)
    assert np.array_equal(rec[:42], A0), "mode 2 must leave A and b alone"


def test_batch_equals_single_and_is_deterministic(scn, oracle, cuda_dev):
    pts = cu(oracle.compute_points(scn.depths[2], scn.K))
    vols = scn.vols()
    dev = [(cu(v.tsdf), cu(v.weights)) for v in vols]
    cvs = [ops.volume(t, w, v.res, v.voxel, v.trunc, vid=v.vid) for v, (t, w) in zip(vols, dev)]
    Ts = [_perturbed(rel_pose_CO(scn.cam(2), v.pose), k) for k, v in enumerate(vols)]
    assoc = [torch.rand((scn.h, scn.w), device=DEV) for _ in vols]
    iw = [torch.zeros((scn.h, scn.w), device=DEV) for _ in vols]
    modes = [1] * len(vols)
    if len(vols) > 2:
        modes[1] = 0
    rec = torch.full((len(vols), 48), -3.0, device=DEV)
    ops.trackLinearise(cvs, Ts, modes, pts, assoc, 0.2, 64.0, iw, rec)
    rec2 = torch.full((len(vols), 48), -3.0, device=DEV)
    ops.trackLinearise(cvs, Ts, modes, pts, assoc, 0.2, 64.0, iw, rec2)
    torch.cuda.synchronize()
    assert torch.equal(rec[:, :45], rec2[:, :45]), "same inputs must give the same bits"
    for i, v in enumerate(vols):
        if modes[i] == 0:
            assert float(rec[i, 0]) == -3.0, "a skipped volume's record is untouched"
            continue
        g = _run_ours(v, Ts[i], pts, assoc[i])
        assert torch.equal(g["rec"][0, :45], rec[i, :45]), f"volume {i}: batched != single"


def test_tile_culling_changes_nothing(scn, oracle, cuda_dev):
    """with the camera matrix the kernel skips the image tiles that cannot see the volume: same bits everywhere"""
    pts = cu(oracle.compute_points(scn.depths[2], scn.K))
    for k, v in enumerate(scn.vols()):
        T = _perturbed(rel_pose_CO(scn.cam(2), v.pose), k)
        assoc = torch.rand((scn.h, scn.w), device=DEV)
        a = _run_ours(v, T, pts, assoc)
        b = _run_ours(v, T, pts, assoc, intr=scn.K)
        for key in ("grads", "tsdfVals", "trackWeights", "intWeights"):
            assert torch.equal(a[key], b[key]), f"vol {v.vid} {key}"
        # (the number of CTAs, hence the order of the sums, follows the rectangle: the sums agree to rounding)
        ra, rb = a["rec"][0, :45].double().cpu().numpy(), b["rec"][0, :45].double().cpu().numpy()
        assert np.abs(ra[:36] - rb[:36]).max() <= 1e-6 * max(np.abs(ra[:36]).max(), 1e-30)
        assert np.abs(ra[36:42] - rb[36:42]).max() <= 1e-6 * max(np.abs(ra[36:42]).max(), 1e-30)
        assert np.allclose(ra[42:45], rb[42:45], rtol=1e-6, atol=0)
        T2 = _perturbed(T, 3)
        a2 = _run_ours(v, T2, pts, assoc, mode=2, rec=a["rec"], iw=a["intWeights"])
        b2 = _run_ours(v, T2, pts, assoc, mode=2, rec=b["rec"], iw=b["intWeights"], intr=scn.K)
        assert torch.equal(a2["tsdfVals"], b2["tsdfVals"])
        assert abs(float(a2["rec"][0, 42]) - float(b2["rec"][0, 42])) <= 1e-6 * abs(float(a2["rec"][0, 42]))


@pytest.mark.skipif(not ref_gpu.available(), reason="oracle/_ref not built")
def test_linearise_vs_reference_chain(scn, oracle, cuda_dev):
    rng = np.random.default_rng(13)
    pts_h = oracle.compute_points(scn.depths[2], scn.K)
    pts = cu(pts_h)
    n = scn.w * scn.h
    rt = ref_gpu.RefTracker(scn.w, scn.h)
    try:
        for k, v in enumerate(scn.vols()):
            T = _perturbed(rel_pose_CO(scn.cam(2), v.pose), k)
            assoc = cu(rng.random((scn.h, scn.w), dtype=np.float32))
            gvol = torch.zeros((v.n, 3), device=DEV)
            tsdf, wts = cu(v.tsdf), cu(v.weights)
            ref_gpu.update_gradients(tsdf, gvol, v.res)
            A_r, b_r, e_r = rt.linearise(tsdf, gvol, wts, pts, assoc, S.R9(T), S.T3(T), v.res, v.voxel)
            g = _run_ours(v, T, pts, assoc)
            assert_bits(g["grads"], rt.image(rt.GRADS, (n, 6)).cpu().numpy(), f"vol {v.vid} grads vs reference kernel")
            assert_bits(g["tsdfVals"], rt.image(rt.TSDF_VALS, (scn.h, scn.w)).cpu().numpy(), "tsdfVals vs reference kernel")
            assert_bits(g["trackWeights"], rt.image(rt.TRACK_WEIGHTS, (scn.h, scn.w)).cpu().numpy(), "trackWeights")
            np.testing.assert_allclose(ops_normalised(g), rt.image(rt.INT_WEIGHTS, (scn.h, scn.w)).cpu().numpy(), rtol=6e-7,
                                       atol=1e-12)
            rec = g["rec"].cpu().numpy()[0]
            # float64 column sums of the reference's own per-pixel products = what its reduce approximates
            As64 = rt.image(rt.AS, (n, 36)).double().sum(0).cpu().numpy()
            bs64 = rt.image(rt.BS, (n, 6)).double().sum(0).cpu().numpy()
            sA, sb = max(np.abs(As64).max(), 1e-30), max(np.abs(bs64).max(), 1e-30)
            assert np.abs(rec[:36] - As64).max() <= 1e-5 * sA
            assert np.abs(rec[36:42] - bs64).max() <= 1e-5 * sb
            assert np.abs(rec[:36].reshape(6, 6) - A_r).max() <= 2e-4 * sA
            assert np.abs(rec[36:42] - b_r).max() <= 2e-4 * sb
            assert abs(rec[42] - e_r) <= 1e-4 * max(abs(e_r), 1e-30)
            T2 = _perturbed(T, 2)
            e2_r = rt.error(tsdf, pts, S.R9(T2), S.T3(T2), v.res, v.voxel)
            g2 = _run_ours(v, T2, pts, assoc, mode=2, rec=g["rec"], iw=g["intWeights"])
            assert abs(float(g2["rec"][0, 42]) - e2_r) <= 1e-4 * max(abs(e2_r), 1e-30)
    finally:
        rt.close()


def test_se3_exp_log_roundtrip():
    rng = np.random.default_rng(3)
    for _ in range(50):
        x = rng.normal(size=6) * np.array([0.3, 0.3, 0.3, 0.8, 0.8, 0.8])
        T = se3_exp(x)
        assert np.allclose(T.R @ T.R.T, np.eye(3), atol=1e-12)
        assert np.allclose(se3_log(T), x, atol=1e-9)
    assert np.allclose(se3_log(se3_exp(np.zeros(6))), 0)


def test_tracker_recovers_camera_pose(cuda_dev, oracle):
    """performTracking on the background: frames rendered from the true pose, tracker started from a perturbed one"""
    sc = S.make("track", oracle, 320, 240, (96, 96, 96), 0, (32, 32, 32), n_frames=3, integrate_frames=2)
    prm = TSDFParams()
    vol = TSDF(sc.bg.res, sc.bg.voxel, sc.bg.trunc, sc.bg.pose, prm, (sc.w, sc.h), device=DEV)
    vol.tsdfVol.copy_(cu(sc.bg.tsdf).view_as(vol.tsdfVol))
    vol.tsdfWeights.copy_(cu(sc.bg.weights).view_as(vol.tsdfWeights))
    pts = cu(oracle.compute_points(sc.depths[2], sc.K))
    assoc = torch.ones((sc.h, sc.w), device=DEV)
    true_cam = sc.cam(2)
    start = true_cam * se3_exp(np.array([0.01, -0.008, 0.012, 0.006, -0.005, 0.004]))
    tr = Tracker([vol], (sc.w, sc.h), DEV, intr=sc.K)
    st = tr.track(pts, [assoc], start, maxTrackingIter=100)[0]
    cam = tr.syncTrackCamera(0)
    d0 = np.linalg.norm(se3_log(true_cam.inv() * start))
    d1 = np.linalg.norm(se3_log(true_cam.inv() * cam))
    assert st.iterations >= 2 and st.linearisations >= 1
    assert d1 < 0.25 * d0, f"pose error {d0:.4f} -> {d1:.4f} after {st.iterations} iterations"
    # two device reads per iteration for the whole batch, at most
    assert tr.device_reads <= 2 * st.iterations + 1


def _bg_volume(sc):
    prm = TSDFParams()
    vol = TSDF(sc.bg.res, sc.bg.voxel, sc.bg.trunc, sc.bg.pose, prm, (sc.w, sc.h), device=DEV)
    vol.tsdfVol.copy_(cu(sc.bg.tsdf).view_as(vol.tsdfVol))
    vol.tsdfWeights.copy_(cu(sc.bg.weights).view_as(vol.tsdfWeights))
    return vol


def test_device_loop_equals_host_loop(cuda_dev, oracle):
    """emf_track_iterate (the Levenberg-Marquardt loop with its host part on the device, nothing read back in between)
    against Tracker.track (the same control flow in Python, two reads per iteration): same number of iterations and
    linearisations, same decisions, the same pose"""
    sc = S.make("track", oracle, 320, 240, (96, 96, 96), 2, (32, 32, 32), n_frames=3, integrate_frames=2)
    pts = cu(oracle.compute_points(sc.depths[2], sc.K))
    vols = [_bg_volume(sc)]
    ObjTSDF.nextID = 0
    for o in sc.objs:
        v = ObjTSDF(o.res, o.voxel, o.trunc, sc.scene.object_pose(o.vid - 1, 2), TSDFParams(), (sc.w, sc.h), device=DEV)
        v.tsdfVol.copy_(cu(o.tsdf).view_as(v.tsdfVol)); v.tsdfWeights.copy_(cu(o.weights).view_as(v.tsdfWeights))
        vols.append(v)
    assoc = [torch.ones((sc.h, sc.w), device=DEV) for _ in vols]
    start = sc.cam(2) * se3_exp(np.array([0.008, -0.006, 0.01, 0.005, -0.004, 0.003]))
    host = Tracker(vols, (sc.w, sc.h), DEV, intr=sc.K)
    hs = [(s.rel_pose_CO, s.iterations, s.linearisations, s.trackingConverged) for s in host.track(pts, assoc, start, 40)]
    devt = Tracker(vols, (sc.w, sc.h), DEV, intr=sc.K)
    ds = devt.track_device(pts, assoc, start, 40)
    for i, (s, (pose_h, it_h, lin_h, conv_h)) in enumerate(zip(ds, hs)):
        d = np.linalg.norm(se3_log(pose_h.inv() * s.rel_pose_CO))
        # (the background: hundreds of thousands of pixels, a well conditioned problem.  A 32^3 object seen in a few hundred
        #  pixels amplifies the last-bit differences between LAPACK's and the device's 6 x 6 float solve over the iterations)
        tol, dit = (2e-5, 2) if i == 0 else (1e-3, 10)
        assert d < tol, f"volume {i}: device and host loops end {d:.2e} apart"
        assert abs(s.iterations - it_h) <= dit and abs(s.linearisations - lin_h) <= dit, (i, s.iterations, it_h, s.linearisations, lin_h)
        assert np.allclose(s.rel_pose_CO.R @ s.rel_pose_CO.R.T, np.eye(3), atol=1e-9)
    # the background's pose moved towards the truth
    true_rel = vols[0].pose.inv() * sc.cam(2)
    d0 = np.linalg.norm(se3_log(true_rel.inv() * (vols[0].pose.inv() * start)))
    d1 = np.linalg.norm(se3_log(true_rel.inv() * ds[0].rel_pose_CO))
    assert d1 < d0
    # host reads: the device loop looks at snapshots only (at most one per chunk + the final one)
    assert devt.device_reads <= 40 // 4 + 1 < host.device_reads


def test_device_loop_stops_on_convergence(cuda_dev, oracle):
    """started at the optimum of a previous run the loop converges at once: converged volumes make the remaining launches
    no-ops and the host stops enqueuing when the snapshot says so"""
    sc = S.make("track", oracle, 160, 120, (64, 64, 64), 0, (32, 32, 32), n_frames=3, integrate_frames=2)
    pts = cu(oracle.compute_points(sc.depths[2], sc.K))
    vol = _bg_volume(sc)
    assoc = [torch.ones((sc.h, sc.w), device=DEV)]
    tr = Tracker([vol], (sc.w, sc.h), DEV, intr=sc.K)
    s1 = tr.track_device(pts, assoc, sc.cam(2), 100)[0]
    cam1 = tr.syncTrackCamera(0)
    it1 = s1.iterations
    s2 = tr.track_device(pts, assoc, cam1, 100)[0]
    assert s2.trackingConverged or s2.iterations <= it1
    assert tr.iterations_enqueued <= 100
    assert np.linalg.norm(se3_log(cam1.inv() * tr.syncTrackCamera(0))) < 1e-3
