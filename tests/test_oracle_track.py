"""CPU: the oracle's restatement of one tracker iteration (oracle/emf_oracle.c emfo_track_linearise) is self-consistent:
b is half the derivative of the weighted error along each twist direction (weights held fixed), A is symmetric PSD, and
the SE(3) helpers of the host mirror invert each other."""
import numpy as np

from emfusion_b200.poses import rel_pose_CO
from emfusion_b200.tracking import orthonormalise, se3_exp, se3_log
from tests import scenario as S


def test_track_oracle_jacobian_matches_finite_differences(oracle):
    sc = S.make("trk", oracle, 160, 120, (64, 64, 64), 0, (32, 32, 32), n_frames=2, integrate_frames=2)
    v = sc.bg
    pts = oracle.compute_points(sc.depths[1], sc.K)
    T = se3_exp(np.array([0.004, -0.003, 0.005, 0.003, -0.002, 0.004])) * rel_pose_CO(sc.cam(1), v.pose)
    assoc = np.ones((sc.h, sc.w), np.float32)
    gvol = oracle.compute_grads(v.tsdf, v.res)
    o = oracle.track_linearise(v.tsdf, gvol, v.weights, pts, assoc, S.R9(T), S.T3(T), v.res, v.voxel)
    A, b = o["A"], o["b"]
    assert np.allclose(A, A.T, rtol=0, atol=1e-9 * np.abs(A).max())
    assert np.linalg.eigvalsh(A).min() > -1e-6 * np.abs(A).max()
    w = o["intWeights"].astype(np.float64)

    def err(Tk):
        f, _ = oracle.get_volume_vals(v.tsdf, pts, S.R9(Tk), S.T3(Tk), v.res, v.voxel)
        return float((w * f.astype(np.float64) ** 2).sum())

    assert abs(err(T) - o["err"]) <= 1e-6 * o["err"]
    num = np.zeros(6)
    for k in range(6):
        d = np.zeros(6); d[k] = 2e-3 if k < 3 else 1e-3
        num[k] = (err(se3_exp(d) * T) - err(se3_exp(-d) * T)) / (4 * d[k])      # dE/dxi = 2 b
    # the tracker's gradient is the interpolated FORWARD difference of the TSDF, not the derivative of the trilinear
    # interpolant, so the two agree in direction and roughly in size only
    cos = float(num @ b / (np.linalg.norm(num) * np.linalg.norm(b)))
    assert cos > 0.97, (cos, num, b)
    assert 0.6 < np.linalg.norm(num) / np.linalg.norm(b) < 1.6


def test_orthonormalise_and_se3():
    rng = np.random.default_rng(0)
    T = se3_exp(rng.normal(size=6) * 0.5)
    T.R = T.R + 1e-4 * rng.normal(size=(3, 3))
    Q = orthonormalise(T)
    assert np.allclose(Q.R @ Q.R.T, np.eye(3), atol=1e-12) and np.linalg.det(Q.R) > 0
    assert np.abs(Q.R - T.R).max() < 1e-3
    x = rng.normal(size=6) * 0.4
    assert np.allclose(se3_log(se3_exp(x)), x, atol=1e-10)
