"""CPU: pins the C oracle (oracle/emf_oracle.c) against golden vectors produced by the REFERENCE'S OWN CUDA
kernels on a B200 (tests/golden/ref_kernels_small.npz, written by tests/golden/generate_golden.py from
oracle/_ref/libemf_ref.so = /root/reference/src/core/cuda/{TSDF,ObjTSDF}.cu compiled unchanged).

The reference ships no tests or fixtures of its own; this fixture is what turns the
oracle from "parity unpinned" into "pinned on the reference kernels' outputs".  Bar: bit-exact on every
output (floats are compared by bit pattern, +0 == -0): the oracle places fmaf() exactly where the reference
build contracts, so nothing on these kernels needs a tolerance.
"""
import os

import numpy as np
import pytest

from tests import oracle_c

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_kernels_small.npz")


@pytest.fixture(scope="module")
def g():
    return np.load(GOLDEN)


def same_bits(a, b):
    a = np.ascontiguousarray(a).reshape(-1)
    b = np.ascontiguousarray(b).reshape(-1)
    assert a.shape == b.shape and a.dtype == b.dtype
    if a.dtype == np.float32:
        return (a.view(np.uint32) == b.view(np.uint32)) | ((a == 0) & (b == 0)) | (np.isnan(a) & np.isnan(b))
    return a == b


def assert_bits(a, b, what):
    ok = same_bits(a, b)
    assert ok.all(), f"{what}: {int((~ok).sum())} of {ok.size} differ, first at {np.flatnonzero(~ok)[:5]}"


def _res(g, i):
    return tuple(int(r) for r in (g["bg_res"] if i == 0 else g["obj_res"]))


def test_golden_is_nontrivial(g):
    assert int(g["mask_0"].sum()) > 1000, "background raycast of the fixture should hit"
    assert int(g["mask_1"].sum()) > 50, "object raycast of the fixture should hit"
    assert (g["fgbg2"] != 0).any()
    assert (g["depth0"] == 0).any(), "fixture has dropouts (invalid depth branch)"


def test_integrate_three_frames_matches_reference_kernels(g, oracle):
    """A.2 kernel_updateTSDF, three frames accumulated (frame 0 assoc == 1, then random association images)."""
    H, W = int(g["H"]), int(g["W"])
    for i in range(2):
        res = _res(g, i)
        n = int(np.prod(res))
        tsdf, wts = np.zeros(n, np.float32), np.zeros(n, np.float32)
        for f in range(3):
            oracle.update_tsdf(g[f"depth{f}"], g[f"assoc{f}"], tsdf, wts, g[f"R_oc{f}_{i}"], g[f"t_oc{f}_{i}"], g["K"],
                               res, float(g["voxel"][i]), float(g["trunc"][i]), 64.0)
            assert_bits(tsdf, g[f"tsdf{f}_{i}"], f"tsdf vol {i} frame {f}")
            assert_bits(wts, g[f"weights{f}_{i}"], f"weights vol {i} frame {f}")
        assert (wts > 0).sum() > n // 50


def test_fgbg_matches_reference_kernel(g, oracle):
    """kernel_updateFgBgProbs on the reference's volume state of each frame, with a random occlusion mask."""
    res = _res(g, 1)
    fgbg = np.zeros(2 * int(np.prod(res)), np.float32)
    for f in range(3):
        oracle.update_fgbg((g[f"inst{f}"] == 1).astype(np.uint8), g[f"occl{f}"], g[f"tsdf{f}_1"].copy(),
                           g[f"weights{f}_1"].copy(), fgbg, g[f"R_oc{f}_1"], g[f"t_oc{f}_1"], g["K"], res,
                           float(g["voxel"][1]))
        assert_bits(fgbg, g[f"fgbg{f}"], f"fgbg frame {f}")


def test_gradients_match_reference_kernel(g, oracle):
    for i in range(2):
        assert_bits(oracle.compute_grads(g[f"tsdf2_{i}"].copy(), _res(g, i)), g[f"grads_{i}"], f"grads vol {i}")


def test_raycast_matches_reference_kernel(g, oracle):
    """A.4 kernel_raycastTSDF: hit mask, raylength, vertex and normal bit-exact for every pixel."""
    H, W = int(g["H"]), int(g["W"])
    for i in range(2):
        r = oracle.raycast(g[f"tsdf2_{i}"].copy(), g[f"grads_{i}"].copy(), g[f"weights2_{i}"].copy(), g[f"R_co_{i}"],
                           g[f"t_co_{i}"], g["K"], _res(g, i), float(g["voxel"][i]), float(g["trunc"][i]), W, H)
        assert_bits(r["mask"], g[f"mask_{i}"], f"mask vol {i}")
        assert_bits(r["ray"], g[f"ray_{i}"], f"raylength vol {i}")
        assert_bits(r["vert"], g[f"vert_{i}"], f"vertex vol {i}")
        assert_bits(r["norm"], g[f"norm_{i}"], f"normal vol {i}")
        hit = r["hit"].reshape(-1, 3)
        m = g[f"mask_{i}"].reshape(-1).astype(bool)
        assert (hit[m] >= 0).all() and (hit[~m] == -1).all()


def test_gather_matches_reference_kernel(g, oracle):
    """kernel_getVolumeVals<float>"""
    for i in range(2):
        vals, n_in = oracle.get_volume_vals(g[f"tsdf2_{i}"].copy(), g["points2"], g[f"R_co_{i}"], g[f"t_co_{i}"],
                                            _res(g, i), float(g["voxel"][i]))
        assert_bits(vals, g[f"gather_{i}"], f"gather vol {i}")
        assert n_in > 0


def test_points_match_fixture(g, oracle):
    """A.1: the fixture's points were formed in numpy fp32 with the same mul-then-div order."""
    assert_bits(oracle.compute_points(g["depth2"], g["K"]), g["points2"], "points")


def test_nofma_build_bounds_contraction_sensitivity(g):
    """The unfused build differs from the reference kernels by rounding only: TSDF within 1e-5, and the raycast
    hit mask almost everywhere equal -- i.e. FMA placement matters for bit-exactness, not for the 1e-4 bar."""
    o = oracle_c.load(nofma=True)
    assert not o.uses_fma()
    res = _res(g, 0)
    n = int(np.prod(res))
    tsdf, wts = np.zeros(n, np.float32), np.zeros(n, np.float32)
    for f in range(3):
        o.update_tsdf(g[f"depth{f}"], g[f"assoc{f}"], tsdf, wts, g[f"R_oc{f}_0"], g[f"t_oc{f}_0"], g["K"], res,
                      float(g["voxel"][0]), float(g["trunc"][0]), 64.0)
    # a voxel whose projection lands exactly on a .5 pixel boundary may pick the neighbouring pixel
    d = np.abs(tsdf - g["tsdf2_0"])
    assert np.quantile(d, 0.999) <= 1e-5
