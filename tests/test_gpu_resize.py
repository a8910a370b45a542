"""-m gpu: emf_copy_values / emf_resize_volume (csrc/resize.cu) and the host mirror of emf::ObjTSDF::resize against
 (a) the reference's own copyValues kernel (oracle/_ref, compiled unchanged), zero-filled targets as ObjTSDF::resize prepares
     them (src/core/ObjTSDF.cpp:116-147) -- bit-exact, every channel count, offsets that clip on every side;
 (b) a numpy restatement of the index arithmetic;
 (c) end to end: an object volume is resized and keeps raycasting / integrating to the same surface."""
import numpy as np
import pytest
import torch

from emfusion_b200 import ops
from emfusion_b200.poses import Affine
from emfusion_b200.volume import ObjTSDF, TSDFParams
from tests import ref_gpu
from tests.test_gpu_parity import DEV, assert_bits, cu

pytestmark = pytest.mark.gpu

CASES = [
    # src_res, dst_res, offset
    ((32, 32, 32), (40, 40, 40), (-4, -4, -4)),        # centred growth
    ((32, 32, 32), (48, 48, 48), (3, -11, -20)),       # shifted: clips on some sides, leaves gaps on others
    ((24, 20, 28), (16, 16, 16), (5, 2, 9)),           # shrinking
    ((22, 26, 30), (34, 34, 34), (-7, 1, -2)),         # ragged resolutions (scalar path of the fused kernel)
    ((16, 16, 16), (16, 16, 16), (40, 0, 0)),          # no overlap at all
]


def numpy_copy(src, offset, src_res, dst_res, ch):
    sx, sy, sz = src_res
    dx, dy, dz = dst_res
    s = src.reshape(sz, sy, sx, ch)
    d = np.zeros((dz, dy, dx, ch), np.float32)
    for z in range(sz):
        zn = z - offset[2]
        if not 0 <= zn < dz:
            continue
        for y in range(sy):
            yn = y - offset[1]
            if not 0 <= yn < dy:
                continue
            x0, x1 = max(0, offset[0]), min(sx, dx + offset[0])
            if x1 > x0:
                d[zn, yn, x0 - offset[0]:x1 - offset[0]] = s[z, y, x0:x1]
    return d.reshape(-1)


@pytest.mark.parametrize("src_res,dst_res,offset", CASES)
@pytest.mark.parametrize("ch", [1, 2, 3])
def test_copy_values(cuda_dev, src_res, dst_res, offset, ch):
    rng = np.random.default_rng(1)
    n_s, n_d = int(np.prod(src_res)), int(np.prod(dst_res))
    src = rng.standard_normal(n_s * ch).astype(np.float32)
    shape = (dst_res[1] * dst_res[2], dst_res[0]) if ch == 1 else (dst_res[1] * dst_res[2], dst_res[0], ch)
    sshape = (src_res[1] * src_res[2], src_res[0]) if ch == 1 else (src_res[1] * src_res[2], src_res[0], ch)
    s_d = cu(src).view(sshape)
    ours = torch.full(shape, 9.0, device=DEV)
    ops.copyValues(s_d, ours, offset, src_res, dst_res)
    want = numpy_copy(src, offset, src_res, dst_res, ch)
    touched = numpy_copy(np.ones_like(src), offset, src_res, dst_res, ch) > 0
    got = ours.cpu().numpy().reshape(-1)
    assert np.array_equal(got[touched], want[touched])
    assert np.all(got[~touched] == 9.0), "copyValues must not touch voxels without a source"
    if ref_gpu.available():
        ref = torch.full(shape, 9.0, device=DEV)
        ref_gpu.copy_values(s_d, ref, ch, offset, src_res, dst_res)
        assert_bits(ours, ref.cpu().numpy(), "copyValues vs the reference kernel")


@pytest.mark.parametrize("src_res,dst_res,offset", CASES)
def test_resize_volume_fused(cuda_dev, src_res, dst_res, offset):
    rng = np.random.default_rng(2)
    n_s, n_d = int(np.prod(src_res)), int(np.prod(dst_res))
    t, w, f = (rng.standard_normal(n_s).astype(np.float32), rng.random(n_s).astype(np.float32),
               rng.random(2 * n_s).astype(np.float32))
    dt, dw, df = (torch.full((n_d,), 7.0, device=DEV), torch.full((n_d,), 7.0, device=DEV), torch.full((2 * n_d,), 7.0, device=DEV))
    ops.resizeVolume(cu(t), cu(w), cu(f), src_res, dt, dw, df, dst_res, offset)
    assert np.array_equal(dt.cpu().numpy(), numpy_copy(t, offset, src_res, dst_res, 1))
    assert np.array_equal(dw.cpu().numpy(), numpy_copy(w, offset, src_res, dst_res, 1))
    assert np.array_equal(df.cpu().numpy(), numpy_copy(f, offset, src_res, dst_res, 2))
    if ref_gpu.available():   # the reference's sequence: setTo(0), then copyValues
        rt, rf = torch.zeros((n_d,), device=DEV), torch.zeros((2 * n_d,), device=DEV)
        ref_gpu.copy_values(cu(t), rt, 1, offset, src_res, dst_res)
        ref_gpu.copy_values(cu(f), rf, 2, offset, src_res, dst_res)
        assert_bits(dt, rt.cpu().numpy(), "resize tsdf vs reference")
        assert_bits(df, rf.cpu().numpy(), "resize fg/bg vs reference")
    # without fg/bg counts (a plain TSDF)
    dt2, dw2 = torch.full((n_d,), 7.0, device=DEV), torch.full((n_d,), 7.0, device=DEV)
    ops.resizeVolume(cu(t), cu(w), None, src_res, dt2, dw2, None, dst_res, offset)
    assert torch.equal(dt2, dt) and torch.equal(dw2, dw)


def test_objtsdf_resize_host_logic(cuda_dev):
    """ObjTSDF::resize: contained boxes change nothing; otherwise the grid is recentred on a voxel multiple, gets the
    reference's even resolution, and every old voxel that still fits keeps its value at its new index."""
    ObjTSDF.nextID = 0
    prm = TSDFParams()
    vs = 0.01
    o = ObjTSDF((32, 32, 32), vs, 10 * vs, Affine.translation([0.1, 0.2, 1.5]), prm, (160, 120), device=DEV)
    rng = np.random.default_rng(4)
    o.tsdfVol.copy_(cu(rng.standard_normal(32 ** 3).astype(np.float32)).view_as(o.tsdfVol))
    o.tsdfWeights.copy_(cu(rng.random(32 ** 3).astype(np.float32)).view_as(o.tsdfWeights))
    o.fgBgProbs.copy_(cu(rng.integers(0, 5, 2 * 32 ** 3).astype(np.float32)).view_as(o.fgBgProbs))
    old = o.tsdfVol.cpu().numpy().reshape(32, 32, 32).copy()
    old_fg = o.fgBgProbs.cpu().numpy().reshape(32, 32, 32, 2).copy()
    pose0 = o.pose.copy()
    assert np.all(o.resize([-0.1, -0.1, -0.1], [0.1, 0.1, 0.1], 2.0) == 0) and o.volumeRes == (32, 32, 32)
    c = o.resize([-0.05, -0.02, 0.03], [0.21, 0.12, 0.20], 2.0)
    # newCenter = round((p10 + p90) / 2 / voxel) * voxel ; resolution = next even >= ceil(volPad * max extent / voxel)
    assert np.allclose(c, np.rint(np.array([0.08, 0.05, 0.115], np.float32) / np.float32(vs)) * np.float32(vs), atol=1e-6)
    n = (int(np.ceil(np.float32(2.0) * np.float32(0.26) / np.float32(vs))) + 1) // 2 * 2
    assert o.volumeRes == (n, n, n) and o.tsdfVol.shape == (n * n, n) and o.fgProbs.shape == (n * n, n)
    assert np.allclose(o.pose.t, pose0.t + pose0.R @ c.astype(np.float64))
    off = np.rint(c / np.float32(vs)).astype(int) - (n - 32) // 2
    new = o.tsdfVol.cpu().numpy().reshape(n, n, n)
    new_fg = o.fgBgProbs.cpu().numpy().reshape(n, n, n, 2)
    checked = 0
    for z in range(0, 32, 3):
        for y in range(0, 32, 5):
            for x in range(0, 32, 7):
                xn, yn, zn = x - off[0], y - off[1], z - off[2]
                if 0 <= xn < n and 0 <= yn < n and 0 <= zn < n:
                    assert new[zn, yn, xn] == old[z, y, x] and np.array_equal(new_fg[zn, yn, xn], old_fg[z, y, x])
                    checked += 1
    assert checked > 50
    # the world position of a kept voxel is unchanged: old centre + (x - 15.5) vs = new centre + (xn - (n-1)/2) vs
    x, xn = 10, 10 - off[0]
    w_old = pose0.t[0] + (x - 15.5) * vs
    w_new = o.pose.t[0] + (xn - (n - 1) / 2) * vs
    assert abs(w_old - w_new) < 1e-6
    # fg probabilities were recomputed for the new grid
    s = new_fg[..., 0] + new_fg[..., 1]
    want = np.where(s != 0, new_fg[..., 0] / np.where(s != 0, s, 1), 0).astype(np.float32)
    assert np.allclose(o.fgProbs.cpu().numpy().reshape(n, n, n), want)
