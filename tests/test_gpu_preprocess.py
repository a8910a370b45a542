"""-m gpu: emf_preprocess_depth (bilateral pre-filter + NaN / zero patch + optional un-projection, one launch) against the
C oracle's restatement of OpenCV-CUDA's published bilateral kernel.  PARITY UNPINNED with respect to OpenCV itself (the
dependency is absent from this image); tolerance 2e-6 relative: expf differs by <= 2 ulp between libm and CUDA."""
import numpy as np
import pytest
import torch

from emfusion_b200 import ops
from emfusion_b200.synth import Scene
from tests.test_gpu_parity import DEV, assert_bits, cu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("w,h,ksize", [(640, 480, 7), (161, 119, 7), (64, 48, 5), (40, 30, 0)])
def test_preprocess_depth_vs_oracle(cuda_dev, oracle, w, h, ksize):
    scene = Scene(n_objects=3, width=w, height=h, seed=2, noise_sigma=0.003, dropout=0.03)
    raw, _ = scene.render(1)
    raw[5:9, 3:20] = np.nan if ksize == 5 else raw[5:9, 3:20]     # NaNs in the input must come out as 0 around them
    want = oracle.preprocess_depth(raw, ksize, 0.04, 4.5)
    out = torch.full((h, w), 7.0, device=DEV)
    pts = torch.full((h, w, 3), 7.0, device=DEV)
    ops.preprocessDepth(cu(raw), out, pts, scene.K, ksize, 0.04, 4.5)
    got = out.cpu().numpy()
    assert not np.isnan(got).any()
    assert np.array_equal(got == 0, want == 0)
    assert (got[raw == 0] == 0).all()
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-7)
    # the fused un-projection is computePoints of the filtered depth, bit for bit
    pts2 = torch.zeros((h, w, 3), device=DEV)
    ops.computePoints(out, pts2, scene.K)
    assert_bits(pts, pts2.cpu().numpy(), "fused points")
    # smoothing: noise goes down where there is no depth edge
    if ksize == 7 and w == 640:
        clean, _ = Scene(n_objects=3, width=w, height=h, seed=2).render(1)
        flat = (np.abs(clean - np.roll(clean, 1, 1)) < 0.01) & (raw > 0) & (want > 0)
        flat[:, :8] = False; flat[:, -8:] = False; flat[:8] = False; flat[-8:] = False
        assert np.abs(got - clean)[flat].std() < 0.7 * np.abs(raw - clean)[flat].std()
    # without points
    out2 = torch.zeros((h, w), device=DEV)
    ops.preprocessDepth(cu(raw), out2, None, None, ksize, 0.04, 4.5)
    assert torch.equal(out2, out)
