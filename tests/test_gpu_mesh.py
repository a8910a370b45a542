"""-m gpu: marching cubes (csrc/mcubes.cu, emf_mesh_count / emf_mesh_extract = emf::TSDF::getMesh) against the reference's own
emf::cuda::TSDF::marchingCubes (oracle/_ref) on GPU-integrated volumes: the same number of vertices and triangles, vertex
positions and (un-normalised) normals bit for bit in the reference's order, every triangle inside its cube's vertex range, and
the same oriented patch boundaries -- the interior diagonals of patches with more than three vertices are this repo's own
(scripts/gen_mc_tables.py)."""
import numpy as np
import pytest
import torch

from emfusion_b200 import ops
from emfusion_b200.poses import Affine, rel_pose_OC
from emfusion_b200.synth import Scene
from emfusion_b200.volume import ObjTSDF, Params, TSDF
from tests import ref_gpu
from tests.test_gpu_parity import DEV, assert_bits, cu

pytestmark = pytest.mark.gpu


def boundary(tris):
    """directed edges (a, b) of a triangle list whose reverse (b, a) does not occur: the patch boundaries"""
    t = tris[:, 1:].astype(np.int64)
    e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]])
    key = e[:, 0] * (1 << 32) + e[:, 1]
    rev = e[:, 1] * (1 << 32) + e[:, 0]
    return np.sort(key[~np.isin(key, rev)])


@pytest.mark.skipif(not ref_gpu.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("res,obj", [((96, 96, 96), False), ((80, 64, 72), False), ((64, 64, 64), True)], ids=["bg96", "ragged", "object"])
def test_mesh_equals_reference(cuda_dev, res, obj):
    w, h = 320, 240
    scene = Scene(n_objects=2, width=w, height=h, seed=4)
    prm = Params(frameSize=(w, h), intr=scene.K)
    ObjTSDF.nextID = 0
    if obj:
        voxel = scene.object_voxel_size(0, res[0])
        vol = ObjTSDF(res, voxel, float(np.float32(10) * np.float32(voxel)), scene.object_pose(0, 0), prm.tsdfParams, (w, h), DEV)
    else:
        voxel = float(np.float32(5.12 / max(res)))
        vol = TSDF(res, voxel, float(np.float32(10) * np.float32(voxel)), Affine.translation([0, 0, 2.56]), prm.tsdfParams, (w, h), DEV)
    ones = torch.ones((h, w), device=DEV)
    zeros = torch.zeros((h, w), dtype=torch.uint8, device=DEV)
    for f in range(3):
        depth, inst = scene.render(f)
        vol.integrate(cu(depth), ones, scene.cam_pose(f), scene.K)
        if obj:
            vol.integrateMask(cu((inst == 1).astype(np.uint8)), zeros, scene.cam_pose(f), scene.K)
    mesh = vol.getMesh()
    # the reference: gradient volume, mask volume (weights > 0 [& fgProb > 0.5]), its launcher
    n = int(np.prod(res))
    grads = torch.zeros(3 * n, device=DEV)
    ops.computeTSDFGrads(vol.tsdfVol, grads, res)
    mask = (vol.tsdfWeights.reshape(-1) > 0)
    if obj:
        mask = mask & (vol.fgProbs.reshape(-1) > 0.5)
    mask = mask.to(torch.uint8) * 255
    rv, rn, rt = ref_gpu.marching_cubes(vol.tsdfVol.reshape(-1), grads, mask, res, voxel)
    assert mesh["cloud"].shape[0] == rv.shape[0] > 1000 and mesh["polygons"].shape[0] == rt.shape[0] > 500
    assert_bits(mesh["cloud"], rv.cpu().numpy(), "vertices")
    assert_bits(mesh["normals"], rn.cpu().numpy(), "normals")
    a, b = mesh["polygons"].cpu().numpy(), rt.cpu().numpy()
    assert (a[:, 0] == 3).all() and (b[:, 0] == 3).all()
    # triangle k of both lists belongs to the same cube: its indices lie in the same range
    assert (a[:, 1:].min(1) // 1 >= 0).all() and np.abs(a[:, 1:].min(1) - b[:, 1:].min(1)).max() <= 11
    assert np.array_equal(boundary(a), boundary(b)), "patch boundaries differ"
    # (about a third of the triangles are identical; the others differ by a patch's interior diagonal only -- the boundary check)
    # every vertex is used by both
    assert np.array_equal(np.unique(a[:, 1:]), np.unique(b[:, 1:]))


def test_mesh_of_an_empty_volume(cuda_dev):
    prm = Params(frameSize=(64, 48))
    vol = TSDF((32, 32, 32), 0.05, 0.5, Affine.identity(), prm.tsdfParams, (64, 48), DEV)
    m = vol.getMesh()
    assert m["cloud"].shape == (0, 3) and m["polygons"].shape == (0, 4)
