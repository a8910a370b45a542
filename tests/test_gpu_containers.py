"""-m gpu: BASELINE.json's configs 2 and 3 through the reference's own stream containers (emfusion_b200/io.py) into the engine.
 * config 2 -- 512^3 background, no objects, TUM RGB-D container (associations.txt, 16-bit PNG depth x 5000,
   reference src/utils/TUMRGBDReader.cpp:38-104): the frames the reader hands out drive the engine to exactly the state the
   same (quantised) frames give in memory, and the raycast sees the scene;
 * config 3 -- 512^3 background + 8 objects @64^3, Co-Fusion container (colour/Color%04d.png, depth/Depth%04d.exr float
   metres, src/utils/ImageReader.cpp:41-116): the EXR round trip is lossless, so the engine ends bit-identical to the run
   on the rendered frames."""
import numpy as np
import pytest
import torch

from emfusion_b200 import io as emfio
from emfusion_b200.native import NativeEngine
from emfusion_b200.synth import Scene
from emfusion_b200.volume import ObjTSDF, Params
from tests.test_gpu_parity import DEV, assert_bits, cu

pytestmark = pytest.mark.gpu


def engine(scene, w, h, bg, n_obj, obj):
    prm = Params(frameSize=(w, h), intr=scene.K, globalVolumeDims=(bg,) * 3, globalVoxelSize=5.12 / bg, objVolumeDims=(obj,) * 3)
    ObjTSDF.nextID = 0
    eng = NativeEngine(prm, DEV)
    for k in range(n_obj):
        eng.add_object(scene.object_pose(k, 0), scene.object_voxel_size(k, obj))
    return eng


def drive(eng, scene, depths, insts):
    for f, d in enumerate(depths):
        eng.processFrame(cu(d), scene.cam_pose(f), {o.id: scene.object_pose(o.id - 1, f) for o in eng.objects})
        if f == 0:
            zeros = torch.zeros((eng.h, eng.w), dtype=torch.uint8, device=DEV)
            for o in eng.objects:
                o.integrateMask(cu((insts[0] == o.id).astype(np.uint8)), zeros, eng.pose, eng.params.intr)
    torch.cuda.synchronize()


def test_config2_through_tum_container(tmp_path, cuda_dev):
    w, h, n = 640, 480, 4
    scene = Scene(n_objects=0, width=w, height=h, seed=0)
    frames = [scene.render(f) for f in range(n)]
    seen = emfio.write_tum_stream(str(tmp_path), [d for d, _ in frames])
    rd = emfio.TUMRGBDReader(str(tmp_path))
    assert rd.numFrames() == n
    read = [rd.readFrame(f)[1] for f in range(n)]
    a, b = engine(scene, w, h, 512, 0, 64), engine(scene, w, h, 512, 0, 64)
    drive(a, scene, read, [i for _, i in frames])
    drive(b, scene, seen, [i for _, i in frames])
    assert_bits(a.background.tsdfVol, b.background.tsdfVol.cpu().numpy(), "background tsdf")
    assert_bits(a.raylengths, b.raylengths.cpu().numpy(), "ray lengths")
    assert int(a.bg_mask.sum()) > 0.8 * w * h                      # the room is seen
    assert float(np.abs(read[1] - frames[1][0]).max()) <= 0.5 / 5000 + 1e-6


def test_config3_through_cofusion_container(tmp_path, cuda_dev):
    w, h, n, k = 640, 480, 4, 8
    scene = Scene(n_objects=k, width=w, height=h, seed=0)
    frames = [scene.render(f) for f in range(n)]
    cp, dp = emfio.write_cofusion_stream(str(tmp_path), [d for d, _ in frames])
    rd = emfio.ImageReader(cp, dp)
    assert rd.numFrames() == n and rd.currFrame == 0
    read = [rd.readFrame(f)[1] for f in range(n)]
    for r, (d, _) in zip(read, frames):
        assert np.array_equal(r, d)
    a, b = engine(scene, w, h, 512, k, 64), engine(scene, w, h, 512, k, 64)
    drive(a, scene, read, [i for _, i in frames])
    drive(b, scene, [d for d, _ in frames], [i for _, i in frames])
    for va, vb in zip(a.local_volumes(), b.local_volumes()):
        assert_bits(va.tsdfVol, vb.tsdfVol.cpu().numpy(), f"tsdf of volume {va.id}")
    assert_bits(a.modelSegmentation, b.modelSegmentation.cpu().numpy(), "segmentation")
    assert_bits(a.raylengths, b.raylengths.cpu().numpy(), "ray lengths")
    assert len(a.vis_objs) >= 4 and int((a.modelSegmentation > 0).sum()) > 1000
