"""-m gpu: the C++ frame engine (emf_engine_*, csrc/engine.cu, through emfusion_b200.native.NativeEngine) against the
per-stage level-3 path (engine.EMFusionEngine) -- every image, every volume, bit for bit over a moving multi-object
stream, including objects added mid-stream, device-side visibility gating, and the raycast cull / brick-map options."""
import numpy as np
import pytest
import torch

from emfusion_b200.engine import EMFusionEngine
from emfusion_b200.native import NativeEngine
from emfusion_b200.synth import Scene
from emfusion_b200.volume import ObjTSDF, Params
from tests.test_gpu_parity import DEV, assert_bits, cu

pytestmark = pytest.mark.gpu


def make(cls, scene, w, h, bg, n_obj, obj, accelerate=False):
    prm = Params(frameSize=(w, h), intr=scene.K, globalVolumeDims=(bg,) * 3, globalVoxelSize=5.12 / bg,
                 objVolumeDims=(obj,) * 3, visibilityThresh=(40 * 40 * w * h) // (640 * 480), boundary=max(2, 20 * w // 640))
    ObjTSDF.nextID = 0
    eng = cls(prm, DEV, accelerate=accelerate)
    for k in range(n_obj):
        eng.add_object(scene.object_pose(k, 0), scene.object_voxel_size(k, obj))
    return eng


def compare(a, b, f):
    for name in ("raylengths", "vertices", "normals", "modelSegmentation", "bg_raylengths", "bg_vertices", "bg_normals",
                 "bg_mask", "bg_associationWeights", "points"):
        assert_bits(getattr(a, name), getattr(b, name).cpu().numpy(), f"frame {f} {name}")
    for o in a.objects:
        assert_bits(a.obj_raylengths[o.id], b.obj_raylengths[o.id].cpu().numpy(), f"frame {f} obj ray {o.id}")
        assert_bits(a.obj_modelSegmentation[o.id], b.obj_modelSegmentation[o.id].cpu().numpy(), f"frame {f} obj mask {o.id}")
        assert_bits(a.associationWeights[o.id], b.associationWeights[o.id].cpu().numpy(), f"frame {f} assoc {o.id}")
    assert a.vis_objs == b.vis_objs, (f, a.vis_objs, b.vis_objs)


@pytest.mark.parametrize("accelerate", [False, True], ids=["plain", "brickmaps"])
def test_native_engine_equals_staged_engine(cuda_dev, accelerate):
    w, h, bg, n_obj, obj = 320, 240, 96, 5, 32
    scene = Scene(n_objects=n_obj + 1, width=w, height=h, seed=3, dropout=0.01)
    nat = make(NativeEngine, scene, w, h, bg, n_obj, obj, accelerate)
    ref = make(EMFusionEngine, scene, w, h, bg, n_obj, obj, accelerate)
    for f in range(6):
        depth, inst = scene.render(f)
        d = cu(depth)
        if f == 3:   # an object appears mid-stream (created after the previous frame's raycast)
            for eng in (nat, ref):
                ObjTSDF.nextID = n_obj
                eng.add_object(scene.object_pose(n_obj, f), scene.object_voxel_size(n_obj, obj))
        for eng in (nat, ref):
            poses = {o.id: scene.object_pose(o.id - 1, f) for o in eng.objects}
            eng.processFrame(d, scene.cam_pose(f), poses)
            if f == 0:
                zeros = torch.zeros((h, w), dtype=torch.uint8, device=DEV)
                for o in eng.objects:
                    o.integrateMask(cu((inst == o.id).astype(np.uint8)), zeros, eng.pose, eng.params.intr)
        for va, vb in zip(nat.local_volumes(), ref.local_volumes()):
            assert_bits(va.tsdfVol, vb.tsdfVol.cpu().numpy(), f"frame {f} tsdf vol {va.id}")
            assert_bits(va.tsdfWeights, vb.tsdfWeights.cpu().numpy(), f"frame {f} weights vol {va.id}")
        if f > 0:
            compare(nat, ref, f)
        if f == 3:   # the new object's volume was integrated in its first frame, unseen by that frame's raycast (EMFusion.cpp:550,918)
            for eng in (nat, ref):
                new = [o for o in eng.objects if o.id == n_obj + 1][0]
                assert float(new.tsdfWeights.max()) > 0, f"{type(eng).__name__}: an object created mid-stream was never integrated"
    assert len(nat.vis_objs) > 0
    assert int(nat.bg_mask.sum()) > 0.5 * w * h


def test_pose_assigned_between_stage_calls(cuda_dev):
    """the reference's order inside a frame: association, tracking (assigns the poses), association again, raycast,
    integrate.  Poses assigned between stage calls must be the ones the following stages use (native == staged, and the
    second association differs from the first)"""
    w, h, bg, n_obj, obj = 160, 120, 64, 2, 32
    scene = Scene(n_objects=n_obj, width=w, height=h, seed=6)
    nat = make(NativeEngine, scene, w, h, bg, n_obj, obj)
    ref = make(EMFusionEngine, scene, w, h, bg, n_obj, obj)
    for eng in (nat, ref):
        depth, inst = scene.render(0)
        eng.processFrame(cu(depth), scene.cam_pose(0), {o.id: scene.object_pose(o.id - 1, 0) for o in eng.objects})
        zeros = torch.zeros((h, w), dtype=torch.uint8, device=DEV)
        for o in eng.objects:
            o.integrateMask(cu((inst == o.id).astype(np.uint8)), zeros, eng.pose, eng.params.intr)
    depth, _ = scene.render(5)
    first = {}
    for eng in (nat, ref):
        eng.set_depth(cu(depth))
        eng.computeAssociationWeights()                 # at the poses of frame 0
        first[type(eng).__name__] = eng.bg_associationWeights.clone()
        eng.pose = scene.cam_pose(5)                    # "tracking" moves everything
        for o in eng.objects:
            o.pose = scene.object_pose(o.id - 1, 5)
        eng.computeAssociationWeights()
        eng.raycast()
        eng.integrateDepth()
    compare(nat, ref, 5)
    assert not torch.equal(first["NativeEngine"], nat.bg_associationWeights), "the second association used the stale poses"
    for va, vb in zip(nat.local_volumes(), ref.local_volumes()):
        assert_bits(va.tsdfVol, vb.tsdfVol.cpu().numpy(), f"tsdf vol {va.id}")


def test_native_engine_stage_times(cuda_dev):
    w, h = 160, 120
    scene = Scene(n_objects=2, width=w, height=h, seed=4)
    nat = make(NativeEngine, scene, w, h, 64, 2, 32)
    for f in range(3):
        depth, _ = scene.render(f)
        nat.processFrame(cu(depth), scene.cam_pose(f), {o.id: scene.object_pose(o.id - 1, f) for o in nat.objects}, timed=True)
    ms = nat.stage_ms()
    assert len(ms) == 3 and all(m >= 0 for m in ms) and sum(ms) > 0


def test_host_frame_pipeline_equals_direct_frames(cuda_dev):
    """HostFramePipeline (pinned depth in, pinned composite out, copies overlapped with the kernels) returns, frame by
    frame, exactly what processFrame on device-resident depth produces"""
    from emfusion_b200.pipeline import HostFramePipeline
    w, h, n_obj = 160, 120, 3
    scene = Scene(n_objects=n_obj, width=w, height=h, seed=5)
    a = make(NativeEngine, scene, w, h, 64, n_obj, 32)
    b = make(NativeEngine, scene, w, h, 64, n_obj, 32)
    pipe = HostFramePipeline(b)
    frames = [scene.render(f) for f in range(7)]
    pinned = [torch.from_numpy(d).pin_memory() for d, _ in frames]
    want, tickets = [], []
    for f in range(7):
        poses = {o.id: scene.object_pose(o.id - 1, f) for o in a.objects}
        a.processFrame(cu(frames[f][0]), scene.cam_pose(f), poses)
        tickets.append(pipe.submit(pinned[f], scene.cam_pose(f), poses))
        if f == 0:
            zeros = torch.zeros((h, w), dtype=torch.uint8, device=DEV)
            for eng in (a, b):
                for o in eng.objects:
                    o.integrateMask(cu((frames[0][1] == o.id).astype(np.uint8)), zeros, eng.pose, eng.params.intr)
        want.append((a.modelSegmentation.cpu().numpy().copy(), a.raylengths.cpu().numpy().copy()))
        if f >= 1:   # consume one frame behind, as a streaming host would
            seg, ray = pipe.result(tickets[f - 1])
            assert np.array_equal(seg.numpy(), want[f - 1][0]), f"frame {f - 1} segmentation"
            assert_bits(torch.from_numpy(ray.numpy().copy()), want[f - 1][1], f"frame {f - 1} ray lengths")
    seg, ray = pipe.result(tickets[-1])
    assert np.array_equal(seg.numpy(), want[-1][0])
    assert int((want[-1][0] > 0).sum()) > 0
    with pytest.raises(ValueError):
        pipe.result(tickets[0])
