"""Worker of tests/test_gpu_multi.py (run under torch.distributed.run, one rank per GPU): the NativeEngine with objects
sharded over the ranks, a few frames of a synthetic stream; rank 0 writes what the merged frame looks like."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from emfusion_b200.native import NativeEngine          # noqa: E402
from emfusion_b200.synth import Scene                  # noqa: E402
from emfusion_b200.volume import ObjTSDF, Params       # noqa: E402


def run(out_path, world, rank, dev, n_frames=5, w=320, h=240, bg=96, n_obj=5, obj=32, group=None, replicate=None, peer=None, cert=None):
    scene = Scene(n_objects=n_obj, width=w, height=h, seed=11, dropout=0.01)
    prm = Params(frameSize=(w, h), intr=scene.K, globalVolumeDims=(bg,) * 3, globalVoxelSize=5.12 / bg,
                 objVolumeDims=(obj,) * 3, visibilityThresh=(40 * 40 * w * h) // (640 * 480), boundary=max(2, 20 * w // 640))
    ObjTSDF.nextID = 0
    eng = NativeEngine(prm, dev, rank=rank, world_size=world, group=group, replicate_background=replicate, peer_exchange=peer)
    if cert is not None:
        eng.set_ray_certificate(cert)
    for k in range(n_obj):
        eng.add_object(scene.object_pose(k, 0), scene.object_voxel_size(k, obj))
    res = {}
    for f in range(n_frames):
        depth, inst = scene.render(f)
        d = torch.from_numpy(depth).to(dev)
        eng.processFrame(d, scene.cam_pose(f), {o.id: scene.object_pose(o.id - 1, f) for o in eng.objects})
        if f == 0:
            zeros = torch.zeros((h, w), dtype=torch.uint8, device=dev)
            for o in eng.objects:
                o.integrateMask(torch.from_numpy((inst == o.id).astype(np.uint8)).to(dev), zeros, eng.pose, prm.intr)
        if f > 0 and rank == 0:
            res[f"seg{f}"] = eng.modelSegmentation.cpu().numpy()
            res[f"ray{f}"] = eng.raylengths.cpu().numpy()
            res[f"vert{f}"] = eng.vertices.cpu().numpy()
            res[f"bgassoc{f}"] = eng.bg_associationWeights.cpu().numpy()
            res[f"vis{f}"] = np.array(sorted(eng.vis_objs), dtype=np.int64)
        elif f > 0:
            _ = eng.vis_objs
    if rank == 0:
        res["bg_tsdf"] = eng.background.tsdfVol.cpu().numpy()
    for o in eng.objects:
        res[f"obj{o.id}_tsdf"] = o.tsdfVol.cpu().numpy()
        res[f"obj{o.id}_assoc"] = eng.associationWeights[o.id].cpu().numpy()
    if world > 1:
        px = eng._px
        res["peer_exchange"] = np.array([1 if px is not None else 0])
        if px is not None:
            torch.cuda.synchronize()
            assert px.check_errors() == 0, f"rank {rank}: a peer wait timed out ({px.check_errors()})"
        if peer:
            assert px is not None, "peer exchange was requested but is not active"
    np.savez(out_path + f".rank{rank}.npz", **res)


SIZES = {"small": dict(w=320, h=240, bg=96, n_obj=5, obj=32, n_frames=5),
         # BASELINE sizes: 512^3 background, 128^3 objects, 640 x 480
         "large": dict(w=640, h=480, bg=512, n_obj=8, obj=128, n_frames=4)}


if __name__ == "__main__":
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    arg = lambda i, d: sys.argv[i] if len(sys.argv) > i else d
    run(sys.argv[1], world, rank, dev, replicate={"auto": None, "0": False, "1": True}[arg(2, "auto")],
        peer={"auto": None, "nccl": False, "peer": True}[arg(3, "auto")],
        cert={"auto": None, "0": False, "1": True}[arg(5, "auto")], **SIZES[arg(4, "small")])
    dist.barrier()
    dist.destroy_process_group()
