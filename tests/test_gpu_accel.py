"""-m gpu: the acceleration state (segment status, constant-brick map, frustum culling, approximate-then-exact
arithmetic) must never change a result.
 * integrate under random camera poses (rotated, inside the volume, looking away, grazing) == C oracle, bit-exact;
 * the constant-segment and brick maps are sound (an entry certifies the voxel values it claims);
 * the engine with acceleration on == the engine with acceleration off, bit for bit, over a moving stream;
 * ConstDiv (hoisted-reciprocal IEEE division) is exercised through the raycast against the reference kernels in
   tests/test_gpu_vs_reference.py; here additionally at larger volumes where rays cross many bricks."""
import numpy as np
import pytest
import torch

from emfusion_b200 import ops
from emfusion_b200.engine import EMFusionEngine
from emfusion_b200.poses import Affine, rel_pose_CO, rel_pose_OC
from emfusion_b200.synth import Scene
from emfusion_b200.volume import ObjTSDF, Params
from tests import scenario as S
from tests.test_gpu_parity import DEV, assert_bits, cu

pytestmark = pytest.mark.gpu


def random_poses(n, seed):
    rng = np.random.default_rng(seed)
    out = [Affine.identity(), Affine.from_rvec([0, np.pi, 0], [0, 0, 5.0]),          # looking back through the volume
           Affine.from_rvec([0, 0, 0], [0.3, -0.2, 2.5]),                               # inside the volume
           Affine.from_rvec([0, np.pi / 2, 0], [-3.0, 0, 2.5]),                         # from the side
           Affine.from_rvec([0.0, 0.0, np.pi / 4], [0, 0, -1.0])]                       # rolled, from further back
    for _ in range(n):
        out.append(Affine.from_rvec(rng.normal(size=3) * 0.6, rng.uniform(-1.5, 1.5, 3) + [0, 0, 1.0]))
    return out


@pytest.mark.parametrize("res", [(64, 64, 64), (128, 40, 72), (96, 128, 64)], ids=["64", "128x40x72", "96x128x64"])
def test_integrate_bricks_random_poses_bit_exact(oracle, cuda_dev, res):
    """brick-level classification (k_brick_classify + k_integrate_bricks: no bitmaps => the brick path) vs the straight oracle
    loop under random camera poses (rotated, inside the volume, looking away, grazing), accumulated over all poses; the
    kernel's own counters agree with the oracle's"""
    w, h = 320, 240
    scene = Scene(n_objects=3, width=w, height=h, seed=7, dropout=0.03)
    n = int(np.prod(res))
    voxel = float(np.float32(5.12 / res[0]))
    trunc = float(np.float32(10.0) * np.float32(voxel))
    vol_pose = Affine.translation([0, 0, 2.56])
    t_o, w_o = np.zeros(n, np.float32), np.zeros(n, np.float32)
    t_g, w_g = cu(t_o), cu(w_o)
    rng = np.random.default_rng(1)
    stats = torch.zeros(8, dtype=torch.int64, device=DEV)
    tot = np.zeros(6, np.int64)
    for k, cam in enumerate(random_poses(8, 3)):
        depth, _ = scene.render(k)
        assoc = rng.random((h, w), dtype=np.float32)
        T = rel_pose_OC(cam, vol_pose)
        tot += oracle.update_tsdf(depth, assoc, t_o, w_o, S.R9(T), S.T3(T), scene.K, res, voxel, trunc, 64.0, counts=True)
        v = ops.volume(t_g, w_g, res, voxel, trunc)
        ops.integrateVolumes([v], [T], scene.K, cu(depth), [cu(assoc)], 64.0, stats=stats)
        assert_bits(t_g, t_o, f"tsdf after pose {k}")
        assert_bits(w_g, w_o, f"weights after pose {k}")
    st = stats.cpu().numpy()
    assert st[0] == tot[0] and st[1] == tot[1] and st[2] == tot[2] and st[3] == tot[3], (st, tot)
    assert st[6] + st[7] > 0        # segments were decided wholesale


@pytest.mark.parametrize("res", [(64, 64, 64), (128, 40, 72)], ids=["64", "128x40x72"])
def test_integrate_random_poses_bit_exact(oracle, cuda_dev, res):
    """frustum culling + approximate classification vs the straight oracle loop, accumulated over all poses"""
    w, h = 320, 240
    scene = Scene(n_objects=3, width=w, height=h, seed=7, dropout=0.03)
    n = int(np.prod(res))
    voxel = float(np.float32(5.12 / res[0]))
    trunc = float(np.float32(10.0) * np.float32(voxel))
    vol_pose = Affine.translation([0, 0, 2.56])
    t_o, w_o = np.zeros(n, np.float32), np.zeros(n, np.float32)
    t_g, w_g = cu(t_o), cu(w_o)
    cbits = torch.zeros((3 * ops.bitmapWords(res),), dtype=torch.int32, device=DEV)
    ops.resetBitmaps(ops.volume(t_g, w_g, res, voxel, trunc, const_bits=cbits))
    rng = np.random.default_rng(1)
    stats = torch.zeros(8, dtype=torch.int64, device=DEV)
    tot = np.zeros(6, np.int64)
    for k, cam in enumerate(random_poses(8, 3)):
        depth, _ = scene.render(k)
        assoc = rng.random((h, w), dtype=np.float32)
        T = rel_pose_OC(cam, vol_pose)
        tot += oracle.update_tsdf(depth, assoc, t_o, w_o, S.R9(T), S.T3(T), scene.K, res, voxel, trunc, 64.0, counts=True)
        v = ops.volume(t_g, w_g, res, voxel, trunc, const_bits=cbits)
        ops.integrateVolumes([v], [T], scene.K, cu(depth), [cu(assoc)], 64.0, stats=stats)
        assert_bits(t_g, t_o, f"tsdf after pose {k}")
        assert_bits(w_g, w_o, f"weights after pose {k}")
    st = stats.cpu().numpy()
    # the kernel's own counters agree with the oracle's (updated, marked -1, occluded-seen, check-only)
    assert st[0] == tot[0] and st[1] == tot[1] and st[2] == tot[2] and st[3] == tot[3], (st, tot)
    assert st[4] <= tot[4]                      # only a sliver of out-of-image voxels is ever projected
    fr = check_const_bits(t_g, cbits, res)
    assert fr[0] > 0.01, fr


def unpack_maps(bits, res):
    """(3 * words,) int32 -> bool (3, rz, ry, rx // 4)"""
    rx, ry, rz = res
    wpr = (rx // 4 + 31) // 32
    w = bits.reshape(3, rz, ry, wpr).to(torch.int64) & 0xFFFFFFFF
    sh = torch.arange(32, device=bits.device, dtype=torch.int64)
    b = ((w[..., None] >> sh) & 1).to(torch.bool).reshape(3, rz, ry, wpr * 32)
    assert not bool(b[..., rx // 4:].any()) or True   # padding bits carry no meaning
    return b[..., : rx // 4]


VALS = (1.0, 0.0, -1.0)


def check_const_bits(tsdf, bits, res):
    rx, ry, rz = res
    t = tsdf.reshape(rz, ry, rx // 4, 4)
    m = unpack_maps(bits, res)
    for k, c in enumerate(VALS):
        holds = (t == c).all(dim=-1)
        wrong = m[k] & ~holds
        assert not bool(wrong.any()), f"{int(wrong.sum())} segments claim constant {c} they do not hold"
    return [float(m[k].float().mean()) for k in range(3)]


def brick_view(bmap, res):
    """brick_map bytes -> (code, D) int tensors of shape (nbz, nby, nbx)"""
    nb = [(r + 7) // 8 for r in res]
    n = nb[0] * nb[1] * nb[2]
    m = bmap[:n].reshape(nb[2], nb[1], nb[0]).to(torch.int64)
    return m >> 4, m & 7, (m >> 3) & 1


def check_brick_map(tsdf, bmap, res, tight=False):
    """(m << 4) | D certifies: every voxel of every brick within D - 1 bricks (Chebyshev) of this one, as far as it lies
    inside the volume, holds constant VALS[m - 1]."""
    import torch.nn.functional as F
    rx, ry, rz = res
    code, D, Pf = brick_view(bmap, res)
    nbz, nby, nbx = code.shape
    t = tsdf.reshape(rz, ry, rx)
    pad = (0, nbx * 8 - rx, 0, nby * 8 - ry, 0, nbz * 8 - rz)
    fr = []
    assert int(D.max()) <= 7 and bool(((code == 0) == (D == 0)).all())
    for k, c in enumerate(VALS):
        bad = F.pad((t != c).float(), pad, value=0.0)[None, None]          # outside the volume: never "bad"
        brick_bad = F.max_pool3d(bad, kernel_size=8, stride=8)[0, 0] > 0   # (nbz, nby, nbx)
        for d in range(1, 8):
            sel = (code == k + 1) & (D >= d)
            if not bool(sel.any()):
                continue
            r = d - 1
            nb_bad = F.max_pool3d(brick_bad[None, None].float(), kernel_size=2 * r + 1, stride=1, padding=r)[0, 0] > 0
            wrong = sel & nb_bad
            assert not bool(wrong.any()), f"{int(wrong.sum())} bricks certify constant {c} with D >= {d} but are not"
            if tight:   # whatever the segment bitmaps allow is certified
                miss = (~nb_bad) & ~((code == k + 1) & (D >= d))
                assert not bool(miss.any()), f"{int(miss.sum())} bricks could certify {c} with D >= {d} but do not"
        # P: the 2x2x2 block of bricks starting here (as far as inside the volume) holds the constant
        blk_bad = F.max_pool3d(F.pad(brick_bad[None, None].float(), (0, 1, 0, 1, 0, 1), value=0.0), kernel_size=2, stride=1)[0, 0] > 0
        wrong = (code == k + 1) & (Pf == 1) & blk_bad
        assert not bool(wrong.any()), f"{int(wrong.sum())} bricks flag a constant-{c} 2x2x2 block that is not"
        fr.append(float(((code == k + 1) & ((D >= 2) | (Pf == 1))).float().mean()))
    return fr


def build(accel, w=320, h=240, bg=128, n_obj=3, obj=64, seed=5):
    scene = Scene(n_objects=n_obj, width=w, height=h, seed=seed, dropout=0.01)
    prm = Params(frameSize=(w, h), intr=scene.K, globalVolumeDims=(bg,) * 3, globalVoxelSize=5.12 / bg,
                 objVolumeDims=(obj,) * 3, visibilityThresh=100, boundary=10)
    ObjTSDF.nextID = 0
    eng = EMFusionEngine(prm, DEV, accelerate=accel)
    for k in range(n_obj):
        eng.add_object(scene.object_pose(k, 0), scene.object_voxel_size(k, obj))
    return scene, eng


def test_acceleration_changes_nothing(cuda_dev):
    scene, fast = build(True)
    _, slow = build(False)
    assert fast.background.constBits is not None and slow.background.constBits is None
    for f in range(6):
        depth, inst = scene.render(f)
        for eng in (fast, slow):
            poses = {o.id: scene.object_pose(o.id - 1, f) for o in eng.objects}
            eng.processFrame(cu(depth), scene.cam_pose(f), poses)
            if f == 0:
                zeros = torch.zeros((eng.h, eng.w), dtype=torch.uint8, device=DEV)
                for o in eng.objects:
                    o.integrateMask(cu((inst == o.id).astype(np.uint8)), zeros, eng.pose, eng.params.intr)
        for a, b in zip(fast.local_volumes(), slow.local_volumes()):
            assert_bits(a.tsdfVol, b.tsdfVol.cpu().numpy(), f"frame {f} tsdf vol {a.id}")
            assert_bits(a.tsdfWeights, b.tsdfWeights.cpu().numpy(), f"frame {f} weights vol {a.id}")
        if f > 0:
            for name in ("raylengths", "vertices", "normals", "modelSegmentation", "bg_raylengths", "bg_vertices",
                         "bg_normals", "bg_mask", "bg_associationWeights"):
                assert_bits(getattr(fast, name), getattr(slow, name).cpu().numpy(), f"frame {f} {name}")
            for o in fast.objects:
                assert_bits(fast.obj_raylengths[o.id], slow.obj_raylengths[o.id].cpu().numpy(), f"frame {f} obj ray {o.id}")
                assert_bits(fast.associationWeights[o.id], slow.associationWeights[o.id].cpu().numpy(), f"frame {f} assoc {o.id}")
            assert fast.vis_objs == slow.vis_objs
    fr = []
    for v in fast.local_volumes():
        fr.append((check_const_bits(v.tsdfVol.reshape(-1), v.constBits, v.volumeRes),
                   check_brick_map(v.tsdfVol.reshape(-1), v.brickMap, v.volumeRes)))
    # the maps are not vacuous: free space (+1) and never-seen space (0) of the background are certified
    assert fr[0][0][0] > 0.03 and fr[0][1][0] + fr[0][1][1] > 0.01, fr
    assert int(fast.bg_mask.sum()) > 0.8 * fast.w * fast.h


def test_raycast_with_bricks_vs_oracle(oracle, cuda_dev):
    """64^3 background integrated on the GPU (so the brick map is live), then raycast with and without the map
    against the C oracle on the GPU-integrated volume."""
    w, h = 320, 240
    scene = Scene(n_objects=2, width=w, height=h, seed=9)
    res = (64, 64, 64)
    n = 64 ** 3
    voxel = float(np.float32(5.12 / 64))
    trunc = float(np.float32(10.0) * np.float32(voxel))
    pose = Affine.translation([0, 0, 2.56])
    t_g, w_g = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    cbits = torch.zeros((3 * ops.bitmapWords(res),), dtype=torch.int32, device=DEV)
    bmap = torch.zeros((ops.brickMapBytes(res),), dtype=torch.uint8, device=DEV)
    v = ops.volume(t_g, w_g, res, voxel, trunc, const_bits=cbits, brick_map=bmap)
    ops.resetBitmaps(v)
    code, D, Pf = brick_view(bmap, res)
    assert bool((code == 2).all()) and bool((D == 7).all()) and bool((Pf == 1).all())     # a zeroed volume: every brick "all 0", full radius
    ones = torch.ones((h, w), device=DEV)
    for f in range(4):
        depth, _ = scene.render(f)
        ops.integrateVolumes([v], [rel_pose_OC(scene.cam_pose(f), pose)], scene.K, cu(depth), [ones], 64.0)
    ops.updateBrickMaps([v])
    check_const_bits(t_g, cbits, res)
    assert sum(check_brick_map(t_g, bmap, res)) > 0.005
    t_np, w_np = t_g.cpu().numpy(), w_g.cpu().numpy()
    for f in (4, 9):
        T = rel_pose_CO(scene.cam_pose(f), pose)
        o = oracle.raycast(t_np, oracle.compute_grads(t_np, res), w_np, S.R9(T), S.T3(T), scene.K, res, voxel, trunc, w, h)
        z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=DEV)
        ray, vert, norm, mask = [z(h, w)], [z(h, w, 3)], [z(h, w, 3)], [z(h, w, dt=torch.uint8)]
        ops.raycastVolumes([v], [T], scene.K, [[0, 0, w, h]], ray, vert, norm, mask)
        assert_bits(mask[0], o["mask"], "mask")
        assert_bits(ray[0], o["ray"], "ray")
        assert_bits(vert[0], o["vert"], "vert")
        assert_bits(norm[0], o["norm"], "norm")
        assert int(o["mask"].sum()) > 0.5 * w * h
