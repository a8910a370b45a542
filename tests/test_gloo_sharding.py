"""CPU, world_size 2 over gloo: the multi-GPU decomposition of the path on the C oracle --
 * objects sharded round-robin, rank 0 also owns the background;
 * association: per-rank un-normalised weights + partial per-pixel normaliser, ONE all-reduce (sum), local divide
   == the single-process normalised weights (<= 1e-6: only the summation order changes);
 * composite: per-rank pre-composite of the rank's own objects against an empty background, ONE gather to rank 0, merge in
   (raylength, list index) order + background rule == the reference's sequential composite over all objects, bit-exact
   (the merge rule is the one emf_composite_merge implements on the GPU)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def merge_numpy(parts, ids, bg, boundary, w, h):
    """numpy restatement of k_composite_merge (emfusion_b200/csrc/raycast.cu)"""
    lut = {min(i, 255): k for k, i in reversed(list(enumerate(ids)))}
    r = np.zeros((h, w), np.float32); best = np.full((h, w), 1 << 30); win = np.full((h, w), -1); seg = np.zeros((h, w), np.uint8)
    for p, (ray, vert, norm, sg) in enumerate(parts):
        has = sg != 0
        idx = np.vectorize(lambda v: lut.get(int(v), 1 << 29))(sg)
        take = has & ((win < 0) | ((r <= 0) & (idx > best)) | (ray < r) | ((ray == r) & (idx < best) & ~(r <= 0)))
        r = np.where(take, ray, r); best = np.where(take, idx, best); win = np.where(take, p, win); seg = np.where(take, sg, seg)
    bgm = bg["mask"] != 0
    seg = np.where(bgm & ((r - bg["ray"]) > np.float32(0.05)), 0, seg).astype(np.uint8)
    vert = np.zeros((h, w, 3), np.float32)
    for p, (ray, v, n, sg) in enumerate(parts):
        vert = np.where(((win == p) & (seg != 0))[..., None], v, vert)
    vert = np.where(((seg == 0) & bgm)[..., None], bg["vert"], vert)
    return dict(ray=r, seg=seg, vert=vert)


def worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import oracle_c, scenario as S
    from emfusion_b200.poses import rel_pose_CO
    o = oracle_c.load()
    w, h = 160, 120
    sc = S.make("gloo", o, w, h, (48, 48, 48), 5, (24, 24, 24), n_frames=2, integrate_frames=1, seed=4)
    cam = sc.cam(1)
    pts = o.compute_points(sc.depths[1], sc.K)
    vols = sc.vols()                                   # [bg, obj1..obj5], list order
    mine = [i for i, v in enumerate(vols) if (i == 0 and rank == 0) or (i > 0 and (i - 1) % world == rank)]
    # ---- association
    imgs = {}
    part = np.zeros((h, w), np.float32)
    for i in mine:
        v = vols[i]
        T = rel_pose_CO(cam, v.pose)
        a, _ = o.assoc_volume(v.tsdf, v.fg_probs, pts, S.R9(T), S.T3(T), v.res, v.voxel, v.trunc)
        imgs[i] = a
        part = part + a
    norm = torch.from_numpy(part.copy())
    dist.all_reduce(norm, op=dist.ReduceOp.SUM)
    norm = norm.numpy()
    for i in mine:
        imgs[i] = np.where(norm != 0, imgs[i] / np.where(norm != 0, norm, 1), 0).astype(np.float32)
    # ---- raycast + per-rank pre-composite against an empty background
    rc = {}
    for i in mine:
        v = vols[i]
        T = rel_pose_CO(cam, v.pose)
        g = o.compute_grads(v.tsdf, v.res)
        wt = v.weights if v.fg_probs is None else o.raycast_weights(v.weights, (v.fg_probs > 0.5).astype(np.uint8) * 255)
        rc[i] = o.raycast(v.tsdf, g, wt, S.R9(T), S.T3(T), sc.K, v.res, v.voxel, v.trunc, w, h)
    objs = [i for i in mine if i > 0]
    z1, z3, zm = np.zeros((h, w), np.float32), np.zeros((h, w, 3), np.float32), np.zeros((h, w), np.uint8)
    pre = o.composite([vols[i].vid for i in objs], [rc[i]["ray"] for i in objs], [rc[i]["vert"] for i in objs],
                      [rc[i]["norm"] for i in objs], [rc[i]["mask"] for i in objs], z1, z3, z3, zm, 4)
    packed = torch.from_numpy(np.concatenate([pre["ray"].ravel(), pre["vert"].ravel(), pre["norm"].ravel(),
                                              pre["seg"].ravel().astype(np.float32)]))
    bufs = [torch.empty_like(packed) for _ in range(world)] if rank == 0 else None
    dist.gather(packed, bufs, dst=0)
    np.savez(os.path.join(tmp, f"rank{rank}.npz"), **{f"assoc{i}": imgs[i] for i in mine})
    if rank == 0:
        n = h * w
        parts = []
        for b in bufs:
            b = b.numpy()
            parts.append((b[:n].reshape(h, w), b[n:4 * n].reshape(h, w, 3), b[4 * n:7 * n].reshape(h, w, 3),
                          b[7 * n:].reshape(h, w).astype(np.uint8)))
        merged = merge_numpy(parts, [v.vid for v in vols[1:]], rc[0], 4, w, h)
        np.savez(os.path.join(tmp, "merged.npz"), **merged)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_decomposition_matches_single_process(tmp_path, oracle):
    from tests import scenario as S
    from emfusion_b200.poses import rel_pose_CO
    port = 29600 + os.getpid() % 300
    mp.spawn(worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    # single process, the reference's order
    o = oracle
    w, h = 160, 120
    sc = S.make("gloo", o, w, h, (48, 48, 48), 5, (24, 24, 24), n_frames=2, integrate_frames=1, seed=4)
    cam = sc.cam(1)
    pts = o.compute_points(sc.depths[1], sc.K)
    imgs, rc = [], []
    for v in sc.vols():
        T = rel_pose_CO(cam, v.pose)
        a, _ = o.assoc_volume(v.tsdf, v.fg_probs, pts, S.R9(T), S.T3(T), v.res, v.voxel, v.trunc)
        imgs.append(a)
        g = o.compute_grads(v.tsdf, v.res)
        wt = v.weights if v.fg_probs is None else o.raycast_weights(v.weights, (v.fg_probs > 0.5).astype(np.uint8) * 255)
        rc.append(o.raycast(v.tsdf, g, wt, S.R9(T), S.T3(T), sc.K, v.res, v.voxel, v.trunc, w, h))
    o.normalise(imgs)
    comp = o.composite([v.vid for v in sc.objs], [r["ray"] for r in rc[1:]], [r["vert"] for r in rc[1:]],
                       [r["norm"] for r in rc[1:]], [r["mask"] for r in rc[1:]], rc[0]["ray"], rc[0]["vert"], rc[0]["norm"],
                       rc[0]["mask"], 4)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    owned = set()
    for i in range(len(imgs)):
        src = r0 if f"assoc{i}" in r0.files else r1
        assert f"assoc{i}" in src.files
        owned.add(i)
        assert float(np.abs(src[f"assoc{i}"] - imgs[i]).max()) <= 1e-6, f"association image {i}"
    assert owned == set(range(6)) and "assoc0" in r0.files and "assoc2" in r1.files and "assoc1" in r0.files
    m = np.load(tmp_path / "merged.npz")
    assert np.array_equal(m["seg"], comp["seg"])
    assert np.array_equal(m["ray"], comp["ray"])
    assert np.array_equal(m["vert"], comp["vert"])
    assert int((comp["seg"] != 0).sum()) > 50
