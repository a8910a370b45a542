"""-m gpu, needs >= 2 GPUs (skipped otherwise): objects sharded over 2 ranks with the NCCL normaliser all-reduce and the
gathered, merged composite (NativeEngine, world_size 2) against the single-GPU engine on the same stream: segmentation,
ray lengths, vertices and visibility bit-exact; association within 1e-5 (the all-reduce changes the summation order);
integrated volumes within 1e-4."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("exchange", ["nccl", "peer"], ids=["nccl_collectives", "nvlink_peer_memory"])
@pytest.mark.parametrize("replicate", ["0", "1"], ids=["background_on_rank0", "background_replicated"])
def test_two_rank_engine_equals_single_gpu(tmp_path, replicate, exchange):
    from tests import mgpu_check
    out2 = str(tmp_path / "w2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "2953" + replicate, os.path.join(ROOT, "tests", "mgpu_check.py"), out2, replicate, exchange]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    out1 = str(tmp_path / "w1")
    mgpu_check.run(out1, 1, 0, torch.device("cuda", 0))
    a = np.load(out1 + ".rank0.npz")
    b0, b1 = np.load(out2 + ".rank0.npz"), np.load(out2 + ".rank1.npz")
    assert int(b0["peer_exchange"][0]) == (1 if exchange == "peer" else 0)
    for f in range(1, 5):
        assert np.array_equal(a[f"seg{f}"], b0[f"seg{f}"]), f"frame {f}: segmentation"
        assert np.array_equal(a[f"ray{f}"], b0[f"ray{f}"]), f"frame {f}: ray lengths"
        assert np.array_equal(a[f"vert{f}"], b0[f"vert{f}"]), f"frame {f}: vertices"
        assert np.array_equal(a[f"vis{f}"], b0[f"vis{f}"]), f"frame {f}: visibility"
        assert float(np.abs(a[f"bgassoc{f}"] - b0[f"bgassoc{f}"]).max()) <= 1e-5
    assert float(np.abs(a["bg_tsdf"] - b0["bg_tsdf"]).max()) <= 1e-4
    seen = 0
    for k in range(1, 6):
        key = f"obj{k}_tsdf"
        src = b0 if key in b0.files else b1
        assert key in src.files, f"object {k} is owned by no rank"
        assert float(np.abs(a[key] - src[key]).max()) <= 1e-4
        assert float(np.abs(a[f"obj{k}_assoc"] - src[f"obj{k}_assoc"]).max()) <= 1e-5
        seen += 1
    assert seen == 5
