"""GPU: the sharding wrapper drives the CUDA operators (single process = world 1 here; the N-GPU run is
bench.py --gpus N, which shards the same way with one process per GPU)."""
import numpy as np
import pytest

from pytorchocr_b200 import sharding, synth

pytestmark = pytest.mark.gpu


def test_sharded_wrapper_single_process():
    import torch
    from pytorchocr_b200.postprocess import build_post_process
    H, W, N = 96, 160, 4
    maps = synth.db_batch(N, seed=3, H=H, W=W)
    sl = np.array([[H, W, 1.0, 1.0]] * N)
    op = build_post_process({"name": "DBPostProcess", "thresh": 0.3, "box_thresh": 0.5, "unclip_ratio": 1.7,
                             "cuda_speedup": True})
    full = op({"maps": torch.from_numpy(maps).cuda()}, sl)
    parts = []
    for r in range(2):   # what two ranks would each compute
        m, s = sharding.shard_detection_inputs(torch.from_numpy(maps).cuda(), sl, r, 2)
        parts.extend(op({"maps": m}, s))
    assert len(parts) == N
    for a, b in zip(parts, full):
        assert np.array_equal(a["points"], b["points"])
    res = sharding.ShardedPostProcess(op)({"maps": torch.from_numpy(maps).cuda()}, sl)
    assert all(np.array_equal(a["points"], b["points"]) for a, b in zip(res, full))
