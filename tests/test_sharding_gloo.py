"""Multi-process (gloo, world_size 2, CPU) test of the sharding host logic. The per-rank operator is
the CPU oracle here (no GPU in this container); on the GPU box the same wrapper drives the CUDA
operators (tests/test_sharding_gpu.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pytorchocr_b200 import sharding, synth


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 128, 65536):
        for world in (1, 2, 3, 8):
            b = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.ctc_oracle import CTCLabelDecodeOracle
        from oracle.db_oracle import DBPostProcessOracle
        H, W, N = 96, 160, 5
        maps = synth.db_batch(N, seed=3, H=H, W=W)
        shape_list = np.array([[H, W, 1.0, 1.0]] * N)
        op = DBPostProcessOracle(thresh=0.3, box_thresh=0.5, unclip_ratio=1.7)
        # full batch on every rank, results everywhere
        res_all = sharding.ShardedPostProcess(op)({"maps": maps}, shape_list)
        # pre-sharded (DistributedSampler style), results on rank 0 only
        lo, hi = sharding.shard_bounds(N, rank, world)
        res_dst = sharding.ShardedPostProcess(op, dst=0, presharded=True)({"maps": maps[lo:hi]}, shape_list[lo:hi])
        # CTC: shard dimension 1 of [T,B,C]
        d = synth.write_char_dict(os.path.join(out_dir, "dict_%d.txt" % rank), 96)
        probs, _ = synth.ctc_probs_numpy(5, 12, 7, 97)
        txt = sharding.ShardedPostProcess(CTCLabelDecodeOracle(d))(torch.from_numpy(probs))
        if rank == 0:
            full = op({"maps": maps}, shape_list)
            assert len(res_all) == N and len(res_dst) == N
            for a, b, c in zip(res_all, res_dst, full):
                assert np.array_equal(a["points"], c["points"]) and np.array_equal(b["points"], c["points"])
            full_txt = CTCLabelDecodeOracle(d)(torch.from_numpy(probs))
            assert [t[0] for t in txt] == [t[0] for t in full_txt]
        else:
            assert res_dst is None and len(res_all) == N
        open(os.path.join(out_dir, "ok_%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_operator_gloo_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok_%d" % r)) for r in range(world))
