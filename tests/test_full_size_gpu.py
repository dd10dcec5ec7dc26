"""GPU: BASELINE.json's full batch sizes, checked through size-independent properties (the oracle only
finishes small samples in seconds): batch independence / determinism (a map gives the same boxes wherever
it sits in the batch and however often the call is repeated), bounds, and the oracle on a sample."""
import numpy as np
import pytest

from pytorchocr_b200 import synth

pytestmark = pytest.mark.gpu


def _rows(r):
    return sorted(map(tuple, np.asarray(r["points"], np.int16).reshape(-1, 8).tolist()))


def test_db_batch_256_full_size():
    import torch
    from oracle.db_oracle import DBPostProcessOracle
    from pytorchocr_b200.postprocess import build_post_process
    H, W, U, N = 736, 1280, 8, 256
    uniq = synth.db_batch(U, seed=synth.BASE_SEED + 1000)
    idx = np.arange(N) % U
    rng = np.random.default_rng(0)
    rng.shuffle(idx)
    maps = torch.from_numpy(uniq).cuda()[torch.from_numpy(idx).cuda()]            # [256,1,736,1280] on the device
    sl = np.array([[H, W, 1.0, 1.0]] * N)
    cfg = dict(name="DBPostProcess", thresh=0.3, box_thresh=0.5, max_candidates=1000, unclip_ratio=1.7, cpp_speedup=True,
               cuda_speedup=True)
    op = build_post_process(cfg)
    res = op({"maps": maps}, sl)
    res2 = op({"maps": maps}, sl)
    first = {}
    for n in range(N):
        rows = _rows(res[n])
        assert rows == _rows(res2[n])                         # deterministic
        first.setdefault(int(idx[n]), rows)
        assert rows == first[int(idx[n])]                     # independent of the position in the batch
        pts = np.asarray(res[n]["points"])
        assert 150 <= len(pts) <= 260
        assert pts[..., 0].min() >= 0 and pts[..., 0].max() <= W and pts[..., 1].min() >= 0 and pts[..., 1].max() <= H
    want = DBPostProcessOracle(thresh=0.3, box_thresh=0.5, unclip_ratio=1.7, cpp_speedup=True)({"maps": uniq}, sl[:U])
    for u in range(U):                                        # all 8 distinct maps of the batch against the oracle
        a, b = np.array(first[u]), np.array(_rows(want[u]))
        assert a.shape == b.shape
        diff = np.abs(a - b).max(1)
        assert (diff > 0).sum() <= 2 and diff.max() <= 1     # boxes on a rounding discontinuity


@pytest.mark.parametrize("kind", ["pse", "pan"])
def test_expand_batch_full_size(kind):
    import torch
    from pytorchocr_b200.postprocess import build_post_process
    H, W, U, N = 736, 1280, 4, 32
    if kind == "pse":
        uniq = np.stack([synth.pse_maps(synth.BASE_SEED + 2000 + i) for i in range(U)])
        cfg = dict(name="PSEPostProcess", thresh=0, box_thresh=0.85, min_area=16, scale=1)
    else:
        uniq = np.stack([synth.pan_maps(synth.BASE_SEED + 3000 + i) for i in range(U)])
        cfg = dict(name="PANPostProcess", thresh=0, box_thresh=0.85, min_area=16, min_kernel_area=2.6, scale=1)
    idx = np.arange(N) % U
    np.random.default_rng(1).shuffle(idx)
    maps = torch.from_numpy(uniq).cuda()[torch.from_numpy(idx).cuda()]
    sl = np.array([[H, W, 1.0, 1.0]] * N)
    op = build_post_process(dict(cfg, cuda_speedup=True, maps_at_processing_res=True))
    boxes, scores, counts, status, ex = op.run_device(maps, sl, labels=True)
    boxes2, scores2, counts2, _, ex2 = op.run_device(maps, sl, labels=True)
    assert np.array_equal(counts, counts2) and np.array_equal(ex["labels"], ex2["labels"])      # deterministic
    first = {}
    for n in range(N):
        u = int(idx[n])
        k = int(counts[n])
        cur = (ex["labels"][n], boxes[n, :k].copy(), scores[n, :k].copy())
        if u not in first:
            first[u] = cur
            continue
        assert np.array_equal(cur[0], first[u][0])                # label maps: bit-exact, batch independent
        assert np.array_equal(cur[1], first[u][1]) and np.array_equal(cur[2], first[u][2])
        assert 150 <= k <= 300
        # every labelled pixel is a text pixel and every label id is a seed id (monotone, >= 1)
        assert ((cur[0] > 0) <= (uniq[u, 0] > 0)).all()


def test_ctc_8192_lines():
    import torch
    from oracle.ctc_oracle import CTCLabelDecodeOracle
    from pytorchocr_b200.postprocess import build_post_process
    import tempfile, os
    T, B, C = 80, 8192, 6623
    d = synth.write_char_dict(os.path.join(tempfile.mkdtemp(), "dict.txt"), C)
    probs = synth.ctc_probs_torch(synth.BASE_SEED, T, B, C, torch.device("cuda"))
    op = build_post_process({"name": "CTCLabelDecode", "character_dict_path": d, "cuda_speedup": True})
    full = op(probs)
    assert len(full) == B
    # decoding a shard gives the same lines as decoding everything (no cross-line state)
    part = op(probs[:, 4096:4096 + 512].contiguous())
    assert [t for t, _ in part] == [t for t, _ in full[4096:4096 + 512]]
    sel = np.r_[0:512, 8192 - 512:8192]                        # 1024 lines: both ends of the batch
    sample = probs[:, torch.from_numpy(sel).cuda()].cpu()
    want = CTCLabelDecodeOracle(d)(sample)
    assert [t for t, _ in want] == [full[i][0] for i in sel]
    assert np.allclose([c for _, c in want], [full[i][1] for i in sel], rtol=1e-6, equal_nan=True)
