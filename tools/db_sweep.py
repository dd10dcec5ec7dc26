"""Development aid: DB batch-256 step time for the sub-batch / priority tuning knobs."""
import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
from pytorchocr_b200 import synth, _lib
from pytorchocr_b200.postprocess import build_post_process
L = _lib.lib()
N = 256
base = torch.from_numpy(synth.db_batch(16)).cuda()
maps = base.repeat(N // 16, 1, 1, 1).contiguous()
op = build_post_process({"name": "DBPostProcess", "thresh": 0.3, "box_thresh": 0.5, "unclip_ratio": 1.7, "cpp_speedup": True, "cuda_speedup": True})
sl = np.array([[736, 1280, 1.0, 1.0]] * N)
op.run_device(maps, sl)
buf = next(iter(op._cache.values())); key = next(iter(op._cache))
o_box, o_sc, o_cnt, o_st = buf["offs"]; b0 = buf["out_dev"].data_ptr()
stream = torch.cuda.current_stream()
def step():
    _lib.check(L.ocrpp_db_postprocess(maps.data_ptr(), _lib.F32, N, 736, 1280, maps.stride(0), maps.stride(2), buf["wh_dev"].data_ptr(),
        0.3, 0.5, 1.7, key[5], key[4], 0, 0, b0 + o_box, b0 + o_sc, b0 + o_cnt, b0 + o_st, None, None,
        buf["ws"].data_ptr(), buf["ws_bytes"], stream.cuda_stream))
    buf["out_host"].copy_(buf["out_dev"], non_blocking=True)
def timeit(k=20):
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [step() for _ in range(k)]; e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
def phases():
    L.ocrpp_profile_enable(1); step(); torch.cuda.synchronize(); L.ocrpp_profile_reset()
    for _ in range(5): step()
    torch.cuda.synchronize(); L.ocrpp_profile_enable(0)
    calls, ph = _lib.profile_read()
    return {k: round(v / calls, 4) for k, v in ph}
_lib.check(L.ocrpp_set_tuning(2, 1))
for scan in (0, 1, 2):
    _lib.check(L.ocrpp_set_tuning(3, scan))
    for split in (1, 2):
        _lib.check(L.ocrpp_set_tuning(1, split))
        ms = timeit()
        print("scan-tuning %d split %d: %.4f ms  %.0f img/s  whole-step %.3f" % (scan, split, ms, N / ms * 1e3, N * 736 * 1280 * 4 / (ms * 1e-3) / 6546.2e9))
    _lib.check(L.ocrpp_set_tuning(1, 1))
    print("   phases", phases())
_lib.check(L.ocrpp_set_tuning(3, 0))
# db_scan4_kernel: ring slots x CTAs per SM
for stages in (8, 16, 24, 32):
    for ctas in (1, 2, 3, 4):
        if stages * 5120 * ctas > 200 * 1024: continue
        _lib.check(L.ocrpp_set_tuning(4, stages)); _lib.check(L.ocrpp_set_tuning(5, ctas))
        _lib.check(L.ocrpp_set_tuning(1, 1))
        ph = phases()
        _lib.check(L.ocrpp_set_tuning(1, 2))
        ms = timeit()
        print("scan4 stages %d ctas/SM %d: scan %.4f ms | step(split 2) %.4f ms %.0f img/s" % (stages, ctas, ph["db_scan"], ms, N / ms * 1e3))
