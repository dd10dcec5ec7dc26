"""Development aid: DB batch-256 step time over sub-batch count x image-kernel shared memory x scan ring size."""
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
def tune(**kw):
    keys = {"split": 1, "prio": 2, "scan": 3, "stages": 4, "ctas": 5, "img_kb": 6}
    for k, v in kw.items(): _lib.check(L.ocrpp_set_tuning(keys[k], v))
tune(prio=1)
for img_kb in (0, 176, 160):
    tune(img_kb=img_kb, split=1, stages=0, ctas=0)
    print("img_kb %d phases %s" % (img_kb, phases()), flush=True)
    for stages, ctas in ((16, 2), (8, 2), (8, 1), (16, 1)):
        res = []
        for split in (2, 3, 4):
            tune(split=split, stages=stages, ctas=ctas)
            res.append("split %d: %.4f ms" % (split, timeit()))
        print("  stages %d ctas %d | %s" % (stages, ctas, " | ".join(res)), flush=True)
tune(prio=0, img_kb=160, stages=8, ctas=2)
print("with priorities:", " | ".join("split %d: %.4f ms" % (sp, (tune(split=sp), timeit())[1]) for sp in (2, 3, 4)))
