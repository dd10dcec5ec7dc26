"""Quick device-side timing of the CTC kernels (development aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorchocr_b200 import _lib, synth

T, B, C = 80, int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 6623
dev = torch.device("cuda:0")
x = synth.ctc_probs_torch(1, T, B, C, dev)
L = _lib.lib()
idx = torch.empty(B * T + B, dtype=torch.int32, device=dev)
pf = torch.empty(B * T + B, dtype=torch.float32, device=dev)
s = torch.cuda.current_stream().cuda_stream
def run():
    _lib.check(L.ocrpp_ctc_greedy(x.data_ptr(), 0, T, B, C, x.stride(0), x.stride(1), idx.data_ptr(), pf.data_ptr(),
                                  idx.data_ptr() + 4 * B * T, pf.data_ptr() + 4 * B * T, None, s))
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
gb = T * B * C * 4 / 1e9
print(f"B={B} ms={ms:.3f} GB/s={gb/ms*1e3:.1f} lines/s={B/ms*1e3:.0f} frac_of_6546={gb/ms*1e3/6546.2:.3f}")
