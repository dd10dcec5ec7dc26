import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
from pytorchocr_b200 import synth
from pytorchocr_b200.postprocess import build_post_process
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
maps = torch.from_numpy(synth.db_batch(8)).cuda().repeat((N + 7) // 8, 1, 1, 1)[:N].contiguous()
op = build_post_process({"name": "DBPostProcess", "thresh": 0.3, "box_thresh": 0.5, "unclip_ratio": 1.7, "cpp_speedup": True, "cuda_speedup": True})
sl = np.array([[736, 1280, 1.0, 1.0]] * N)
for it in range(3):
    boxes, scores, counts, status, ex = op.run_device(maps, sl, boxes_f=True)
bf = ex["boxes_f"].reshape(N, -1, 8)
clk = bf[:, -3:, :].reshape(N, 24)
names = ["A rowptr", "A runs", "B pass1", "B jump", "B merge", "B flatten", "ids", "C stats", "D tree", "E fill", "F extents", "G hull"]
sub = clk[:, 12:15]
clk = clk[:, :12]
d = np.diff(np.concatenate([np.zeros((N, 1)), clk], 1), axis=1)
print("counts", counts[:8])
for i, nm in enumerate(names):
    print("%-10s mean %8.0f cyc  max %8.0f" % (nm, d[:, i].mean(), d[:, i].max()))
print("total mean %.0f cyc = %.1f us" % (clk[:, 11].mean(), clk[:, 11].mean() / 1965))
ex = bf[:, -3:, :].reshape(N, 24)
print("F detail: fg %.0f | stairs %.0f | hole list %.0f | hole runs %.0f   (nhole unknown)" % ((ex[:,12]-ex[:,9]).mean(), (ex[:,15]-ex[:,12]).mean(), (ex[:,16]-ex[:,15]).mean(), (ex[:,10]-ex[:,16]).mean()))
print("G detail: list+loads %.0f | chains %.0f | pixel sums %.0f | triage+write %.0f" % ((ex[:,14]-ex[:,10]).mean(), (ex[:,17]-ex[:,14]).mean(), (ex[:,18]-ex[:,17]).mean(), (ex[:,11]-ex[:,18]).mean()))
print("F: fg loop %.0f, stairs+holes %.0f | G: list+loads %.0f, chains+triage %.0f" % ((sub[:,0]-clk[:,9]).mean(), (clk[:,10]-sub[:,0]).mean(), (sub[:,2]-clk[:,10]).mean(), (clk[:,11]-sub[:,2]).mean()))
