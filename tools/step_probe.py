import os, sys, torch
sys.path.insert(0, "/root/repo")
import yolov4_b200 as yb
from yolov4_b200.synth import synth_head_outputs
B = 64
raws = synth_head_outputs(B, 608, 80, seed=0, device="cuda")
hp = yb.HeadPostprocessor(B, [76, 38, 19], 80, 1e-4, 0.4, cap_seg=int(os.environ.get("CAP", "1024"))).capture(raws)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(20): hp.replay()
torch.cuda.synchronize(); ev0.record()
N = int(os.environ.get('N', '300'))
for _ in range(N): hp.replay()
ev1.record(); torch.cuda.synchronize()
print(os.environ.get("TAG", ""), "%.2f us/step" % (ev0.elapsed_time(ev1) * 1e3 / N))
