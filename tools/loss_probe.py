"""Forward time of the fused loss (three scales, B=64 @608, 50 GT/image) as a CUDA-graph replay; YL_PDL=0/1 A/B."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yolov4_b200 as yb
from yolov4_b200.synth import synth_head_outputs, synth_labels
from yolov4_b200.yololoss import fused_yolo_loss_components
CFG = {"ANCHORS": yb.ANCHORS_PX, "ANCHOR_MASK": yb.ANCHOR_MASK, "N_CLASSES": 80}
raws = synth_head_outputs(64, 608, 80, seed=0, device="cuda")
labels = synth_labels(64, 608, n_valid=50, seed=2, device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): ref = fused_yolo_loss_components(raws, labels, CFG, 0.7).clone()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = fused_yolo_loss_components(raws, labels, CFG, 0.7)
for _ in range(10): g.replay()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(200): g.replay()
ev1.record(); torch.cuda.synchronize()
same = torch.allclose(out, ref, rtol=1e-9, atol=0)
print("%s loss forward %.1f us (graph replay), components %s, matches eager: %s" % (os.environ.get("TAG", ""), ev0.elapsed_time(ev1) * 5, [round(v, 3) for v in out.tolist()], same))
