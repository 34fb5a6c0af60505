"""Launches every kernel family a few times at the BASELINE.json sizes so that an ncu launch list
(ncu --metrics gpu__time_duration.sum -k regex:^k_ ...) shows per-kernel durations.  No timing of its own."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yolov4_b200 as yb
from yolov4_b200.synth import synth_head_outputs, synth_labels
CFG = {"ANCHORS": yb.ANCHORS_PX, "ANCHOR_MASK": yb.ANCHOR_MASK, "N_CLASSES": 80}
raws = synth_head_outputs(64, 608, 80, seed=0, device="cuda")
for _ in range(3):
    yb.detect_raw(raws, 80, 1e-4, 0.4)                                     # config 2, fused
layers = [yb.YOLOLayer(CFG, l, device="cuda").eval() for l in range(3)]
for _ in range(3):
    dense = torch.cat([layers[l](raws[l]) for l in range(3)], 1)          # A2 contract-literal decode
for _ in range(3):
    yb.postprocess(dense, 80, 1e-4, 0.4)                                   # A5 contract-literal postprocess
del dense
labels = synth_labels(64, 608, n_valid=50, seed=2, device="cuda")
crit = yb.YOLOLoss(CFG, 0.7, device="cuda")
for _ in range(3):
    outs = [yb.YOLOLayer(CFG, l, device="cuda").train()(raws[l]) for l in range(3)]      # A3
    for l in range(3):
        crit.build_target(outs[l]["output"], outs[l]["pred"], l, labels)                # A6/A7
del outs
raws3 = synth_head_outputs(256, 608, 80, seed=1, device="cuda")
for _ in range(3):
    yb.detect_raw(raws3, 80, 0.2, 0.5)                                     # config 3
torch.cuda.synchronize()
raws_g = [r.clone().requires_grad_(True) for r in raws]
for _ in range(3):
    yb.fused_yolo_loss(raws_g, labels, CFG, 0.7).backward()                 # N2
torch.cuda.synchronize()
