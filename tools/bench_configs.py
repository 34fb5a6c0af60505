"""Times the other BASELINE.json configurations on one GPU (CUDA events, eager launches through the Python shim where
noted).  Not the driver's bench; results are recorded in profiles/ and DESIGN.md.
usage: python tools/bench_configs.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yolov4_b200 as yb                                  # noqa: E402
from yolov4_b200.synth import synth_head_outputs, synth_labels   # noqa: E402

CFG = {"ANCHORS": yb.ANCHORS_PX, "ANCHOR_MASK": yb.ANCHOR_MASK, "N_CLASSES": 80}
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        ev[0].record(); fn(); ev[1].record(); torch.cuda.synchronize()
        ts.append(ev[0].elapsed_time(ev[1]) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def graph_time(hp, n=100):
    for _ in range(10):
        hp.replay()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(n):
        hp.replay()
    ev[1].record(); torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) * 1e3 / n


BPI = 22743 * 85 * 4
print("config 2: fused decode+filter+NMS, B=64 @608, conf 1e-4 / nms 0.4")
raws = synth_head_outputs(64, 608, 80, seed=0, device="cuda")
hp = yb.HeadPostprocessor(64, [76, 38, 19], 80, 1e-4, 0.4).capture(raws)
us = graph_time(hp)
rows = sum(0 if r is None else len(r) for r in hp.results())
print("  %.1f us/step  %.0f img/s  %.0f GB/s algorithmic  (%d rows)" % (us, 64 / us * 1e6, 64 * BPI / us / 1e3, rows))

print("config 3: detect setting, B=256 @608, conf 0.2 / nms 0.5 (sparse survivors)")
raws3 = synth_head_outputs(256, 608, 80, seed=1, device="cuda")
hp3 = yb.HeadPostprocessor(256, [76, 38, 19], 80, 0.2, 0.5, cap_seg=256, cap_out=2048).capture(raws3)
us = graph_time(hp3, n=50)
rows = sum(0 if r is None else len(r) for r in hp3.results())
print("  %.1f us/step  %.0f img/s  %.0f GB/s algorithmic  (%d rows)" % (us, 256 / us * 1e6, 256 * BPI / us / 1e3, rows))
del raws3, hp3

print("contract-literal path, B=64 @608: YOLOLayer.forward x3 + torch.cat + postprocess(dense), conf 1e-4 / nms 0.4")
layers = [yb.YOLOLayer(CFG, l, device="cuda").eval() for l in range(3)]
us_dec = timeit(lambda: [layers[l](raws[l]) for l in range(3)])
dense = torch.cat([layers[l](raws[l]) for l in range(3)], 1)
us_cat = timeit(lambda: torch.cat([dense[:, :17328], dense[:, 17328:21660], dense[:, 21660:]], 1))
t0 = time.perf_counter(); out = yb.postprocess(dense, 80, 1e-4, 0.4); torch.cuda.synchronize(); t1 = time.perf_counter()
t0 = time.perf_counter(); out = yb.postprocess(dense, 80, 1e-4, 0.4); torch.cuda.synchronize(); t1 = time.perf_counter()
print("  decode x3 %.1f us (%.0f GB/s r+w)   torch.cat %.1f us   postprocess (wall, incl. the count sync) %.1f us" % (
    us_dec, 2 * 64 * BPI / us_dec / 1e3, us_cat, (t1 - t0) * 1e6))
del dense, out

print("config 4: YOLOLoss.build_target x3 layers, B=64 @608, 50 GT/image")
labels = synth_labels(64, 608, n_valid=50, seed=2, device="cuda")
crit = yb.YOLOLoss(CFG, 0.7, device="cuda")
outs = [yb.YOLOLayer(CFG, l, device="cuda").train()(raws[l]) for l in range(3)]
def bt():
    for l in range(3):
        crit.build_target(outs[l]["output"], outs[l]["pred"], l, labels)
us = timeit(bt)
bytes_bt = 64 * 16.01e6
print("  %.1f us/step  %.0f img/s  %.0f GB/s of dense outputs (eager, incl. allocation of 1 GB of outputs)" % (us, 64 / us * 1e6, bytes_bt / us / 1e3))
us_dt = timeit(lambda: [yb.YOLOLayer(CFG, l, device="cuda").train()(raws[l]) for l in range(3)])
print("  train-mode YOLOLayer.forward x3: %.1f us (%.0f GB/s r+w)" % (us_dt, 2 * 64 * BPI / us_dt / 1e3))

print("N2: fused YOLO loss (raw head tensors -> loss -> grad_raw), B=64 @608, 50 GT/image, all three layers")
raws_g = [r.clone().requires_grad_(True) for r in raws]
def fused_fb():
    for r in raws_g:
        r.grad = None
    yb.fused_yolo_loss(raws_g, labels, CFG, 0.7).backward()
us_f = timeit(lambda: yb.fused_yolo_loss_components(raws, labels, CFG, 0.7))
us_fb = timeit(fused_fb)
def unfused_fb():
    for r in raws_g:
        r.grad = None
    outs_ = [yb.YOLOLayer(CFG, l, device="cuda").train()(raws_g[l]) for l in range(3)]
    crit(outs_, {"padded_labels": labels.double()}).backward()
us_ufb = timeit(unfused_fb, n=5, warm=2)
print("  fused forward %.1f us; fused forward+backward %.1f us (eager, CUDA events); unfused CUDA path "
      "(YOLOLayer.train + build_target + torch loss arithmetic + autograd) %.1f us" % (us_f, us_fb, us_ufb))
