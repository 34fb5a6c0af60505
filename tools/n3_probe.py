"""N3: what a fused head-convolution + filter could gain.  Times the reference's three head convolutions
(yolov4.py:237,243,249: 3x3 256->255 @76x76, 1x1 512->255 @38x38, 1x1 1024->255 @19x19, fp32 NCHW, B=64) with cuDNN in exact fp32
and with TF32 tensor cores, the error TF32 puts on the logits / decoded values, and the fused decode+filter+NMS step behind them."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yolov4_b200 as yb
B = 64
torch.manual_seed(0)
dev = "cuda"
specs = [(256, 3, 76), (512, 1, 38), (1024, 1, 19)]
convs, xs = [], []
for cin, k, F in specs:
    c = torch.nn.Conv2d(cin, 255, k, 1, k // 2, bias=True).to(dev)
    torch.nn.init.normal_(c.weight, 0, (2.0 / (cin * k * k)) ** 0.5)
    convs.append(c)
    xs.append(torch.randn(B, cin, F, F, device=dev))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ev[0].record()
    for _ in range(n): fn()
    ev[1].record(); torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n * 1e3
out = {}
with torch.no_grad():
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        ts = [timeit(lambda c=c, x=x: c(x)) for c, x in zip(convs, xs)]
        out[tf32] = [c(x) for c, x in zip(convs, xs)]
        flops = [2 * 255 * cin * k * k * B * F * F for cin, k, F in specs]
        print("cuDNN %s: %s us  (%s TFLOP/s)" % ("TF32" if tf32 else "fp32", ", ".join("%.0f" % t for t in ts),
                                                 ", ".join("%.0f" % (f / t / 1e6) for f, t in zip(flops, ts))))
    for l in range(3):
        a, b = out[False][l], out[True][l]
        print("scale %d: max |logit_tf32 - logit_fp32| = %.3e (logits std %.2f)" % (l, float((a - b).abs().max()), float(a.std())))
    cfg = {"ANCHORS": yb.ANCHORS_PX, "ANCHOR_MASK": yb.ANCHOR_MASK, "N_CLASSES": 80}
    d32 = yb.decode_dense_cat(out[False], cfg); dtf = yb.decode_dense_cat(out[True], cfg)
    rel = ((d32 - dtf).abs() / d32.abs().clamp_min(1e-30))
    print("decoded values: max relative difference TF32 vs fp32 = %.3e; fraction of elements beyond 1e-5 relative: %.4f"
          % (float(rel.max()), float((rel > 1e-5).float().mean())))
hp = yb.HeadPostprocessor(B, [76, 38, 19], 80, 0.5, 0.4).capture(out[False])
print("fused decode+filter+NMS step behind the convolutions (these i.i.d. logits, conf 0.5): %.0f us" % timeit(hp.replay, 20))
