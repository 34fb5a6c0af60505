"""Runs detect_raw on a few seeded inputs and compares with the oracle, bit for bit.  Executed by test_gpu_alt_paths.py in
a fresh process per environment setting (the kernel-form switches are read once at library load)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc                                  # noqa: E402  (checker only)
import yolov4_b200 as yb                                           # noqa: E402
from yolov4_b200.synth import synth_head_outputs                   # noqa: E402


def same(got, want):
    for g, w in zip(got, want):
        if (g is None) != (w is None):
            return False
        if w is not None:
            g = g.cpu().numpy()
            if g.shape != w.shape or not np.array_equal(g.view(np.uint32), w.view(np.uint32)):
                return False
    return True


ok = True
for img, C, B, conf, nmst, kw in ((608, 80, 3, 1e-4, 0.4, {}), (416, 80, 2, 0.005, 0.4, dict(fg_prob=0.03, clustered=True)),
                                  (608, 80, 2, 0.2, 0.5, dict(fg_prob=0.02, clustered=True)), (96, 20, 4, 0.01, 0.45, dict(fg_prob=0.05))):
    raws = synth_head_outputs(B, img, C, seed=img + B, device="cuda", **kw)
    want = orc.detect([r.cpu().numpy() for r in raws], C, conf, nmst, nthreads=8)
    got = yb.detect_raw(raws, C, conf, nmst)
    good = same(got, want)
    print("img %d C %d conf %g: %s (%d rows)" % (img, C, conf, "ok" if good else "MISMATCH", sum(0 if w is None else len(w) for w in want)))
    ok = ok and good
raws = synth_head_outputs(4, 608, 80, seed=5, device="cuda")
hp = yb.HeadPostprocessor(4, [76, 38, 19], 80, 1e-4, 0.4, n_groups=2).capture(raws)
hp.replay(); torch.cuda.synchronize()
want = orc.detect([r.cpu().numpy() for r in raws], 80, 1e-4, 0.4, nthreads=8)
good = same(hp.results(), want)
print("graph replay, 2 image groups: %s" % ("ok" if good else "MISMATCH"))
sys.exit(0 if ok and good else 1)
