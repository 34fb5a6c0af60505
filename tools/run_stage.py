"""Runs one stage of the hot path a few times (for ncu captures and A/B timing).
usage: python tools/run_stage.py {filter|nms|all} [--batch 64] [--iters 5] [--conf 1e-4]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yolov4_b200 as yb                      # noqa: E402
from yolov4_b200 import _cabi                 # noqa: E402
from yolov4_b200.synth import synth_head_outputs   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("stage", choices=["filter", "nms", "all"])
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--conf", type=float, default=1e-4)
ap.add_argument("--nms", type=float, default=0.4)
ap.add_argument("--groups", type=int, default=1)
a = ap.parse_args()
B = a.batch
raws = synth_head_outputs(B, 608, 80, seed=0, device="cuda")
hp = yb.HeadPostprocessor(B, [76, 38, 19], 80, a.conf, a.nms, n_groups=a.groups)
L = _cabi.lib()
st = torch.cuda.current_stream().cuda_stream
rp = _cabi.ptrs([r.data_ptr() for r in raws])


def filt():
    _cabi.check(L.yl_post_reset(hp.ws.ptr(), hp.ws.nbytes, B, hp.M, 80, hp.cap_seg, st))
    _cabi.check(L.yl_filter_raw(rp, hp.fs, 3, B, 80, hp.anch, hp.mask, hp.conf, hp.ws.ptr(), hp.ws.nbytes, hp.M, hp.cap_seg, 0, B, st))


def nms():
    _cabi.check(L.yl_nms(hp.ws.ptr(), hp.ws.nbytes, B, hp.M, 80, hp.cap_seg, hp.nms, hp.rows.data_ptr(), hp.cap_out,
                         hp.meta.data_ptr(), 0, B, st))


ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for stage, fn in (("filter", filt), ("nms", nms)):
    if a.stage not in (stage, "all") and not (a.stage == "nms" and stage == "filter"):
        continue
    reps = a.iters if a.stage in (stage, "all") else 1
    if stage == "nms":
        # the NMS stage compacts segments in place, so every repetition needs a fresh filter pass
        ts = []
        for _ in range(reps):
            filt()
            ev[0].record(); fn(); ev[1].record(); torch.cuda.synchronize()
            ts.append(ev[0].elapsed_time(ev[1]) * 1e3)
        print("%s: min %.1f us  median %.1f us" % (stage, min(ts), sorted(ts)[len(ts) // 2]))
    else:
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            ev[0].record(); fn(); ev[1].record(); torch.cuda.synchronize()
            ts.append(ev[0].elapsed_time(ev[1]) * 1e3)
        print("%s: min %.1f us  median %.1f us  (%.0f GB/s at min)" % (stage, min(ts), sorted(ts)[len(ts) // 2],
                                                                       B * 7732620 / min(ts) / 1e3))
