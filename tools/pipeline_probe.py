"""Does replaying two independent postprocessors on two streams (double-buffered batches) raise throughput?"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yolov4_b200 as yb
from yolov4_b200.synth import synth_head_outputs
B = 64
raws = [synth_head_outputs(B, 608, 80, seed=s, device="cuda") for s in range(2)]
hps = [yb.HeadPostprocessor(B, [76, 38, 19], 80, 1e-4, 0.4).capture(r) for r in raws]
streams = [torch.cuda.Stream() for _ in range(2)]
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def run(n, depth):
    torch.cuda.synchronize()
    ev0.record()
    if depth == 1:
        for i in range(n):
            hps[0].replay()
    else:
        cur = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(cur)
        for i in range(n):
            with torch.cuda.stream(streams[i & 1]):
                hps[i & 1].replay()
        for s in streams:
            cur.wait_stream(s)
    ev1.record(); torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) * 1e3 / n
for d in (1, 2):
    run(20, d)
    print("depth %d: %.1f us/step  %.0f img/s" % (d, run(200, d), B / run(200, d) * 1e6))
