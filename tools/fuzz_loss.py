"""Randomised parity run of the fused loss (value and gradient) and of build_target against the CPU oracle.
usage: python tools/fuzz_loss.py [n_cases] [seed]"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yolov4_b200 as yb
from yolov4_b200.synth import synth_head_outputs
from oracle import oracle as orc

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
for it in range(n_cases):
    img = int(rng.choice([64, 96, 160, 224, 320]))
    C = int(rng.choice([1, 4, 20, 80]))
    B = int(rng.randint(1, 5))
    K = int(rng.choice([8, 30, 60]))
    ign = float(rng.choice([0.3, 0.5, 0.7, 0.9]))
    cfg = {"ANCHORS": yb.ANCHORS_PX, "ANCHOR_MASK": yb.ANCHOR_MASK, "N_CLASSES": C}
    labels = np.zeros((B, K, 5), np.float32)
    for b in range(B):
        n = int(rng.randint(0, K + 1)) if rng.rand() < 0.8 else 0
        wh = np.exp(rng.uniform(np.log(2.0), np.log(img * 0.9), (n, 2)))
        xy = rng.uniform(0.5, img - 0.5, (n, 2))
        labels[b, :n, 0:2], labels[b, :n, 2:4], labels[b, :n, 4] = xy, wh, rng.randint(0, C, n)
        if n >= 4 and rng.rand() < 0.5:                      # same-cell collisions, duplicates, another class on the same box
            labels[b, 1] = labels[b, 0]; labels[b, 1, 4] = (labels[b, 0, 4] + 1) % C
            labels[b, 2, :2] = labels[b, 0, :2] + 0.3; labels[b, 2, 2:4] = labels[b, 0, 2:4] * 1.03
            labels[b, 3] = labels[b, 0]
    raws_np = [r.numpy() for r in synth_head_outputs(B, img, C, seed=int(rng.randint(1 << 30)), fg_prob=0.05)]
    nch = 5 + C
    for l in range(3):                                       # plant some predictions on ground truths (ignore-mask zeros)
        s = 8 << l; F = img // s
        for b in range(B):
            for t in range(K):
                if labels[b, t].sum() > 0 and rng.rand() < 0.4:
                    gx, gy, gw, gh = labels[b, t, :4] / s
                    i, j = min(int(gx), F - 1), min(int(gy), F - 1); a = int(rng.randint(0, 3))
                    aw, ah = yb.ANCHORS_PX[yb.ANCHOR_MASK[l][a]][0] / s, yb.ANCHORS_PX[yb.ANCHOR_MASK[l][a]][1] / s
                    fx, fy = min(max(gx - i, 0.02), 0.98), min(max(gy - j, 0.02), 0.98)
                    raws_np[l][b, a * nch + 0, j, i] = np.log(fx / (1 - fx)); raws_np[l][b, a * nch + 1, j, i] = np.log(fy / (1 - fy))
                    raws_np[l][b, a * nch + 2, j, i] = np.log(gw / aw) + 0.05 * rng.randn(); raws_np[l][b, a * nch + 3, j, i] = np.log(gh / ah)
    raws = [torch.from_numpy(r).cuda().requires_grad_(True) for r in raws_np]
    lab = torch.from_numpy(labels).cuda()
    loss = yb.fused_yolo_loss(raws, lab, cfg, ign)
    loss.backward()
    want, ok = 0.0, True
    for l in range(3):
        ls, gr = orc.yolo_loss_layer(raws_np[l], labels, l, C, ign)
        want += ls.sum()
        got = raws[l].grad.cpu().numpy().astype(np.float64)
        ok &= np.array_equal(got != 0, gr != 0) and np.allclose(got, gr, rtol=2e-5, atol=1e-7)
        # build_target itself, bit-exact
        d = yb.YOLOLayer(cfg, l, device="cuda").train()(raws[l].detach())
        t = yb.YOLOLoss(cfg, ign, device="cuda").build_target(d["output"], d["pred"], l, lab)
        w = orc.build_target(d["pred"].cpu().numpy(), labels, l, C, ign)
        ok &= all(np.array_equal(a.cpu().numpy(), b_, equal_nan=True) for a, b_ in zip(t, w))
    ok &= abs(loss.item() - want) <= 1e-5 * abs(want) + 1e-6
    if not ok:
        bad += 1
        print("MISMATCH", dict(img=img, C=C, B=B, K=K, ign=ign), loss.item(), want, flush=True)
    elif it % 10 == 0:
        print("case %d ok (img %d C %d B %d K %d ign %g loss %.4f)" % (it, img, C, B, K, ign, want), flush=True)
print("fuzz_loss: %d cases, %d mismatches" % (n_cases, bad))
sys.exit(1 if bad else 0)
