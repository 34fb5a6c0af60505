"""GPU: the UNMODIFIED reference (zjykzj/YOLOv4, installed into baseline/_ref by tools/install_reference.sh; git-ignored, it
travels to the GPU box with the snapshot) with and without yolov4_b200.patch_reference(), on the same GPU and the same inputs.

  eval   yolo.model.yolov4.YOLOv4 (yolov4.py:270-324), random init with non-degenerate BN: decoded [B,M,85] within 1e-5 relative
         (north_star's tolerance); postprocess (utils.py:92-223) on the SAME decoded tensor: identical lists, bit for bit
  train  the reference's criterion (yololoss.py:373-443) on the patched model: loss within 1e-5, gradient within 2e-5, both with
         the reference's own forward on top of the B200 YOLOLayer / build_target (its in-place mask multiplies must not break
         autograd) and with the rebound forward

Nothing here reads /root/reference.  Skipped when baseline/_ref is absent.
"""
import os
import sys

import numpy as np
import pytest
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "yolo")), reason="baseline/_ref (the installed reference) is absent")]

IMG = 128           # grids 16 / 8 / 4: the reference's Python postprocess stays fast


@pytest.fixture(scope="module")
def ref():
    import yaml
    sys.path.insert(0, REF)
    import yolo.model.yolov4 as m4
    import yolo.model.yololayer as ml
    import yolo.model.yololoss as mlo
    import yolo.util.utils as mu
    saved = dict(layer=ml.YOLOLayer, layer4=m4.YOLOLayer, post=mu.postprocess, bt=mlo.YOLOLoss.build_target, fwd=mlo.YOLOLoss.forward)

    def restore():
        ml.YOLOLayer = saved["layer"]
        m4.YOLOLayer = saved["layer4"]
        mu.postprocess = saved["post"]
        mlo.YOLOLoss.build_target = saved["bt"]
        mlo.YOLOLoss.forward = saved["fwd"]

    cfg = yaml.safe_load(open(os.path.join(REF, "config", "yolov4_default.cfg")))
    yield dict(m4=m4, ml=ml, mlo=mlo, mu=mu, cfg=cfg, restore=restore, saved=saved)
    restore()
    sys.path.remove(REF)


def _model(m4, cfg, dev):
    torch.manual_seed(0)
    model = m4.YOLOv4(cfg["MODEL"], device=dev).to(dev)
    # the reference initialises BN weights ~ N(0, 0.01) (yolov4.py:292): every head logit is ~0 after 100 layers.
    # Weight 1 and statistics taken from one batch give O(1) logits on both sides of every threshold.
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            nn.init.constant_(m.weight, 1.0)
            m.momentum = 1.0
    return model


def _calibrate(model, x):
    model.train()
    with torch.no_grad():
        model(x)
    return model


def test_eval_model_and_postprocess_patched_vs_unpatched(ref):
    import yolov4_b200 as yb
    dev = torch.device("cuda")
    m4, cfg = ref["m4"], ref["cfg"]
    ref["restore"]()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, IMG, IMG, generator=g).to(dev)
    model_ref = _calibrate(_model(m4, cfg, dev), x).eval()
    assert type(model_ref.head.yolo1[2]).__module__ == "yolo.model.yololayer"
    with torch.no_grad():
        out_ref = model_ref(x)
    # the reference's own postprocess, unpatched, on its own decoded tensor (it overwrites its input: clone)
    conf, nmst = 0.3, 0.45
    want = ref["saved"]["post"](out_ref.clone(), 80, conf, nmst)

    done = yb.patch_reference()
    assert "yolo.model.yolov4.YOLOLayer" in done and "yolo.util.utils.postprocess" in done
    try:
        model_b = _model(m4, cfg, dev)
        assert type(model_b.head.yolo1[2]).__module__ == "yolov4_b200.yololayer"
        model_b.load_state_dict(model_ref.state_dict(), strict=True)          # val.py:82-83: checkpoints load unchanged
        model_b.eval()
        with torch.no_grad():
            out_b = model_b(x)
        assert out_b.shape == out_ref.shape == (2, 3 * (16 * 16 + 8 * 8 + 4 * 4), 85)
        a, b = out_ref.cpu().numpy(), out_b.cpu().numpy()
        assert np.isfinite(a).all()
        rel = np.abs(a - b) / np.maximum(np.abs(a), 1e-30)
        assert rel.max() <= 1e-5, rel.max()                                   # north_star: decoded values within 1e-5 relative
        got = ref["mu"].postprocess(out_ref, 80, conf, nmst)                  # the patched symbol, same decoded input
    finally:
        ref["restore"]()
    n_rows = 0
    for gi, wi in zip(got, want):
        assert (gi is None) == (wi is None)
        if wi is not None:
            assert gi.shape == wi.shape
            assert np.array_equal(gi.cpu().numpy().view(np.uint32), wi.cpu().numpy().view(np.uint32))
            n_rows += wi.shape[0]
    assert n_rows > 50, "the test input must produce detections"


@pytest.mark.parametrize("loss_forward", [False, True])
def test_train_loss_and_gradient_patched_vs_unpatched(ref, loss_forward):
    import yolov4_b200 as yb
    from yolov4_b200.synth import synth_labels
    dev = torch.device("cuda")
    m4, mlo, cfg = ref["m4"], ref["mlo"], ref["cfg"]
    ref["restore"]()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 3, IMG, IMG, generator=g).to(dev)
    labels = synth_labels(2, IMG, n_valid=12, seed=3, device="cpu").double()          # as the loader collates them
    targets = {"padded_labels": labels}

    def loss_and_grads(model):
        model.train()
        model.zero_grad()
        crit = mlo.YOLOLoss(cfg["MODEL"], ignore_thresh=float(cfg["CRITERION"]["IGNORE_THRESH"]), device=dev).to(dev)
        loss = crit(model(x), targets)
        loss.backward()
        gs = [p.grad.detach().clone() for p in list(model.head.yolo1[1].parameters()) + list(model.head.yolo3[1].parameters())]
        return float(loss), gs

    model_ref = _calibrate(_model(m4, cfg, dev), x)
    loss_ref, g_ref = loss_and_grads(model_ref)
    assert np.isfinite(loss_ref) and loss_ref > 0

    done = yb.patch_reference(loss_forward=loss_forward)
    assert ("yolo.model.yololoss.YOLOLoss.forward" in done) == loss_forward
    try:
        model_b = _model(m4, cfg, dev)
        model_b.load_state_dict(model_ref.state_dict(), strict=True)
        for mr, mb in zip(model_ref.modules(), model_b.modules()):
            if isinstance(mr, nn.BatchNorm2d):
                mb.momentum = mr.momentum
        loss_b, g_b = loss_and_grads(model_b)      # the reference's forward multiplies dict['output'] in place: must back-propagate
    finally:
        ref["restore"]()
    assert abs(loss_b - loss_ref) <= 1e-5 * abs(loss_ref), (loss_b, loss_ref)
    for a, b in zip(g_ref, g_b):
        scale = float(a.abs().max())
        assert scale > 0
        assert float((a - b).abs().max()) <= 2e-5 * scale, (float((a - b).abs().max()), scale)
