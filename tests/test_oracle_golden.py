"""Pin the CPU oracle (oracle/yolo_head_oracle.c) against outputs of the reference itself.

The goldens in tests/golden/*.npz were produced by tests/golden/make_golden.py, which imports the
unmodified reference.  Tolerances: decode within 1e-5 relative (BASELINE.json north_star); every index,
count, mask and IoU-decision result bit-exact.
"""
import os

import numpy as np
import pytest

from oracle import oracle as orc

REL = 1e-5   # north_star: decoded boxes and scores within 1e-5 relative in fp32


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _split(counts, rows):
    out, o = [], 0
    for n in counts:
        out.append(rows[o:o + n] if n else None)
        o += n
    return out


def assert_rel(a, b, rel=REL, what=""):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, what
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin), what
    assert np.array_equal(a[~fin], b[~fin], equal_nan=True), what
    err = np.abs(a[fin] - b[fin])
    tol = rel * np.abs(b[fin]) + 1e-37
    assert (err <= tol).all(), f"{what}: max rel err {np.max(err / (np.abs(b[fin]) + 1e-37))}"


@pytest.mark.parametrize("tag", ["c80", "c4"])
def test_decode_eval_matches_reference(golden_dir, tag):
    g = _load(golden_dir, "decode.npz")
    C = int(g[f"{tag}_C"])
    raws = [g[f"{tag}_raw{l}"] for l in range(3)]
    out = orc.decode_eval_cat(raws, C)
    ref = g[f"{tag}_eval"]
    # sigmoid outputs below fp32 normal range are flushed differently by SLEEF / the spec math: compare those absolutely
    tiny = np.abs(ref) < 1e-30
    assert np.abs(out[tiny]).max(initial=0.0) < 1e-30
    assert_rel(np.where(tiny, 0, out), np.where(tiny, 0, ref), what=f"decode eval {tag}")


@pytest.mark.parametrize("tag", ["c80", "c4"])
def test_decode_train_matches_reference(golden_dir, tag):
    g = _load(golden_dir, "decode.npz")
    C = int(g[f"{tag}_C"])
    for l in range(3):
        o, p = orc.decode_train(g[f"{tag}_raw{l}"], l, C)
        ro, rp = g[f"{tag}_train_output{l}"], g[f"{tag}_train_pred{l}"]
        tiny = np.abs(ro) < 1e-30
        assert_rel(np.where(tiny, 0, o), np.where(tiny, 0, ro), what=f"train output {tag} l{l}")
        assert_rel(p, rp, what=f"train pred {tag} l{l}")
        # wh channels of `output` stay raw logits (yololayer.py:105 skips 2:4)
        assert np.array_equal(o[..., 2:4], ro[..., 2:4])


def test_nms_cases_bit_exact(golden_dir):
    g = _load(golden_dir, "nms.npz")
    for k in g["names"]:
        keep = orc.nms(g[f"{k}_bbox"], float(g[f"{k}_thr"]), g[f"{k}_score"])
        assert keep.dtype == np.int32
        assert keep.tolist() == g[f"{k}_keep"].tolist(), k


def test_postprocess_bit_exact_on_reference_decoded_input(golden_dir):
    g = _load(golden_dir, "postprocess.npz")
    pred = g["pred"]
    i = 0
    while f"s{i}_conf" in g:
        conf, nmst = float(g[f"s{i}_conf"]), float(g[f"s{i}_nms"])
        out = orc.postprocess(pred, 80, conf, nmst)
        ref = _split(g[f"s{i}_counts"], g[f"s{i}_rows"])
        for b, (o, r) in enumerate(zip(out, ref)):
            assert (o is None) == (r is None), (i, b)
            if r is not None:
                assert o.shape == r.shape, (i, b, o.shape, r.shape)
                assert np.array_equal(o.view(np.uint32), r.view(np.uint32)), (i, b)
        i += 1
    assert i == 5
    # suppression actually happened somewhere (the clustered generator puts IoUs on both sides of the threshold)
    n_cand = int(((pred[:, :, 5:] * pred[:, :, 4:5]) >= np.float32(0.2)).sum())
    assert int(g["s1_counts"].sum()) < n_cand


def test_postprocess_all_empty_and_all_ties(golden_dir):
    g = _load(golden_dir, "postprocess.npz")
    out = orc.postprocess(g["pred"], 80, float(g["empty_conf"]), 0.4)
    assert out == [None, None]
    out = orc.postprocess(g["ties_pred"], 3, 0.2, 0.5)
    ref = _split(g["ties_counts"], g["ties_rows"])
    assert np.array_equal(out[0].view(np.uint32), ref[0].view(np.uint32))


def test_postprocess_threads_agree(golden_dir):
    g = _load(golden_dir, "postprocess.npz")
    a = orc.postprocess(g["pred"], 80, 0.005, 0.4, nthreads=1)
    b = orc.postprocess(g["pred"], 80, 0.005, 0.4, nthreads=4)
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=True)


def test_bboxes_iou_bit_exact(golden_dir):
    g = _load(golden_dir, "iou.npz")
    assert np.array_equal(orc.bboxes_iou(g["a_xywh"], g["b_xywh"], xyxy=False), g["iou_xywh"])
    assert np.array_equal(orc.bboxes_iou(g["a_xyxy"], g["b_xyxy"], xyxy=True), g["iou_xyxy"])


@pytest.mark.parametrize("layer", [0, 1, 2])
def test_build_target_matches_reference(golden_dir, layer):
    g = _load(golden_dir, "build_target.npz")
    C = int(g["C"])
    pred = g[f"pred{layer}"]
    target, obj_mask, tgt_mask, tgt_scale = orc.build_target(pred, g["labels"], layer, C, 0.7)
    assert np.array_equal(obj_mask, g[f"obj_mask{layer}"])
    assert np.array_equal(tgt_mask, g[f"tgt_mask{layer}"])
    assert np.array_equal(tgt_scale, g[f"tgt_scale{layer}"], equal_nan=True)   # sqrt(2 - wh/F^2) is NaN for boxes larger than ~1.4 grids
    rt = g[f"target{layer}"]
    # everything except the two log() channels is bit-exact; log within a few ulp (spec math vs SLEEF)
    sel = np.ones(rt.shape[-1], bool)
    sel[2:4] = False
    assert np.array_equal(target[..., sel], rt[..., sel])
    np.testing.assert_allclose(target[..., 2:4], rt[..., 2:4], rtol=1e-5, atol=1e-6)
    assert (tgt_mask[..., 0].sum() > 0)


def test_build_target_strided_pred_view(golden_dir):
    """The reference hands build_target a non-contiguous pred view; strides must be honoured."""
    g = _load(golden_dir, "build_target.npz")
    pred = g["pred1"]
    planar = np.ascontiguousarray(pred.transpose(0, 1, 4, 2, 3))      # [B,3,4,F,F] storage
    view = planar.transpose(0, 1, 3, 4, 2)                            # [B,3,F,F,4] strided
    assert not view.flags.c_contiguous
    a = orc.build_target(view, g["labels"], 1, int(g["C"]), 0.7)
    b = orc.build_target(pred, g["labels"], 1, int(g["C"]), 0.7)
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=True)


def test_spec_math_accuracy():
    x = np.concatenate([np.linspace(-87, 88, 20001), np.random.RandomState(0).randn(20000) * 4]).astype(np.float32)
    e = orc.expf(x).astype(np.float64)
    ref = np.exp(x.astype(np.float64))
    assert (np.abs(e - ref) / ref).max() < 2.5e-7          # < ~2 ulp
    xs = np.exp(np.random.RandomState(1).randn(20000) * 5).astype(np.float32)
    l = orc.logf(xs).astype(np.float64)
    assert np.abs(l - np.log(xs.astype(np.float64))).max() < 2e-6
    assert np.isnan(orc.expf(np.float32("nan")))
    assert orc.expf(np.float32(100.0)) == np.inf and orc.expf(np.float32(-200.0)) == 0.0
    assert orc.sigmoidf(np.float32(0.0)) == np.float32(0.5)


def test_loss_oracle_vs_reference_forward_and_autograd(golden_dir):
    """N2: the oracle's restatement of YOLOLoss.forward (yololoss.py:390-432) against the reference's loss value and the
    gradient its autograd produces with respect to the raw head tensors."""
    g = _load(golden_dir, "loss.npz")
    C = int(g["C"])
    total = 0.0
    for l in range(3):
        losses, grad = orc.yolo_loss_layer(g[f"raw{l}"], g["labels"], l, C, 0.7)
        assert abs(losses.sum() - float(g[f"loss{l}"])) <= 1e-5 * abs(float(g[f"loss{l}"])), (l, losses, float(g[f"loss{l}"]))
        ref = g[f"grad{l}"].astype(np.float64)
        assert np.array_equal(grad != 0, ref != 0), "layer %d: gradient support differs" % l
        assert np.allclose(grad, ref, rtol=2e-5, atol=1e-7), (l, np.abs(grad - ref).max())
        total += losses.sum()
    assert abs(total - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))


# ------------------------------------------------------------------------------------------------ BASELINE-size pins
def _fullsize(golden_dir):
    import sys
    if golden_dir not in sys.path:
        sys.path.insert(0, golden_dir)
    import fullsize_inputs as fi
    return fi, _load(golden_dir, "fullsize.npz")


def test_fullsize_postprocess_608_matches_reference(golden_dir):
    """BASELINE configs 2 and 3 at 608x608 (val 1e-4 / 0.4 and detect 0.2 / 0.5): the oracle on its own decoded tensor against
    the reference's postprocess on the same tensor (kept pairs + SHA-256 of the reference's rows; rows for the detect setting)."""
    import hashlib
    fi, g = _fullsize(golden_dir)
    raws = [r.numpy() for r in fi.raws_cpu()]
    pred = orc.decode_eval_cat(raws, fi.C)
    for tag, conf, nmst in fi.SETTINGS:
        got = orc.postprocess(pred, fi.C, conf, nmst, nthreads=4)
        counts = g[f"{tag}_counts"]
        assert [0 if o is None else len(o) for o in got] == counts.tolist()
        rows = np.concatenate([o for o in got if o is not None], 0)
        if tag == "det":
            assert np.array_equal(rows.view(np.uint32), g["det_rows"].view(np.uint32))
        else:
            want = np.concatenate(fi.rows_from_kept(pred, counts, g["val_kept_idx"], g["val_kept_cls"]), 0)
            assert hashlib.sha256(np.ascontiguousarray(want).tobytes()).hexdigest() == str(g["val_sha256"])
            assert np.array_equal(rows.view(np.uint32), want.view(np.uint32))
    # raw -> detections: the reference's own decode (ATen sigmoid / exp) leads to the same kept pairs up to borderline ones
    fused = orc.detect(raws, fi.C, fi.SETTINGS[0][1], fi.SETTINGS[0][2], nthreads=4)
    _check_refdec(fi, g, pred, fused)


def _check_refdec(fi, g, pred, got_rows):
    """Kept (box, class) pairs of a raw -> detections run against those of the reference decode + postprocess.  The two decodes
    differ by <= 2.1e-7 relative, so a pair may flip only when its score is within 1e-5 relative of conf or an IoU within 1e-5 of the
    NMS threshold; the flips are counted and bounded, not ignored."""
    co = fi.corners_obj(pred)
    o = 0
    flips = 0
    for b, n in enumerate(g["refdec_counts"]):
        want = set(zip(g["refdec_kept_idx"][o:o + n].tolist(), g["refdec_kept_cls"][o:o + n].tolist()))
        o += n
        lut = {k.tobytes(): i for i, k in enumerate(co[b].view(np.uint32))}
        r = got_rows[b]
        have = set((lut[x[:5].view(np.uint32).tobytes()], int(x[6])) for x in np.ascontiguousarray(r))
        flips += len(want ^ have)
    assert float(g["refdec_max_rel"]) <= 1e-5
    assert flips <= 4, "%d kept pairs differ between the reference-decoded and the spec-decoded run" % flips
    return flips


@pytest.mark.parametrize("layer", [0, 1, 2])
def test_fullsize_build_target_608_matches_reference(golden_dir, layer):
    """BASELINE config 4 at 608x608, 50 GT / image (B=2 of it): the reference's four dense tensors, stored sparse."""
    fi, g = _fullsize(golden_dir)
    raws = fi.raws_cpu()
    labels = fi.labels_cpu().numpy()
    _, p = orc.decode_train(raws[layer].numpy(), layer, fi.C)
    p = fi.plant_pred(np.ascontiguousarray(p).copy(), labels, layer)
    got = orc.build_target(p, labels, layer, fi.C, 0.7)
    _check_fullsize_bt(fi, g, layer, got)


def _check_fullsize_bt(fi, g, layer, got):
    for name, t, bg in zip(("target", "obj_mask", "tgt_mask", "tgt_scale"), got, (0.0, 1.0, 0.0, 0.0)):
        want = fi.sparse_unpack(tuple(g[f"bt{layer}_{name}_shape"]), bg, g[f"bt{layer}_{name}_idx"], g[f"bt{layer}_{name}_val"])
        t = np.asarray(t)
        assert t.shape == want.shape
        if name == "target":
            sel = np.ones(t.shape[-1], bool)
            sel[2:4] = False                        # log() targets: 1e-5 relative (north_star), everything else bit-exact
            assert np.array_equal(t[..., sel], want[..., sel])
            np.testing.assert_allclose(t[..., 2:4], want[..., 2:4], rtol=1e-5, atol=1e-6)
        else:
            assert np.array_equal(t, want, equal_nan=True), name
    assert (np.asarray(got[1]) == 0).sum() > 0 and np.asarray(got[2]).sum() > 0
