"""Generate golden vectors by running the UNMODIFIED reference (zjykzj/YOLOv4) in the build container.

    PYTHONPATH=/root/reference PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference has no tests / fixtures of its own for this path ("parity unpinned"), so the pin
for the oracle (oracle/yolo_head_oracle.c) and for the CUDA path is the reference's own output on seeded
synthetic inputs, committed next to this script as .npz files.  /root/reference does not exist on the GPU
box; nothing but this script reads it.

Tie handling: the reference sorts with NumPy's unstable argsort (utils.py:58).  Every postprocess case is run
twice: unpatched, and with `score.argsort()` replaced *at import time, in memory* by
`score.argsort(kind='stable')`.  Tie-free cases must agree (asserted here) and are stored once; cases with
score ties store the stable-order result and carry `ties > 0`.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

REF = os.environ.get("YOLO_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.dont_write_bytecode = True

from yolo.model.yololayer import YOLOLayer          # noqa: E402
from yolo.model.yololoss import YOLOLoss, bboxes_iou  # noqa: E402
from yolo.util import utils as ref_utils            # noqa: E402

from yolov4_b200.synth import synth_head_outputs    # noqa: E402

ANCHORS = [[12, 16], [19, 36], [40, 28], [36, 75], [76, 55], [72, 146], [142, 110], [192, 243], [459, 401]]
MASK = [[0, 1, 2], [3, 4, 5], [6, 7, 8]]


def cfg(C):
    return {"ANCHORS": ANCHORS, "ANCHOR_MASK": MASK, "N_CLASSES": C}


def stable_utils():
    """The reference's utils module with only the argsort call made stable (in memory)."""
    path = os.path.join(REF, "yolo", "util", "utils.py")
    src = open(path, encoding="utf-8").read()
    assert src.count("score.argsort()[::-1]") == 1
    src = src.replace("score.argsort()[::-1]", "score.argsort(kind='stable')[::-1]")
    spec = importlib.util.spec_from_loader("ref_utils_stable", loader=None)
    mod = importlib.util.module_from_spec(spec)
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def ref_decode_eval(raws, C):
    outs = []
    for l, r in enumerate(raws):
        m = YOLOLayer(cfg(C), l, device="cpu").eval()
        with torch.no_grad():
            outs.append(m(r.clone()))
    return torch.cat(outs, 1)


def ref_decode_train(raw, l, C):
    m = YOLOLayer(cfg(C), l, device="cpu").train()
    with torch.no_grad():
        d = m(raw.clone())
    return d["output"], d["pred"]


def pack_list(lst):
    counts = np.array([0 if o is None else o.shape[0] for o in lst], dtype=np.int32)
    rows = [o.numpy() if isinstance(o, torch.Tensor) else o for o in lst if o is not None]
    rows = np.concatenate(rows, 0).astype(np.float32) if rows else np.zeros((0, 7), np.float32)
    return counts, rows


def count_ties(decoded, conf):
    """Number of (image,class) segments that contain at least two candidates with equal fp32 score."""
    ties = 0
    d = decoded.numpy()
    for b in range(d.shape[0]):
        obj = d[b, :, 4:5]
        sc = (d[b, :, 5:] * obj).astype(np.float32)
        keep = sc >= np.float32(conf)
        for c in range(sc.shape[1]):
            s = sc[keep[:, c], c]
            if s.size and np.unique(s).size != s.size:
                ties += 1
    return ties


def gen_decode():
    out = {}
    for tag, img, C, seed in (("c80", 64, 80, 1), ("c4", 96, 4, 2)):
        raws = synth_head_outputs(2, img, C, seed=seed, fg_prob=0.05)
        # widen the dynamic range so exp/sigmoid tails are exercised
        raws[0][0, :, :2, :2] *= 6.0
        out[f"{tag}_img"] = np.int32(img)
        out[f"{tag}_C"] = np.int32(C)
        for l, r in enumerate(raws):
            out[f"{tag}_raw{l}"] = r.numpy()
            o, p = ref_decode_train(r, l, C)
            out[f"{tag}_train_output{l}"] = o.contiguous().numpy()
            out[f"{tag}_train_pred{l}"] = p.contiguous().numpy()
        out[f"{tag}_eval"] = ref_decode_eval(raws, C).numpy()
    np.savez_compressed(os.path.join(HERE, "decode.npz"), **out)
    print("decode.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim})


def gen_postprocess():
    st = stable_utils()
    out = {}
    img, C = 128, 80
    raws = synth_head_outputs(2, img, C, seed=3, fg_prob=0.04, clustered=True)
    decoded = ref_decode_eval(raws, C)
    out["pred"] = decoded.numpy()
    settings = [(1e-4, 0.4), (0.2, 0.5), (0.001, 0.4), (0.005, 0.4), (0.05, 0.1)]
    for i, (conf, nmst) in enumerate(settings):
        a = ref_utils.postprocess(decoded.clone(), C, conf, nmst)
        b = st.postprocess(decoded.clone(), C, conf, nmst)
        ties = count_ties(decoded, conf)
        ca, ra = pack_list(a)
        cb, rb = pack_list(b)
        if ties == 0:
            assert np.array_equal(ca, cb) and np.array_equal(ra, rb), "tie-free case must not depend on sort stability"
        out[f"s{i}_conf"], out[f"s{i}_nms"], out[f"s{i}_ties"] = np.float64(conf), np.float64(nmst), np.int32(ties)
        out[f"s{i}_counts"], out[f"s{i}_rows"] = cb, rb
        print(f"postprocess conf={conf} nms={nmst}: rows={cb.tolist()} ties={ties} unpatched_equal={np.array_equal(ra, rb)}")
    # all-empty batch -> [None, None]
    e = ref_utils.postprocess(decoded.clone(), C, 0.99999, 0.4)
    assert all(o is None for o in e)
    out["empty_conf"] = np.float64(0.99999)
    # degenerate all-ties case (random-init network: every logit 0 -> every score 0.25), tiny grid
    z = [torch.zeros(1, 3 * (5 + 3), f, f) for f in (4, 2, 1)]
    dz = ref_decode_eval(z, 3)
    b = st.postprocess(dz.clone(), 3, 0.2, 0.5)
    out["ties_pred"] = dz.numpy()
    out["ties_counts"], out["ties_rows"] = pack_list(b)
    print("all-ties case rows", out["ties_counts"].tolist())
    np.savez_compressed(os.path.join(HERE, "postprocess.npz"), **out)


def gen_nms():
    st = stable_utils()
    cases = {}
    rng = np.random.RandomState(7)
    # random clustered boxes
    c = rng.rand(40, 2).astype(np.float32) * 50
    wh = (rng.rand(40, 2).astype(np.float32) * 30 + 5)
    bb = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    sc = rng.rand(40).astype(np.float32)
    cases["rand"] = (bb, sc, 0.3)
    # IoU exactly == thresh must suppress (utils.py:77 uses >=): two boxes with IoU 0.5 exactly
    bb = np.array([[0, 0, 2, 2], [0, 0, 2, 1], [10, 10, 12, 12]], np.float32)
    cases["iou_eq_thr"] = (bb, np.array([0.9, 0.8, 0.7], np.float32), 0.5)
    # zero-area duplicates: 0/0 = NaN -> NaN >= thr False -> all kept
    bb = np.array([[5, 5, 5, 5], [5, 5, 5, 5], [5, 5, 5, 5]], np.float32)
    cases["zero_area"] = (bb, np.array([0.3, 0.2, 0.1], np.float32), 0.4)
    # NaN / inf coordinates
    bb = np.array([[0, 0, 10, 10], [np.nan, 0, 10, 10], [0, 0, np.inf, 10], [1, 1, 9, 9]], np.float32)
    cases["nan_inf"] = (bb, np.array([0.9, 0.8, 0.7, 0.6], np.float32), 0.4)
    # ties (stable order: equal scores -> higher index first)
    bb = np.array([[0, 0, 10, 10], [1, 1, 11, 11], [20, 20, 30, 30], [0, 0, 10, 10]], np.float32)
    cases["ties"] = (bb, np.array([0.5, 0.5, 0.5, 0.5], np.float32), 0.4)
    # empty
    cases["empty"] = (np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 0.4)
    # single
    cases["single"] = (np.array([[1, 2, 3, 4]], np.float32), np.array([0.1], np.float32), 0.4)
    out = {"names": np.array(list(cases.keys()))}
    with np.errstate(all="ignore"):
        for k, (bb, sc, thr) in cases.items():
            keep = st.nms(bb.copy(), thr, score=sc.copy())
            out[f"{k}_bbox"], out[f"{k}_score"], out[f"{k}_thr"], out[f"{k}_keep"] = bb, sc, np.float64(thr), keep
            print("nms", k, keep.tolist())
    np.savez_compressed(os.path.join(HERE, "nms.npz"), **out)


def gen_iou():
    rng = np.random.RandomState(11)
    a = (rng.rand(17, 4) * 20).astype(np.float32)
    b = (rng.rand(9, 4) * 20).astype(np.float32)
    a2 = a.copy(); a2[:, 2:] += a2[:, :2]
    b2 = b.copy(); b2[:, 2:] += b2[:, :2]
    out = {"a_xywh": a, "b_xywh": b, "a_xyxy": a2, "b_xyxy": b2,
           "iou_xywh": bboxes_iou(torch.from_numpy(a), torch.from_numpy(b), xyxy=False).numpy(),
           "iou_xyxy": bboxes_iou(torch.from_numpy(a2), torch.from_numpy(b2), xyxy=True).numpy()}
    np.savez_compressed(os.path.join(HERE, "iou.npz"), **out)


def gen_build_target():
    img, C, B, K = 96, 80, 5, 60
    raws = synth_head_outputs(B, img, C, seed=5, fg_prob=0.05)
    rng = np.random.RandomState(13)
    labels = np.zeros((B, K, 5), np.float64)

    def rand_gt(n, wmax):
        wh = rng.rand(n, 2) * wmax + 4
        xy = rng.rand(n, 2) * (img - 2) + 1
        cls = rng.randint(0, C, size=(n, 1))
        return np.concatenate([xy, wh, cls], 1)

    labels[0, :12] = rand_gt(12, 90)                    # ordinary image
    # image 1: no labels at all (n == 0 -> obj_mask stays 1)
    g = rand_gt(8, 60)                                  # image 2: collisions in one cell / same anchor
    g[1, :4] = g[0, :4]; g[1, 4] = (g[0, 4] + 1) % C    # same box, different class (classes accumulate)
    g[2, :2] = g[0, :2] + 0.25; g[2, 2:4] = g[0, 2:4] * 1.02   # same cell, slightly different xy/wh: last wins
    g[3] = g[0]                                          # exact duplicate
    labels[2, :8] = g
    g = rand_gt(9, 400)                                  # image 3: big boxes (layer-2 anchors), zero row inside
    g[4] = 0.0                                           # zero row among the first rows: nlabel counts 8, uses rows 0..7
    labels[3, :9] = g
    g = rand_gt(20, 30)                                  # image 4: small boxes only (layer-0 anchors)
    labels[4, :20] = g
    out = {"labels": labels.astype(np.float32), "img": np.int32(img), "C": np.int32(C)}
    crit = YOLOLoss(cfg(C), ignore_thresh=0.7, device="cpu")
    for l, r in enumerate(raws):
        o, p = ref_decode_train(r, l, C)
        # make some predictions overlap GTs strongly so the ignore mask has zeros: copy GT boxes into pred
        F = p.shape[2]
        s = crit.strides[l]
        p = p.clone()
        for b in range(B):
            for t in range(K):
                if labels[b, t].sum() > 0 and rng.rand() < 0.7:
                    i, j = int(labels[b, t, 0] / s), int(labels[b, t, 1] / s)
                    a = rng.randint(0, 3)
                    p[b, a, j, i, :] = torch.tensor(labels[b, t, :4] / s, dtype=torch.float32) * float(1 + 0.05 * rng.randn())
        with torch.no_grad():
            tgt, om, tm, ts = crit.build_target(o, p, l, torch.from_numpy(labels))
        out[f"pred{l}"] = p.contiguous().numpy()
        out[f"target{l}"], out[f"obj_mask{l}"] = tgt.numpy(), om.numpy()
        out[f"tgt_mask{l}"], out[f"tgt_scale{l}"] = tm.numpy(), ts.numpy()
        print(f"build_target layer {l}: assigned cells={int(tm[..., 0].sum())} ignored={(om == 0).sum().item()}")
    np.savez_compressed(os.path.join(HERE, "build_target.npz"), **out)


def gen_epilogue():
    """N1: the per-detection arithmetic of validate() (engine/build.py:146-164, importable parts: utils.yolobox2xywh) and of
    detect.parse_info (detect.py:171-179, utils.yolobox2yxyx).  engine/build.py itself needs apex + pycocotools, so the
    five lines of its loop body are driven here around the reference's own converter functions."""
    rng = np.random.RandomState(17)
    n_img = 3
    counts = [40, 0, 25]
    rows = []
    for k in counts:
        xy = rng.rand(k, 2).astype(np.float32) * 500
        wh = rng.rand(k, 2).astype(np.float32) * 100 + 1
        r = np.concatenate([xy, xy + wh, rng.rand(k, 2).astype(np.float32), rng.randint(0, 80, (k, 1)).astype(np.float32)], 1)
        rows.append(torch.from_numpy(r.astype(np.float32)))
    img_info = [[480, 640, 456, 608], [1080, 1920, 342, 608], [333, 500, 405, 608]]      # src_h, src_w, dst_h, dst_w
    image_ids = [139, 285, 632]
    class_ids = [int(i * 1.13) + 1 for i in range(80)]                                    # COCO-like sparse category ids
    coco, det = [], []
    for b in range(n_img):
        outputs = rows[b].cpu().data
        for output in outputs:                                                            # engine/build.py:146-164
            x1 = float(output[0]); y1 = float(output[1]); x2 = float(output[2]); y2 = float(output[3])
            label = class_ids[int(output[6])]
            bbox = ref_utils.yolobox2xywh((y1, x1, y2, x2), img_info[b][:4])
            score = float(output[4].data.item() * output[5].data.item())
            coco.append([image_ids[b], label] + bbox + [score])
        for x1, y1, x2, y2, conf, cls_conf, cls_pred in rows[b].numpy():                  # detect.py:171-179
            box = ref_utils.yolobox2yxyx([y1, x1, y2, x2], img_info[b][:4])
            det.append([b, class_ids[int(cls_pred)]] + [float(v) for v in box] + [float(cls_conf.item())])
    out = {"rows": np.concatenate([r.numpy() for r in rows], 0), "counts": np.array(counts, np.int32),
           "img_info": np.array(img_info, np.float64), "image_ids": np.array(image_ids, np.int64),
           "class_ids": np.array(class_ids, np.int32), "coco": np.array(coco, np.float64), "detect": np.array(det, np.float64)}
    np.savez_compressed(os.path.join(HERE, "epilogue.npz"), **out)
    print("epilogue.npz", out["coco"].shape, out["detect"].shape)


def gen_loss():
    """N2: YOLOLoss.forward (yololoss.py:373-443) on the three train-mode YOLOLayer outputs, and its gradient with respect
    to the raw head tensors by the reference's own autograd.  Some raw cells are planted so that their decoded box
    coincides with a ground truth (ignore mask zeros, yololoss.py:276-294); labels include same-cell collisions."""
    img, C, B, K = 96, 20, 4, 60
    raws = [r.clone() for r in synth_head_outputs(B, img, C, seed=21, fg_prob=0.05)]
    rng = np.random.RandomState(31)
    labels = np.zeros((B, K, 5), np.float64)

    def rand_gt(n, wmax):
        wh = rng.rand(n, 2) * wmax + 4
        xy = rng.rand(n, 2) * (img - 2) + 1
        cls = rng.randint(0, C, size=(n, 1))
        return np.concatenate([xy, wh, cls], 1)

    labels[0, :10] = rand_gt(10, 80)
    g = rand_gt(6, 50)
    g[1, :4] = g[0, :4]; g[1, 4] = (g[0, 4] + 1) % C        # same box, other class: one-hots accumulate
    g[2, :2] = g[0, :2] + 0.25; g[2, 2:4] = g[0, 2:4] * 1.02  # same cell: last writer wins on xy / wh / scale
    labels[1, :6] = g
    # image 2: no labels
    labels[3, :14] = rand_gt(14, 30)
    nch = 5 + C
    for l in range(3):                                       # plant predictions on top of ground truths
        s = 8 << l
        F = img // s
        for b in range(B):
            for t in range(K):
                if labels[b, t].sum() > 0 and rng.rand() < 0.6:
                    gx, gy, gw, gh = labels[b, t, :4] / s
                    i, j = min(int(gx), F - 1), min(int(gy), F - 1)
                    a = rng.randint(0, 3)
                    aw, ah = ANCHORS[MASK[l][a]][0] / s, ANCHORS[MASK[l][a]][1] / s
                    fx, fy = min(max(gx - i, 0.02), 0.98), min(max(gy - j, 0.02), 0.98)
                    raws[l][b, a * nch + 0, j, i] = float(np.log(fx / (1 - fx)))
                    raws[l][b, a * nch + 1, j, i] = float(np.log(fy / (1 - fy)))
                    raws[l][b, a * nch + 2, j, i] = float(np.log(gw / aw))
                    raws[l][b, a * nch + 3, j, i] = float(np.log(gh / ah))
    out = {"labels": labels.astype(np.float32), "img": np.int32(img), "C": np.int32(C)}
    crit = YOLOLoss(cfg(C), ignore_thresh=0.7, device="cpu")
    leaves = [r.clone().requires_grad_(True) for r in raws]
    # the reference layer works in place on (a view of) its input, which in the model is a conv output, not a leaf
    outs = [YOLOLayer(cfg(C), l, device="cpu").train()(leaves[l] * 1.0) for l in range(3)]
    total = crit(outs, {"padded_labels": torch.from_numpy(labels)})
    total.backward()
    out["loss"] = np.float64(total.item())
    for l in range(3):
        out[f"raw{l}"] = raws[l].numpy()
        out[f"grad{l}"] = leaves[l].grad.numpy()
        # per-layer value (fresh graph: the reference's forward overwrites `output` in place)
        leaf = raws[l].clone()
        with torch.no_grad():
            d = YOLOLayer(cfg(C), l, device="cpu").train()(leaf)
            out[f"loss{l}"] = np.float64(crit([d], {"padded_labels": torch.from_numpy(labels)}).item())
    print("loss", out["loss"], [out[f"loss{l}"] for l in range(3)])
    np.savez_compressed(os.path.join(HERE, "loss.npz"), **out)


def gen_fullsize():
    """BASELINE-size pins without big files (configs 2, 3 and 4 at 608x608): the inputs are seeded (fullsize_inputs.py), the
    decoded tensor is the oracle's (bit-equal to the CUDA decode, which the GPU test asserts), and what is stored is
      * val setting (conf 1e-4, nms 0.4): per image the kept (box row, class) pairs of the REFERENCE postprocess in output
        order + a SHA-256 of its output rows (the test rebuilds the rows from the pairs and the decoded tensor);
      * detect setting (conf 0.2, nms 0.5): the reference's output rows;
      * build_target at 50 GT/image, three layers: the reference's four dense tensors, stored sparse;
      * raw -> detections through the reference's OWN decode (ATen sigmoid / exp) + postprocess: kept pairs, to be matched by
        the fused CUDA path up to a counted set of threshold-borderline pairs."""
    import hashlib
    from oracle import oracle as orc
    sys.path.insert(0, HERE)
    import fullsize_inputs as fi
    st = stable_utils()
    raws = fi.raws_cpu()
    pred = torch.from_numpy(orc.decode_eval_cat([r.numpy() for r in raws], fi.C))
    out = {}

    def kept_pairs(decoded, lst):
        co = fi.corners_obj(decoded.numpy())
        idx_all, cls_all = [], []
        for b, o in enumerate(lst):
            if o is None:
                continue
            lut = {}
            for i, k in enumerate(co[b].view(np.uint32)):
                lut.setdefault(k.tobytes(), []).append(i)
            o = o.numpy()
            for r in o:
                cand = lut[r[:5].view(np.uint32).tobytes()]
                cls = int(r[6])
                hit = [i for i in cand if decoded[b, i, 5 + cls].numpy().view(np.uint32) == r[5:6].view(np.uint32)[0]]
                assert len(hit) == 1, "ambiguous row -> box match"
                idx_all.append(hit[0]); cls_all.append(cls)
        return np.array(idx_all, np.int32), np.array(cls_all, np.int8)

    for tag, conf, nmst in fi.SETTINGS:
        a = ref_utils.postprocess(pred.clone(), fi.C, conf, nmst)
        b = st.postprocess(pred.clone(), fi.C, conf, nmst)
        ties = count_ties(pred, conf)
        ca, ra = pack_list(a)
        cb, rb = pack_list(b)
        if ties == 0:
            assert np.array_equal(ca, cb) and np.array_equal(ra, rb)
        out[f"{tag}_ties"], out[f"{tag}_counts"] = np.int32(ties), cb
        out[f"{tag}_unpatched_equal"] = np.int32(np.array_equal(ca, cb) and np.array_equal(ra, rb))
        if tag == "det":
            out["det_rows"] = rb
        else:
            ki, kc = kept_pairs(pred, b)
            rebuilt = np.concatenate(fi.rows_from_kept(pred.numpy(), cb, ki, kc), 0)
            assert np.array_equal(rebuilt.view(np.uint32), rb.view(np.uint32)), "rows_from_kept must rebuild the reference's rows"
            out["val_kept_idx"], out["val_kept_cls"] = ki, kc
            out["val_sha256"] = np.array(hashlib.sha256(np.ascontiguousarray(rb).tobytes()).hexdigest())
        print(f"fullsize {tag}: rows {cb.tolist()} ties {ties} unpatched_equal {int(out[f'{tag}_unpatched_equal'])}")

    # raw -> detections through the reference's own decode
    dec_ref = ref_decode_eval(raws, fi.C)
    conf, nmst = fi.SETTINGS[0][1], fi.SETTINGS[0][2]
    lst = st.postprocess(dec_ref.clone(), fi.C, conf, nmst)
    cr, _ = pack_list(lst)
    ki, kc = kept_pairs(dec_ref, lst)
    out["refdec_counts"], out["refdec_kept_idx"], out["refdec_kept_cls"] = cr, ki, kc
    rel = (dec_ref - pred).abs() / pred.abs().clamp_min(1e-30)
    out["refdec_max_rel"] = np.float64(rel.max().item())
    print("fullsize raw->detections via the reference decode: rows", cr.tolist(), "max rel decode difference", float(rel.max()))

    # build_target, 50 GT / image
    labels = fi.labels_cpu()
    crit = YOLOLoss(cfg(fi.C), ignore_thresh=0.7, device="cpu")
    for l, r in enumerate(raws):
        o, p = orc.decode_train(r.numpy(), l, fi.C)
        p = fi.plant_pred(np.ascontiguousarray(p).copy(), labels.numpy(), l)
        with torch.no_grad():
            tgt, om, tm, ts = crit.build_target(torch.from_numpy(np.ascontiguousarray(o)), torch.from_numpy(p), l, labels.double())
        for name, t, bg in (("target", tgt, 0.0), ("obj_mask", om, 1.0), ("tgt_mask", tm, 0.0), ("tgt_scale", ts, 0.0)):
            idx, val = fi.sparse_pack(t.numpy(), bg)
            out[f"bt{l}_{name}_idx"], out[f"bt{l}_{name}_val"] = idx.astype(np.int32 if idx.size == 0 or idx.max() < 2**31 else np.int64), val
            out[f"bt{l}_{name}_shape"] = np.array(t.shape, np.int64)
        print(f"fullsize build_target layer {l}: assigned cells={int(tm[..., 0].sum())} ignored={(om == 0).sum().item()}")
    np.savez_compressed(os.path.join(HERE, "fullsize.npz"), **out)


if __name__ == "__main__":
    torch.manual_seed(0)
    only = sys.argv[1:]
    for name, fn in (("decode", gen_decode), ("postprocess", gen_postprocess), ("nms", gen_nms), ("iou", gen_iou),
                     ("build_target", gen_build_target), ("epilogue", gen_epilogue), ("loss", gen_loss),
                     ("fullsize", gen_fullsize)):
        if not only or name in only:
            fn()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
