"""ctypes loader for the CPU oracle (oracle/yolo_head_oracle.c).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` leg may import this module; the product package never does.

Every wrapper takes/returns numpy arrays (fp32, C-contiguous) and mirrors one reference function:
  decode_eval      YOLOLayer.forward eval       yolo/model/yololayer.py:88-120,146-166
  decode_train     YOLOLayer.forward train      yolo/model/yololayer.py:122-145
  nms              nms                          yolo/util/utils.py:32-89
  postprocess      postprocess                  yolo/util/utils.py:92-223
  bboxes_iou       bboxes_iou                   yolo/model/yololoss.py:16-91
  build_target     YOLOLoss.build_target        yolo/model/yololoss.py:118-371
  yolo_loss_layer  YOLOLoss.forward, one layer  yolo/model/yololoss.py:390-432  (numpy on top of decode_train + build_target;
                   pinned to the reference's loss value and autograd gradient, tests/golden/loss.npz)
  detect           decode_eval x3 + cat + postprocess (the composition the parity tests and the CPU baseline use)
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

ANCHORS_PX = [[12, 16], [19, 36], [40, 28], [36, 75], [76, 55], [72, 146], [142, 110], [192, 243], [459, 401]]
ANCHOR_MASK = [[0, 1, 2], [3, 4, 5], [6, 7, 8]]
STRIDES = [8, 16, 32]

_f = ctypes.c_float
_i = ctypes.c_int
_l = ctypes.c_long
_fp = ctypes.POINTER(ctypes.c_float)
_ip = ctypes.POINTER(ctypes.c_int)
_lp = ctypes.POINTER(ctypes.c_long)


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc only, no GPU)."""
    src = os.path.join(_HERE, "yolo_head_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.orc_expf.restype = _f
        L.orc_expf.argtypes = [_f]
        L.orc_logf.restype = _f
        L.orc_logf.argtypes = [_f]
        L.orc_sigmoidf.restype = _f
        L.orc_sigmoidf.argtypes = [_f]
        L.orc_decode_eval.restype = None
        L.orc_decode_eval.argtypes = [_fp, _i, _i, _i, _fp, _f, _fp, _l, _l]
        L.orc_decode_train.restype = None
        L.orc_decode_train.argtypes = [_fp, _i, _i, _i, _fp, _fp, _fp]
        L.orc_nms.restype = _i
        L.orc_nms.argtypes = [_fp, _fp, _i, _f, _ip]
        L.orc_postprocess.restype = _l
        L.orc_postprocess.argtypes = [_fp, _i, _l, _i, _i, _f, _f, _fp, _l, _ip, _i]
        L.orc_detect.restype = _l
        L.orc_detect.argtypes = [_fp, _fp, _fp, _i, _i, _i, _i, _i, _fp, _f, _f, _fp, _l, _ip, _i]
        L.orc_bboxes_iou.restype = None
        L.orc_bboxes_iou.argtypes = [_fp, _i, _fp, _i, _i, _fp]
        L.orc_build_target.restype = _i
        L.orc_build_target.argtypes = [_fp, _lp, _fp, _i, _i, _i, _i, _i, _fp, _ip, _f, _fp, _fp, _fp, _fp]
        _lib = L
    return _lib


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_fp)


def masked_anchors_grid(layer_no, anchors=ANCHORS_PX, mask=ANCHOR_MASK):
    """yololayer.py:73-76 -- anchors / stride in Python doubles, then cast to fp32."""
    s = STRIDES[layer_no]
    return np.array([[anchors[i][0] / s, anchors[i][1] / s] for i in mask[layer_no]], dtype=np.float64).astype(np.float32)


def expf(x):
    return np.array([lib().orc_expf(float(v)) for v in np.asarray(x, dtype=np.float32).ravel()], dtype=np.float32).reshape(np.shape(x))


def logf(x):
    return np.array([lib().orc_logf(float(v)) for v in np.asarray(x, dtype=np.float32).ravel()], dtype=np.float32).reshape(np.shape(x))


def sigmoidf(x):
    return np.array([lib().orc_sigmoidf(float(v)) for v in np.asarray(x, dtype=np.float32).ravel()], dtype=np.float32).reshape(np.shape(x))


def decode_eval(raw, layer_no, n_classes=80, anchors=ANCHORS_PX, mask=ANCHOR_MASK):
    """raw [B, 3*(5+C), F, F] -> [B, 3*F*F, 5+C]"""
    raw, rp = _c(raw)
    B, _, F, _ = raw.shape
    anch, ap = _c(masked_anchors_grid(layer_no, anchors, mask).ravel())
    out = np.empty((B, 3 * F * F, 5 + n_classes), dtype=np.float32)
    lib().orc_decode_eval(rp, B, F, n_classes, ap, float(STRIDES[layer_no]), out.ctypes.data_as(_fp), 3 * F * F, 0)
    return out


def decode_eval_cat(raws, n_classes=80, anchors=ANCHORS_PX, mask=ANCHOR_MASK):
    """Three scales -> [B, sum 3F^2, 5+C] (yolov4.py:324 torch.cat order: layer 0 first)."""
    return np.concatenate([decode_eval(r, l, n_classes, anchors, mask) for l, r in enumerate(raws)], axis=1)


def decode_train(raw, layer_no, n_classes=80, anchors=ANCHORS_PX, mask=ANCHOR_MASK):
    """raw [B,3*(5+C),F,F] -> (output [B,3,F,F,5+C] permuted view, pred [B,3,F,F,4] permuted view)"""
    raw, rp = _c(raw)
    B, _, F, _ = raw.shape
    anch, ap = _c(masked_anchors_grid(layer_no, anchors, mask).ravel())
    outp = np.empty((B, 3, 5 + n_classes, F, F), dtype=np.float32)
    pred = np.empty((B, 3, 4, F, F), dtype=np.float32)
    lib().orc_decode_train(rp, B, F, n_classes, ap, outp.ctypes.data_as(_fp), pred.ctypes.data_as(_fp))
    return outp.transpose(0, 1, 3, 4, 2), pred.transpose(0, 1, 3, 4, 2)


def nms(bbox, thresh, score):
    bbox, bp = _c(np.asarray(bbox, dtype=np.float32).reshape(-1, 4))
    R = bbox.shape[0]
    score, sp = _c(np.asarray(score, dtype=np.float32).reshape(-1))
    out = np.empty((max(R, 1),), dtype=np.int32)
    K = lib().orc_nms(bp, sp, R, float(np.float32(thresh)), out.ctypes.data_as(_ip))
    return out[:K].copy()


def postprocess(prediction, num_classes, conf_thre=0.7, nms_thre=0.45, nthreads=1):
    """prediction [B,M,5+C] decoded xywh -> list of [K_i,7] arrays or None (utils.py:92-223)."""
    pred, pp = _c(prediction)
    B, M, nch = pred.shape
    counts = np.zeros((B,), dtype=np.int32)
    cap = 1 << 16
    while True:
        rows = np.empty((cap, 7), dtype=np.float32)
        total = lib().orc_postprocess(pp, B, M, nch, num_classes, float(np.float32(conf_thre)), float(np.float32(nms_thre)),
                                      rows.ctypes.data_as(_fp), cap, counts.ctypes.data_as(_ip), nthreads)
        if total <= cap:
            break
        cap = int(total)
    out, o = [], 0
    for b in range(B):
        n = int(counts[b])
        out.append(rows[o:o + n].copy() if n else None)
        o += n
    return out


def detect(raws, n_classes, conf_thre, nms_thre, anchors=ANCHORS_PX, nthreads=1):
    """Fused oracle: decode three scales + postprocess; returns list of [K_i,7] or None."""
    arrs = [_c(r) for r in raws]
    B = arrs[0][0].shape[0]
    Fs = [a[0].shape[2] for a in arrs]
    anch, ap = _c(np.array(anchors, dtype=np.float32).ravel())
    counts = np.zeros((B,), dtype=np.int32)
    cap = 1 << 16
    while True:
        rows = np.empty((cap, 7), dtype=np.float32)
        total = lib().orc_detect(arrs[0][1], arrs[1][1], arrs[2][1], B, Fs[0], Fs[1], Fs[2], n_classes, ap,
                                 float(np.float32(conf_thre)), float(np.float32(nms_thre)),
                                 rows.ctypes.data_as(_fp), cap, counts.ctypes.data_as(_ip), nthreads)
        if total <= cap:
            break
        cap = int(total)
    out, o = [], 0
    for b in range(B):
        n = int(counts[b])
        out.append(rows[o:o + n].copy() if n else None)
        o += n
    return out


def bboxes_iou(a, b, xyxy=True):
    a, ap = _c(np.asarray(a, dtype=np.float32).reshape(-1, 4))
    b, bp = _c(np.asarray(b, dtype=np.float32).reshape(-1, 4))
    out = np.empty((a.shape[0], b.shape[0]), dtype=np.float32)
    lib().orc_bboxes_iou(ap, a.shape[0], bp, b.shape[0], 1 if xyxy else 0, out.ctypes.data_as(_fp))
    return out


def build_target(pred, labels, layer_no, n_classes=80, ignore_thre=0.7, anchors=ANCHORS_PX, mask=ANCHOR_MASK):
    """pred [B,3,F,F,4] (any strides, fp32 numpy), labels [B,K,5] -> (target, obj_mask, tgt_mask, tgt_scale)."""
    pred = np.asarray(pred)
    assert pred.dtype == np.float32
    B, A, F, _, _ = pred.shape
    assert A == 3
    strides_el = (ctypes.c_long * 5)(*[s // 4 for s in pred.strides])
    labels, lp = _c(labels)
    K = labels.shape[1]
    anch, ap = _c(np.array(anchors, dtype=np.float32).ravel())
    am = (ctypes.c_int * 3)(*mask[layer_no])
    target = np.empty((B, 3, F, F, 5 + n_classes), dtype=np.float32)
    obj_mask = np.empty((B, 3, F, F), dtype=np.float32)
    tgt_mask = np.empty((B, 3, F, F, 4 + n_classes), dtype=np.float32)
    tgt_scale = np.empty((B, 3, F, F, 2), dtype=np.float32)
    rc = lib().orc_build_target(pred.ctypes.data_as(_fp), strides_el, lp, B, F, K, n_classes, layer_no, ap, am,
                                float(np.float32(ignore_thre)), target.ctypes.data_as(_fp), obj_mask.ctypes.data_as(_fp),
                                tgt_mask.ctypes.data_as(_fp), tgt_scale.ctypes.data_as(_fp))
    if rc != 0:
        raise IndexError("a matched GT indexes outside the grid (reference: IndexError at yololoss.py:330)")
    return target, obj_mask, tgt_mask, tgt_scale


def yolo_loss_layer(raw, labels, layer_no, n_classes=80, ignore_thre=0.7, anchors=ANCHORS_PX, mask=ANCHOR_MASK):
    """N2: one layer of YOLOLoss.forward (yolo/model/yololoss.py:390-432) starting from the raw head tensor.

    raw [B,3*(5+C),F,F] -> train-mode YOLOLayer (decode_train, yololayer.py:122-145) -> build_target (:118-371) -> the masked
    BCE / MSE sums of :402-432.  Returns (losses = float64 [xy, wh, obj, cls], grad_raw [B,3*(5+C),F,F] float64), the gradient
    of their sum with respect to raw: ATen's binary_cross_entropy_backward ((x - t) / max((1 - x) x, 1e-12) * weight),
    mse_loss_backward (2 (a - b)), the in-place mask multiplications of :402-407 and sigmoid' on the xy / obj / cls channels.
    numpy restatement (float32 values, float64 accumulation); the mask / target logic is the C oracle's."""
    raw = np.ascontiguousarray(raw, dtype=np.float32)
    B, _, F, _ = raw.shape
    nch = 5 + n_classes
    output, pred = decode_train(raw, layer_no, n_classes, anchors, mask)             # [B,3,F,F,nch], [B,3,F,F,4]
    target, obj_mask, tgt_mask, tgt_scale = build_target(pred, np.asarray(labels, np.float32), layer_no, n_classes, ignore_thre,
                                                         anchors, mask)
    sel = np.r_[0:4, 5:nch]
    out_m = np.array(output, dtype=np.float32)                                       # :402-407 (float32 products, in this order)
    out_m[..., 4] = out_m[..., 4] * obj_mask
    out_m[..., sel] = out_m[..., sel] * tgt_mask
    out_m[..., 2:4] = out_m[..., 2:4] * tgt_scale
    tgt = np.array(target, dtype=np.float32)                                         # :411-413
    tgt[..., 4] = tgt[..., 4] * obj_mask
    tgt[..., sel] = tgt[..., sel] * tgt_mask
    tgt[..., 2:4] = tgt[..., 2:4] * tgt_scale

    def bce(x, t, w=None):                                                           # ATen binary_cross_entropy, reduction sum
        with np.errstate(divide="ignore"):
            lx = np.maximum(np.log(x.astype(np.float32)), np.float32(-100.0)).astype(np.float64)
            l1x = np.maximum(np.log((np.float32(1.0) - x).astype(np.float32)), np.float32(-100.0)).astype(np.float64)
        v = (t.astype(np.float64) - 1.0) * l1x - t.astype(np.float64) * lx
        if w is not None:
            v = v * w.astype(np.float64)
        return v

    def bce_grad(x, t, w=None):
        x64, t64 = x.astype(np.float64), t.astype(np.float64)
        g = (x64 - t64) / np.maximum((1.0 - x64) * x64, 1e-12)
        return g * w.astype(np.float64) if w is not None else g

    w_xy = (tgt_scale * tgt_scale).astype(np.float32)                                # :417
    losses = np.array([
        bce(out_m[..., :2], tgt[..., :2], w_xy).sum(),                               # :421
        (np.square(out_m[..., 2:4].astype(np.float64) - tgt[..., 2:4].astype(np.float64))).sum() / 2.0,   # :423
        bce(out_m[..., 4], tgt[..., 4]).sum(),                                       # :425
        bce(out_m[..., 5:], tgt[..., 5:]).sum(),                                     # :427
    ], dtype=np.float64)
    g_m = np.zeros(out_m.shape, np.float64)                                          # d loss / d out_m
    g_m[..., :2] = bce_grad(out_m[..., :2], tgt[..., :2], w_xy)
    g_m[..., 2:4] = out_m[..., 2:4].astype(np.float64) - tgt[..., 2:4].astype(np.float64)
    g_m[..., 4] = bce_grad(out_m[..., 4], tgt[..., 4])
    g_m[..., 5:] = bce_grad(out_m[..., 5:], tgt[..., 5:])
    g_o = np.array(g_m)                                                              # back through the mask products
    g_o[..., 2:4] = g_o[..., 2:4] * tgt_scale
    g_o[..., sel] = g_o[..., sel] * tgt_mask
    g_o[..., 4] = g_o[..., 4] * obj_mask
    o64 = np.asarray(output, np.float64)                                             # sigmoid' on xy / obj / cls, identity on wh
    dsig = o64 * (1.0 - o64)
    dsig[..., 2:4] = 1.0
    g_raw = (g_o * dsig).transpose(0, 1, 4, 2, 3).reshape(B, 3 * nch, F, F)
    return losses, g_raw
