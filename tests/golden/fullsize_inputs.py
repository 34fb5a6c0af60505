"""Inputs of the BASELINE-size goldens (fullsize.npz), shared by the generator (make_golden.py, build container, imports the
reference) and the tests (GPU box, no reference): everything is seeded and generated on the CPU, so both sides see the
same bits."""
import numpy as np
import torch

IMG, C, B = 608, 80, 2
SEED_RAW, SEED_LAB = 41, 43
SETTINGS = (("val", 1e-4, 0.4), ("det", 0.2, 0.5))          # BASELINE configs 2 and 3 (val.py / detect.py settings)
N_GT = 50                                                   # BASELINE config 4


def raws_cpu():
    from yolov4_b200.synth import synth_head_outputs
    return synth_head_outputs(B, IMG, C, seed=SEED_RAW)


def labels_cpu():
    from yolov4_b200.synth import synth_labels
    lab = synth_labels(B, IMG, n_valid=N_GT, seed=SEED_LAB)
    lab[1, 7] = lab[1, 6]                                   # duplicate GT (same cell, same anchor: last writer wins)
    lab[1, 8, :4] = lab[1, 6, :4]
    lab[1, 8, 4] = (lab[1, 6, 4] + 3) % C                   # same box, other class: one-hots accumulate
    return lab


def plant_pred(pred, labels, layer):
    """Copies a seeded subset of the GT boxes (jittered) into `pred` [B,3,F,F,4] (numpy, grid units, modified in place) so that
    the ignore mask of build_target has zeros (a random prediction almost never reaches IoU 0.7 with a GT)."""
    rng = np.random.RandomState(1000 + layer)
    s = float(8 << layer)
    F = pred.shape[2]
    lab = np.asarray(labels, np.float32)
    for b in range(pred.shape[0]):
        for t in range(lab.shape[1]):
            if lab[b, t].sum() > 0 and rng.rand() < 0.6:
                i, j = min(int(lab[b, t, 0] / s), F - 1), min(int(lab[b, t, 1] / s), F - 1)
                a = rng.randint(0, 3)
                pred[b, a, j, i, :] = (lab[b, t, :4] / np.float32(s)) * np.float32(1.0 + 0.04 * rng.randn())
    return pred


def corners_obj(pred):
    """(x1, y1, x2, y2, obj) of every box of a decoded [B,M,5+C] tensor, with the reference's fp32 arithmetic (utils.py:117-126)."""
    p = torch.as_tensor(pred)
    c = torch.empty(p.shape[0], p.shape[1], 5, dtype=torch.float32)
    c[:, :, 0] = p[:, :, 0] - p[:, :, 2] / 2
    c[:, :, 1] = p[:, :, 1] - p[:, :, 3] / 2
    c[:, :, 2] = p[:, :, 0] + p[:, :, 2] / 2
    c[:, :, 3] = p[:, :, 1] + p[:, :, 3] / 2
    c[:, :, 4] = p[:, :, 4]
    return c.numpy()


def rows_from_kept(pred, counts, kept_idx, kept_cls):
    """Rebuilds postprocess' output rows (x1,y1,x2,y2,obj,cls_conf,cls) from the kept (box, class) pairs (utils.py:177-184)."""
    pred = np.asarray(pred)
    co = corners_obj(pred)
    rows, o = [], 0
    for b, n in enumerate(counts):
        idx, cls = kept_idx[o:o + n].astype(np.int64), kept_cls[o:o + n].astype(np.int64)
        r = np.empty((n, 7), np.float32)
        r[:, :5] = co[b, idx]
        r[:, 5] = pred[b, idx, 5 + cls]
        r[:, 6] = cls.astype(np.float32)
        rows.append(r)
        o += n
    return rows


def sparse_pack(dense, background):
    flat = np.ascontiguousarray(dense).reshape(-1)
    idx = np.flatnonzero(flat.view(np.uint32) != np.float32(background).view(np.uint32)).astype(np.int64)
    return idx, flat[idx].copy()


def sparse_unpack(shape, background, idx, val):
    flat = np.full(int(np.prod(shape)), background, np.float32)
    flat[idx] = val
    return flat.reshape(shape)
