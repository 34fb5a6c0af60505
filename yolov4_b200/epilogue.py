"""Per-detection epilogue of the reference's callers, one kernel instead of a Python loop.

coco_rows(outputs, img_infos, image_ids, class_ids)   validate(): engine/build.py:146-164 + utils.yolobox2xywh
detect_rows(outputs, img_infos, class_ids)            detect.parse_info(): detect.py:171-179 + utils.yolobox2yxyx
"""
import torch

from . import _cabi


def _run(outputs, img_infos, image_ids, class_ids, mode):
    present = [(i, o) for i, o in enumerate(outputs) if o is not None and o.shape[0] > 0]
    dev = present[0][1].device if present else torch.device("cuda")
    if not present:
        return torch.zeros((0, 7), dtype=torch.float64, device=dev)
    if not present[0][1].is_cuda:
        raise TypeError("epilogue needs CUDA tensors (postprocess outputs); there is no CPU fallback")
    rows = torch.cat([o.to(torch.float32) for _, o in present], 0).contiguous()
    row_image = torch.cat([torch.full((o.shape[0],), i, dtype=torch.int32, device=dev) for i, o in present])
    info = torch.tensor([[float(v) for v in inf[:4]] for inf in img_infos], dtype=torch.float64, device=dev)
    ids = torch.tensor([int(v) for v in image_ids], dtype=torch.int64, device=dev)
    cids = torch.tensor([int(v) for v in class_ids], dtype=torch.int32, device=dev)
    out = torch.empty((rows.shape[0], 7), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().yl_coco_rows(rows.data_ptr(), row_image.data_ptr(), rows.shape[0], info.data_ptr(), ids.data_ptr(),
                                             cids.data_ptr(), len(class_ids), mode, out.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream))
    return out


def coco_rows(outputs, img_infos, image_ids, class_ids):
    """outputs: list (one entry per image) of postprocess rows [K_i,7] or None; img_infos[i] = (src_h, src_w, dst_h, dst_w, ...).
    Returns float64 [sum K_i, 7] = (image_id, category_id, x, y, w, h, score), the numbers validate() puts in its COCO dicts."""
    return _run(outputs, img_infos, image_ids, class_ids, 0)


def coco_dicts(outputs, img_infos, image_ids, class_ids):
    """The reference's `data_list` (engine/build.py:158-164) built from coco_rows with a single device->host copy."""
    t = coco_rows(outputs, img_infos, image_ids, class_ids).cpu().tolist()
    return [{"image_id": int(r[0]), "category_id": int(r[1]), "bbox": [r[2], r[3], r[4], r[5]], "score": r[6], "segmentation": []}
            for r in t]


def detect_rows(outputs, img_infos, class_ids):
    """Returns float64 [sum K_i, 7] = (image index, category_id, y1, x1, y2, x2, cls_conf) as detect.parse_info computes them."""
    return _run(outputs, img_infos, list(range(len(outputs))), class_ids, 1)


def coco_rows_padded(rows, counts, img_infos, image_ids, class_ids, mode=0, total=None):
    """The same rows from the padded device output of the fixed-shape path (HeadPostprocessor.run: rows [B, cap_out, 7], counts =
    meta[:B]) with ONE launch for the whole batch -- no torch.cat, no per-image index tensor.  `total` = sum of the counts when the
    caller already knows it (HeadPostprocessor.results() has read them); otherwise one D2H read here.
    mode 0: validate()'s COCO rows, mode 1: detect.parse_info()'s rows."""
    if not rows.is_cuda or rows.dtype != torch.float32 or rows.dim() != 3 or rows.shape[2] != 7 or not rows.is_contiguous():
        raise TypeError("rows must be a contiguous float32 CUDA tensor [B, cap_out, 7]; there is no CPU fallback")
    B, cap_out = int(rows.shape[0]), int(rows.shape[1])
    dev = rows.device
    counts = counts[:B].to(device=dev, dtype=torch.int32).contiguous()
    if total is None:
        total = int(counts.clamp(0, cap_out).sum())
    info = torch.tensor([[float(v) for v in inf[:4]] for inf in img_infos], dtype=torch.float64, device=dev)
    ids = torch.tensor([int(v) for v in image_ids], dtype=torch.int64, device=dev)
    cids = torch.tensor([int(v) for v in class_ids], dtype=torch.int32, device=dev)
    out = torch.empty((total, 7), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().yl_coco_rows_padded(rows.data_ptr(), counts.data_ptr(), B, cap_out, info.data_ptr(), ids.data_ptr(),
                                                    cids.data_ptr(), len(class_ids), int(mode), out.data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream))
    return out
