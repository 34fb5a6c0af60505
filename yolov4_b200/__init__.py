"""yolov4_b200 -- B200-native (sm_100a) detection-head hot path of zjykzj/YOLOv4.

Host side mirrors the reference's call surface:
    YOLOLayer(cfg, layer_no, device).forward          yolo/model/yololayer.py
    postprocess(prediction, num_classes, conf, nms)   yolo/util/utils.py:92
    YOLOLoss(cfg, ignore_thresh, device).build_target yolo/model/yololoss.py:118
plus the fused entry detect_raw() / HeadPostprocessor.  All compute runs in libyolohead.so (csrc/, C ABI in
include/yolo_head.h).  Importing the package does not load CUDA; the first call does, and fails loudly if the
library is missing.
"""
from .postprocess import postprocess, detect_raw, HeadPostprocessor, ANCHORS_PX, ANCHOR_MASK  # noqa: F401
from .yololayer import YOLOLayer, decode_dense_cat  # noqa: F401
from .yololoss import YOLOLoss, build_target, build_targets3, fused_yolo_loss, fused_yolo_loss_components  # noqa: F401
from .patch import patch_reference  # noqa: F401
from .epilogue import coco_rows, coco_dicts, detect_rows, coco_rows_padded  # noqa: F401

__all__ = ["postprocess", "detect_raw", "HeadPostprocessor", "YOLOLayer", "decode_dense_cat", "YOLOLoss", "build_target", "fused_yolo_loss", "fused_yolo_loss_components", "patch_reference", "coco_rows", "coco_dicts", "detect_rows",
           "ANCHORS_PX", "ANCHOR_MASK"]
