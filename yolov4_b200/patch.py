"""patch_reference() -- install the B200 path over an importable zjykzj/YOLOv4 checkout."""
import importlib


def patch_reference():
    """Monkey-patches the three symbols of the drop-in boundary:
        yolo.model.yololayer.YOLOLayer, yolo.util.utils.postprocess, yolo.model.yololoss.YOLOLoss.build_target
    and the names other reference modules imported from them.  Returns the list of patched attributes."""
    from . import yololayer as my_layer, postprocess as my_post, yololoss as my_loss
    done = []
    ref_layer = importlib.import_module("yolo.model.yololayer")
    ref_layer.YOLOLayer = my_layer.YOLOLayer
    done.append("yolo.model.yololayer.YOLOLayer")
    ref_utils = importlib.import_module("yolo.util.utils")
    ref_utils.postprocess = my_post.postprocess
    done.append("yolo.util.utils.postprocess")
    ref_loss = importlib.import_module("yolo.model.yololoss")
    ref_loss.YOLOLoss.build_target = my_loss.YOLOLoss.build_target
    done.append("yolo.model.yololoss.YOLOLoss.build_target")
    for modname, attr, val in (("yolo.model.yolov4", "YOLOLayer", my_layer.YOLOLayer),
                               ("yolo.engine.build", "postprocess", my_post.postprocess)):
        try:
            m = importlib.import_module(modname)
        except Exception:        # engine needs apex / pycocotools
            continue
        if hasattr(m, attr):
            setattr(m, attr, val)
            done.append("%s.%s" % (modname, attr))
    return done
