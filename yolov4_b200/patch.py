"""patch_reference() -- install the B200 path over an importable zjykzj/YOLOv4 checkout."""
import importlib


def patch_reference(loss_forward=True):
    """Monkey-patches the three symbols of the drop-in boundary:
        yolo.model.yololayer.YOLOLayer, yolo.util.utils.postprocess, yolo.model.yololoss.YOLOLoss.build_target
    and the names other reference modules imported from them.  With loss_forward=True (default) YOLOLoss.forward is
    rebound as well (row A8: the same arithmetic written out of place, one pinned label upload for the three layers, row
    N4); with False the reference's own forward keeps running on top of the B200 YOLOLayer and build_target -- it
    multiplies dict['output'] by its masks in place (yololoss.py:402-408), which the autograd node of the B200 YOLOLayer
    tolerates because it keeps the raw tensor, not its output.  Returns the list of patched attributes."""
    # (not `from . import postprocess`: the package attribute of that name is the function re-exported by __init__)
    my_layer = importlib.import_module(".yololayer", __package__)
    my_post = importlib.import_module(".postprocess", __package__)
    my_loss = importlib.import_module(".yololoss", __package__)
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
    if loss_forward:
        ref_loss.YOLOLoss.forward = my_loss.YOLOLoss.forward
        done.append("yolo.model.yololoss.YOLOLoss.forward")
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
