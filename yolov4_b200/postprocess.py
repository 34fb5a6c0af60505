"""postprocess -- drop-in for yolo/util/utils.py:92-223 (and the nms() it calls, :32-89), plus the fused entry.

  postprocess(prediction, num_classes, conf_thre, nms_thre)        the reference's exact call surface
  detect_raw(head_outputs, num_classes, conf_thre, nms_thre, ...)  fused: raw head tensors -> detections,
                                                                   never materialising the dense [B,M,5+C] tensor
  HeadPostprocessor                                                fixed-shape, sync-free, CUDA-graph replayable

All compute runs in libyolohead.so (hand-written sm_100a kernels) through the C ABI; there is no eager fallback.
"""
import threading

import numpy as np
import torch

from . import _cabi

ANCHORS_PX = [[12, 16], [19, 36], [40, 28], [36, 75], [76, 55], [72, 146], [142, 110], [192, 243], [459, 401]]
ANCHOR_MASK = [[0, 1, 2], [3, 4, 5], [6, 7, 8]]

_DEFAULT_CAP_SEG = 1024
_DEFAULT_CAP_OUT = 16384


def _f32(x):
    """Python-float thresholds are compared in fp32 by torch / NumPy."""
    return float(np.float32(x))


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _next_pow2(n):
    p = 1
    while p < n:
        p <<= 1
    return p


class _Workspace:
    """Device scratch for one (device, stream, thread, B, M, C, cap_seg): calls of the same shape on different CUDA streams
    or host threads may run concurrently and must not share candidate buffers and counters."""

    _cache = {}

    def __init__(self, device, B, M, C, cap_seg):
        self.B, self.M, self.C, self.cap_seg = B, M, C, cap_seg
        self.nbytes = _cabi.lib().yl_post_workspace_bytes(B, M, C, cap_seg)
        if self.nbytes == 0:
            raise _cabi.YoloHeadError("invalid postprocess shape B=%d M=%d C=%d" % (B, M, C))
        self.buf = torch.empty(self.nbytes, dtype=torch.uint8, device=device)

    @classmethod
    def get(cls, device, B, M, C, cap_seg):
        key = (str(device), int(torch.cuda.current_stream(device).cuda_stream), threading.get_ident(), B, M, C, cap_seg)
        ws = cls._cache.get(key)
        if ws is None:
            # keep one workspace per (device, stream, thread, B, M, C): drop smaller-capacity predecessors
            for k in [k for k in cls._cache if k[:6] == key[:6]]:
                del cls._cache[k]
            ws = cls._cache[key] = cls(device, B, M, C, cap_seg)
        return ws

    def ptr(self):
        return self.buf.data_ptr()


_cap_hint = {}    # (B, M, C) -> (cap_seg, cap_out) that were last sufficient


def _split(rows, meta_host, B, out_device, out_dtype):
    out = []
    for b in range(B):
        k = int(meta_host[b])
        if k == 0:
            out.append(None)                         # utils.py:132,159-160
        else:
            r = rows[b, :k]
            if r.device != out_device or r.dtype != out_dtype:
                r = r.to(device=out_device, dtype=out_dtype)
            out.append(r)
    return out


def _run(front_end, device, B, M, C, nms_thre):
    """Runs front end + NMS, growing cap_seg / cap_out and re-running when the device reports an overflow."""
    L = _cabi.lib()
    cap_seg, cap_out = _cap_hint.get((B, M, C), (_DEFAULT_CAP_SEG, _DEFAULT_CAP_OUT))
    cap_seg = min(cap_seg, _next_pow2(M))
    while True:
        ws = _Workspace.get(device, B, M, C, cap_seg)
        rows = torch.empty((B, cap_out, 7), dtype=torch.float32, device=device)
        meta = torch.empty((3 * B,), dtype=torch.int32, device=device)
        _cabi.check(L.yl_post_reset(ws.ptr(), ws.nbytes, B, M, C, cap_seg, _stream()))
        front_end(ws, cap_seg)
        _cabi.check(L.yl_nms(ws.ptr(), ws.nbytes, B, M, C, cap_seg, _f32(nms_thre), rows.data_ptr(), cap_out,
                             meta.data_ptr(), 0, B, _stream()))
        mh = meta.cpu().numpy()                      # the one D2H sync the list-of-tensors contract requires
        max_seg, max_rows = int(mh[B:2 * B].max()), int(mh[:B].max())
        if max_seg > cap_seg:
            cap_seg = _next_pow2(max_seg)
            cap_out = max(cap_out, _next_pow2(int(mh[2 * B:].max())))
            continue
        if max_rows > cap_out:
            cap_out = _next_pow2(max_rows)
            continue
        _cap_hint[(B, M, C)] = (cap_seg, cap_out)
        return rows, mh


def postprocess(prediction, num_classes, conf_thre=0.7, nms_thre=0.45):
    """Same contract as the reference's postprocess (yolo/util/utils.py:92).

    prediction: [B, M, 5+C] decoded (cx, cy, w, h, obj, cls...), fp32, on a CUDA device (a CPU tensor, as
    detect.py:115-118 passes, is moved to the current CUDA device and the results are moved back).
    Returns a list of B entries: Tensor[K_i, 7] = (x1, y1, x2, y2, obj_conf, cls_conf, cls_idx), classes
    ascending then score descending, or None for images without detections.
    Differences from the reference, none observable by its callers: `prediction[:, :, :4]` is not overwritten
    in place with the corner form (utils.py:126), and score ties are ordered (score desc, box index desc),
    the order of a stable argsort, where NumPy's default sort leaves them unspecified (utils.py:58).
    """
    if prediction.dim() != 3 or prediction.shape[2] < 6:
        raise ValueError("prediction must be [B, M, 5+C]")
    if prediction.dtype != torch.float32:
        raise TypeError("prediction must be float32 (the reference path runs fp32 / apex O0)")
    in_device = prediction.device
    pred = prediction if prediction.is_cuda else prediction.cuda()
    pred = pred.contiguous()
    B, M, nch = pred.shape
    C = nch - 5
    if B == 0:
        return []
    num_classes = int(num_classes)
    L = _cabi.lib()
    with torch.cuda.device(pred.device):
        def front(ws, cap_seg):
            _cabi.check(L.yl_filter_dense(pred.data_ptr(), B, M, C, num_classes, _f32(conf_thre), ws.ptr(), ws.nbytes,
                                          cap_seg, 0, B, _stream()))
        rows, mh = _run(front, pred.device, B, M, C, nms_thre)
    return _split(rows, mh, B, in_device, prediction.dtype)


def _check_raws(head_outputs, num_classes):
    # the library takes the scale of a head tensor from its position (stride 8 << l, anchor_mask[l]; yololayer.py:54,60):
    # a caller with fewer heads passes them in the reference's order starting at the stride-8 head
    if len(head_outputs) < 1 or len(head_outputs) > 3:
        raise ValueError("expected 1..3 head tensors (stride 8, 16, 32 in this order)")
    if any(head_outputs[i].shape[2] < head_outputs[i + 1].shape[2] for i in range(len(head_outputs) - 1)):
        raise ValueError("head tensors must be ordered stride 8, 16, 32 (largest grid first; yolov4.py:324)")
    B = head_outputs[0].shape[0]
    Fs = []
    raws = []
    for r in head_outputs:
        if r.dim() != 4 or r.shape[0] != B or r.shape[1] != 3 * (5 + num_classes) or r.shape[2] != r.shape[3]:
            raise ValueError("head tensor must be [B, 3*(5+C), F, F], got %s" % (tuple(r.shape),))
        if r.dtype != torch.float32 or not r.is_cuda:
            raise TypeError("head tensors must be float32 CUDA tensors")
        raws.append(r.contiguous())
        Fs.append(int(r.shape[2]))
    return raws, B, Fs


def detect_raw(head_outputs, num_classes, conf_thre=0.7, nms_thre=0.45, anchors=ANCHORS_PX, anchor_mask=ANCHOR_MASK):
    """Fused YOLOLayer eval decode (yololayer.py:88-166) + postprocess (utils.py:92-223).

    head_outputs: the three raw head conv outputs [B, 3*(5+C), F_l, F_l] (stride 8, 16, 32; yolov4.py:235-251).
    Equivalent to postprocess(torch.cat([YOLOLayer_l(x_l)], 1), num_classes, conf_thre, nms_thre) but reads the
    raw tensors once and never writes the dense decoded tensor.
    """
    raws, B, Fs = _check_raws(head_outputs, num_classes)
    C = int(num_classes)
    M = sum(3 * f * f for f in Fs)
    L = _cabi.lib()
    dev = raws[0].device
    anch = _cabi.floats([v for wh in anchors for v in wh])
    mask = _cabi.ints([v for m in anchor_mask for v in m])
    rp = _cabi.ptrs([r.data_ptr() for r in raws])
    fs = _cabi.ints(Fs)
    with torch.cuda.device(dev):
        def front(ws, cap_seg):
            _cabi.check(L.yl_filter_raw(rp, fs, len(raws), B, C, anch, mask, _f32(conf_thre), ws.ptr(), ws.nbytes, M,
                                        cap_seg, 0, B, _stream()))
        rows, mh = _run(front, dev, B, M, C, nms_thre)
    return _split(rows, mh, B, dev, torch.float32)


class HeadPostprocessor:
    """Fixed-shape fused decode + filter + NMS with no host synchronisation: the serving / benchmark form.

    run(head_outputs) enqueues the whole chain on the current stream and returns (rows [B,cap_out,7], meta [3B])
    device tensors owned by this object (overwritten by the next run).  With n_groups > 1 image groups are pipelined
    over two streams (NMS of group g next to the streaming filter of group g+1); on B200 one group is fastest because
    every kernel already fills the machine.  capture() records the chain into a CUDA graph bound to the given input
    tensors; replay() launches it.
    """

    def __init__(self, batch, grid_sizes, num_classes, conf_thre, nms_thre, device=None, cap_seg=_DEFAULT_CAP_SEG,
                 cap_out=_DEFAULT_CAP_OUT, n_groups=1, anchors=ANCHORS_PX, anchor_mask=ANCHOR_MASK, side_priority=0,
                 mode="nms_side"):
        self.L = _cabi.lib()
        self.device = torch.device(device if device is not None else "cuda")
        self.B, self.Fs, self.C = int(batch), [int(f) for f in grid_sizes], int(num_classes)
        self.M = sum(3 * f * f for f in self.Fs)
        self.conf, self.nms = _f32(conf_thre), _f32(nms_thre)
        self.cap_seg, self.cap_out = int(cap_seg), int(cap_out)
        self.n_groups = max(1, min(int(n_groups), self.B))
        self.anch = _cabi.floats([v for wh in anchors for v in wh])
        self.mask = _cabi.ints([v for m in anchor_mask for v in m])
        self.fs = _cabi.ints(self.Fs)
        with torch.cuda.device(self.device):
            self.ws = _Workspace(self.device, self.B, self.M, self.C, self.cap_seg)
            self.rows = torch.empty((self.B, self.cap_out, 7), dtype=torch.float32, device=self.device)
            self.meta = torch.zeros((3 * self.B,), dtype=torch.int32, device=self.device)
            self.side = torch.cuda.Stream(device=self.device, priority=side_priority)
            self.side2 = torch.cuda.Stream(device=self.device, priority=side_priority)
        self.mode = mode
        self.graph = None
        # kernels launched per run(): per group flag + emit + segment NMS + fix-up (+ 1 memset node)
        self.launches_per_run = self.n_groups * 5

    def run(self, head_outputs):
        L, B, C, M = self.L, self.B, self.C, self.M
        raws, Bc, Fs = _check_raws(head_outputs, C)      # shape / dtype / device checked, non-contiguous inputs copied
        if Bc != B or Fs != self.Fs:
            raise ValueError("shape mismatch with the configured postprocessor")
        self._live_inputs = raws                         # a contiguous copy must outlive the enqueued kernels
        rp = _cabi.ptrs([r.data_ptr() for r in raws])
        main = torch.cuda.current_stream(self.device)
        G = self.n_groups
        fold_reset = (G == 1 and self.mode not in ("emit_side", "pipe3"))    # one group: the flag kernel zeroes the counters itself
        if not fold_reset:
            _cabi.check(L.yl_post_reset(self.ws.ptr(), self.ws.nbytes, B, M, C, self.cap_seg, main.cuda_stream))
        for g in range(G):
            i0, i1 = B * g // G, B * (g + 1) // G
            if fold_reset:
                _cabi.check(L.yl_filter_raw_stage(rp, self.fs, len(self.Fs), B, C, self.anch, self.mask, self.conf, self.ws.ptr(),
                                                  self.ws.nbytes, M, self.cap_seg, i0, i1 - i0, 7, main.cuda_stream))
                self.side.wait_stream(main)
            elif self.mode in ("emit_side", "pipe3"):
                # streaming flag kernels back to back on the main stream; everything else follows on the side stream
                _cabi.check(L.yl_filter_raw_stage(rp, self.fs, len(self.Fs), B, C, self.anch, self.mask, self.conf, self.ws.ptr(),
                                                  self.ws.nbytes, M, self.cap_seg, i0, i1 - i0, 1, main.cuda_stream))
                self.side.wait_stream(main)
                _cabi.check(L.yl_filter_raw_stage(rp, self.fs, len(self.Fs), B, C, self.anch, self.mask, self.conf, self.ws.ptr(),
                                                  self.ws.nbytes, M, self.cap_seg, i0, i1 - i0, 2, self.side.cuda_stream))
            else:
                _cabi.check(L.yl_filter_raw(rp, self.fs, len(self.Fs), B, C, self.anch, self.mask, self.conf, self.ws.ptr(),
                                            self.ws.nbytes, M, self.cap_seg, i0, i1 - i0, main.cuda_stream))
                self.side.wait_stream(main)
            nms_stream = self.side
            if self.mode == "pipe3":
                # three-deep: flag(g+1) on the main stream, emit(g) on the first side stream, NMS + gather(g-1) on the second
                self.side2.wait_stream(self.side)
                nms_stream = self.side2
            _cabi.check(L.yl_nms(self.ws.ptr(), self.ws.nbytes, B, M, C, self.cap_seg, self.nms, self.rows.data_ptr(),
                                 self.cap_out, self.meta.data_ptr(), i0, i1 - i0, nms_stream.cuda_stream))
        main.wait_stream(self.side)
        if self.mode == "pipe3":
            main.wait_stream(self.side2)
        return self.rows, self.meta

    def capture(self, head_outputs):
        raws, B, Fs = _check_raws(head_outputs, self.C)
        if B != self.B or Fs != self.Fs:
            raise ValueError("shape mismatch with the configured postprocessor")
        self._captured_inputs = raws
        with torch.cuda.device(self.device):
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self.run(raws)                       # warm-up outside capture
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.run(raws)
        return self

    def replay(self):
        self.graph.replay()
        return self.rows, self.meta

    def results(self):
        """Host-side view of the last run: list of [K_i,7] tensors / None (one D2H sync)."""
        mh = self.meta.cpu().numpy()
        if int(mh[self.B:2 * self.B].max()) > self.cap_seg or int(mh[:self.B].max()) > self.cap_out:
            raise _cabi.YoloHeadError("capacity exceeded (max segment %d / cap_seg %d, max rows %d / cap_out %d)" % (
                int(mh[self.B:2 * self.B].max()), self.cap_seg, int(mh[:self.B].max()), self.cap_out))
        return _split(self.rows, mh, self.B, self.device, torch.float32)
