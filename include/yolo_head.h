/*
 * yolo_head.h -- C ABI of libyolohead.so, the B200 (sm_100a) implementation of the YOLOv4 detection-head hot
 * path of zjykzj/YOLOv4.  The reference is pure Python and has no FFI of its own: the
 * reference-side boundary is the Python call surface of three symbols, and this header is what the Python
 * shim (yolov4_b200/_cabi.py, ctypes) binds to replace each of them.  Every entry point cites the
 * reference interface it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; no torch / C++ types.
 *   - every function returns int: 0 = YL_OK, 1..99 = argument errors, 1000+e = cudaError_t e.  Never throws.
 *   - all device-pointer functions are stream-ordered on `stream` (a cudaStream_t passed as void*), do not
 *     allocate, do not synchronise, and are CUDA-graph capturable.
 *   - fp32 everywhere (the reference trains and evaluates with apex O0, README.md:106,112).
 *   - sigmoid / exp / log are the "spec math" of DESIGN.md section 4 (fixed IEEE op sequences), so results are
 *     bit-reproducible and identical to the CPU oracle.
 */
#ifndef YOLO_HEAD_H_
#define YOLO_HEAD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YL_OK 0
#define YL_ERR_ARG 1          /* null pointer / non-positive size / unsupported shape */
#define YL_ERR_CLASSES 2      /* more classes than the candidate bitmask supports (YL_MAX_CLASSES) */
#define YL_ERR_WORKSPACE 3    /* workspace smaller than yl_post_workspace_bytes() */
#define YL_ERR_CAPACITY 4     /* yl_detect_host: a class segment overflowed cap_seg; recreate the context larger */
#define YL_ERR_CUDA_BASE 1000 /* 1000 + cudaError_t */

#define YL_MAX_CLASSES 128
#define YL_ABI_VERSION 1

typedef void *yl_stream_t; /* cudaStream_t */

int yl_abi_version(void);
/* Hash of the sources the library was compiled from (build.py:source_hash); the loader refuses a stale library. */
const char *yl_source_hash(void);
const char *yl_error_string(int code);
/* Self-test of the library's spec math (no reference counterpart): the bounded reciprocal used by the sigmoid against the
 * IEEE quotient for every float in [1, 2^125]; *bad_dev (device u64) receives the number of mismatches. */
int yl_selftest_rcp(unsigned long long *bad_dev, yl_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * A2  YOLOLayer.forward, eval branch                       replaces yolo/model/yololayer.py:88-120,146-166
 *   raw   [B, 3*(5+C), F, F] contiguous device fp32 (head conv output, yolov4.py:235-251)
 *   anch_grid[6] = masked anchors / stride, (w0,h0,w1,h1,w2,h2)               (yololayer.py:73-76)
 *   out   rows of (5+C) floats; box (b,a,y,x) -> row b*rows_per_image + row_offset + a*F*F + y*F + x, so the
 *         three layers can write straight into the torch.cat((x1,x2,x3),1) buffer of yolov4.py:324.
 * The reference's in-place overwrite of `raw` (a side effect nobody reads) is not reproduced.
 * --------------------------------------------------------------------------------------------------------- */
int yl_decode_dense(const float *raw, int B, int F, int C, const float *anch_grid, float stride,
                    float *out, long rows_per_image, long row_offset, yl_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * A3  YOLOLayer.forward, train branch                      replaces yolo/model/yololayer.py:122-145
 *   output_planar [B,3,5+C,F,F]: sigmoid on xy/obj/cls, raw wh.  The reference's dict['output'] is the
 *                 permute(0,1,3,4,2) view of this storage (strides (255F^2, 85F^2, F, 1, F^2)).
 *   pred_planar   [B,3,4,F,F]: grid-unit boxes (no *stride); dict['pred'] is its permuted view.
 * yl_decode_train_backward: grad_raw = grad_out * d(output)/d(raw) (sigmoid' on xy/obj/cls, 1 on wh), for
 * the autograd.Function that keeps `output` differentiable (yololoss.py:402-432 back-props through it).
 * --------------------------------------------------------------------------------------------------------- */
int yl_decode_train(const float *raw, int B, int F, int C, const float *anch_grid,
                    float *output_planar, float *pred_planar, yl_stream_t stream);
int yl_decode_train_backward(const float *output_planar, const float *grad_out_planar, int B, int F, int C,
                             float *grad_raw, yl_stream_t stream);
/* The same gradient from the RAW head tensor (sigmoid recomputed with the bits of the forward pass).  This is the form the
 * autograd node uses: the reference's YOLOLoss.forward multiplies dict['output'] by its masks in place
 * (yololoss.py:402-408), so the tensor the forward returned cannot be the one kept for the backward pass. */
int yl_decode_train_backward_raw(const float *raw, const float *grad_out_planar, int B, int F, int C,
                                 float *grad_raw, yl_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * A4+A5  postprocess(prediction, num_classes, conf_thre, nms_thre)   replaces yolo/util/utils.py:92-223
 *        (and nms, utils.py:32-89, which it calls per (image, class))
 *
 * Two front ends fill the same candidate workspace, one back end consumes it:
 *   yl_filter_raw     fused decode + confidence filter straight from the three raw head tensors (the 495 MB
 *                     dense [B,M,5+C] tensor is never materialised)          yololayer.py:88-166 + utils.py:117-184
 *   yl_filter_dense   confidence filter of an already decoded [B,M,5+C] tensor -- the literal
 *                     postprocess(prediction, ...) entry                      utils.py:117-184
 *   yl_nms            per (image,class) sort (score desc, box index desc) + greedy NMS (drop at IoU >= thr)
 *                     + class-ascending concatenation                        utils.py:32-89, :191-220
 *
 * Workspace: one device buffer of yl_post_workspace_bytes(B, M, C, cap_seg) bytes.  cap_seg = capacity of one
 * (image,class) candidate segment.  Counting is always exact; a segment that receives more than cap_seg
 * candidates is reported through meta[] and the caller re-runs with a larger cap_seg (the shim does).
 *
 * Output: out_rows [B, cap_out, 7] = (x1,y1,x2,y2,obj_conf,cls_conf,cls_idx), image b's K_b rows first;
 *         meta [3*B] int32: meta[b] = K_b (true count, may exceed cap_out -> rows truncated),
 *                            meta[B+b] = max candidates seen in any class segment of image b,
 *                            meta[2B+b] = total candidates of image b.
 * --------------------------------------------------------------------------------------------------------- */
size_t yl_post_workspace_bytes(int B, long M, int C, int cap_seg);

/* Zeroes the per-run counters.  Must precede yl_filter_* on the same stream. */
int yl_post_reset(void *ws, size_t ws_bytes, int B, long M, int C, int cap_seg, yl_stream_t stream);

/* raw[l] = [B, 3*(5+C), F[l], F[l]], l = 0..n_layers-1 (n_layers <= 3), strides 8/16/32 (yololayer.py:54);
 * anchors_px[18] and anchor_mask[9] are cfg MODEL.ANCHORS / ANCHOR_MASK.  M must equal sum 3*F[l]^2.
 * img_first/img_count select a contiguous image range (used to pipeline image groups on several streams). */
int yl_filter_raw(const float *const *raw, const int *F, int n_layers, int B, int C,
                  const float *anchors_px, const int *anchor_mask, float conf_thre,
                  void *ws, size_t ws_bytes, long M, int cap_seg, int img_first, int img_count,
                  yl_stream_t stream);

/* The two kernels of yl_filter_raw run separately (measurement / pipelining): stages bit 0 = streaming flag kernel
 * (reads every raw byte once, writes 16 B per box), bit 1 = emit kernel (resolves the flagged pairs exactly).
 * Bit 2 (with bit 0 or 1): the call also resets the workspace counters, i.e. it replaces yl_post_reset -- inside the flag
 * kernel where that exists, so a step has one graph node less; only for the first (or only) image group of a step. */
int yl_filter_raw_stage(const float *const *raw, const int *F, int n_layers, int B, int C,
                        const float *anchors_px, const int *anchor_mask, float conf_thre,
                        void *ws, size_t ws_bytes, long M, int cap_seg, int img_first, int img_count,
                        int stages, yl_stream_t stream);

/* pred [B, M, 5+C] decoded (cx,cy,w,h,obj,cls...), not modified (the reference's in-place xyxy overwrite,
 * utils.py:126, is an unobserved side effect).  num_classes <= C limits the row pre-filter (utils.py:139). */
int yl_filter_dense(const float *pred, int B, long M, int C, int num_classes, float conf_thre,
                    void *ws, size_t ws_bytes, int cap_seg, int img_first, int img_count, yl_stream_t stream);

int yl_nms(void *ws, size_t ws_bytes, int B, long M, int C, int cap_seg, float nms_thre,
           float *out_rows, long cap_out, int *meta, int img_first, int img_count, yl_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * A6+A7  YOLOLoss.build_target(output, pred, layer_no, labels)      replaces yolo/model/yololoss.py:118-371
 *        (and bboxes_iou, yololoss.py:16-91, which it calls twice per image)
 *   pred          element (b,a,j,i,k) at pred[b*ps[0] + a*ps[1] + j*ps[2] + i*ps[3] + k*ps[4]] (floats); the
 *                 reference passes a non-contiguous view with strides (255F^2, 85F^2, F, 1, F^2)
 *   labels        [B,K,5] fp32 (xc,yc,w,h,cls) in input pixels, zero padded (transform.py:464-471)
 *   outputs       dense contiguous fp32: target [B,3,F,F,5+C], obj_mask [B,3,F,F], tgt_mask [B,3,F,F,4+C],
 *                 tgt_scale [B,3,F,F,2]  (yololoss.py:156-167)
 *   status        device int32[1], set non-zero if a matched GT indexes outside the grid (the reference
 *                 raises IndexError there); may be NULL.
 * --------------------------------------------------------------------------------------------------------- */
int yl_build_target(const float *pred, const long *pred_strides, const float *labels, int B, int F, int K,
                    int C, int layer_no, const float *anchors_px, const int *anchor_mask3, float ignore_thre,
                    float *target, float *obj_mask, float *tgt_mask, float *tgt_scale, int *status,
                    yl_stream_t stream);
/* The same for up to three scales in ONE pair of launches (the form YOLOLoss.forward uses, yololoss.py:381-393 loops over the
 * layers): the small grids' latency-bound CTAs run next to the 76x76 ones.  pred[l] / target[l] / ... per scale as above,
 * ps15 = the n_layers stride quintuples back to back, F[l] the grid sizes, layer_no[l] the scale numbers, anchor_mask = the
 * full 3x3 mask table (row layer_no[l] is used).  Results are identical to n_layers calls of yl_build_target. */
int yl_build_target3(const float *const *pred, const long *ps15, const float *labels, int B, const int *F, int K,
                     int C, int n_layers, const int *layer_no, const float *anchors_px, const int *anchor_mask,
                     float ignore_thre, float *const *target, float *const *obj_mask, float *const *tgt_mask,
                     float *const *tgt_scale, int *status, yl_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * N1  per-detection epilogue of the reference's callers
 *   mode 0: validate()  yolo/engine/build.py:146-164 + yolobox2xywh (yolo/util/utils.py:281-309), float64 like the Python loop
 *           out row = (image_id, category_id, x, y, w, h, score = obj_conf*cls_conf)
 *   mode 1: detect.parse_info()  detect.py:171-179 + yolobox2yxyx (utils.py:312-340), fp32 box arithmetic
 *           out row = (image_id, category_id, y1, x1, y2, x2, cls_conf)
 *   rows [K,7] fp32 (postprocess output rows back to back), row_image [K] image index of each row,
 *   img_info [n_img,4] = (src_h, src_w, dst_h, dst_w) float64, image_ids [n_img] int64, class_ids [n_classes] int32 (all device)
 * --------------------------------------------------------------------------------------------------------- */
int yl_coco_rows(const float *rows, const int *row_image, long K, const double *img_info, const long long *image_ids,
                 const int *class_ids, int n_classes, int mode, double *out, yl_stream_t stream);
/* The same epilogue straight on the padded device output of yl_nms (rows [B, cap_out, 7] + counts [B] = meta[0..B)): one launch
 * for the whole batch, no concatenation and no per-image index tensor; out is compact, [sum counts, 7], image b's rows at the
 * exclusive prefix of the counts. */
int yl_coco_rows_padded(const float *rows, const int *counts, int B, long cap_out, const double *img_info,
                        const long long *image_ids, const int *class_ids, int n_classes, int mode, double *out,
                        yl_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * N2  YOLOLoss.forward for one layer, fused            replaces yolo/model/yololoss.py:390-432
 *                    (with the train-mode YOLOLayer, yololayer.py:122-145, and build_target, :118-371, it consumes)
 *   raw        [B, 3*(5+C), F, F] raw head tensor of the layer (device fp32)
 *   labels     [B,K,5] fp32 as for yl_build_target
 *   loss4      device float64[4], ACCUMULATED into: loss_xy (:421), loss_wh (:423), loss_obj (:425), loss_cls (:427);
 *              the layer's loss is their sum (:432); zero it before the first layer
 *   saved for yl_loss_backward: gobj [B,3,F,F] fp32 (d loss / d raw objectness), tcell_all / mcell int32 [B,K],
 *              mgrad [B,K,4+C] fp32 (d loss / d raw xy,wh,classes of the matched cells)
 *   status     device int32[1] (caller zeroes it), set non-zero when a matched GT indexes outside the grid or carries a
 *              class id outside [0, C) (the reference raises IndexError / writes another channel there); may be NULL
 * yl_loss_backward: grad_raw [B, 3*(5+C), F, F] = upstream[0] * d loss / d raw (upstream: device fp32 scalar).
 * None of the reference's dense output / pred / target / mask tensors is materialised: 5 of the 5+C planes are read.
 * Aliasing rule: calls for different layers issued back to back on one stream (yl_loss_forward_chained) overlap on the device
 * (programmatic dependent launch: the next layer's matching runs in the tail of this layer's objectness kernel), so gobj / tcell_all /
 * mcell / mgrad of consecutive calls must be distinct buffers (one set per layer, as yl_loss_backward needs anyway);
 * loss4 and status are shared and only accumulated into.  YL_PDL=0 restores plain stream order.
 * --------------------------------------------------------------------------------------------------------- */
int yl_loss_forward(const float *raw, const float *labels, int B, int F, int K, int C, int layer_no,
                    const float *anchors_px, const int *anchor_mask3, float ignore_thre,
                    double *loss4, float *gobj, int *tcell_all, int *mcell, float *mgrad, int *status,
                    yl_stream_t stream);
/* The same call for the SECOND and later layers of one loss evaluation, issued directly behind the previous layer's
 * yl_loss_forward[_chained] on the same stream: its matching kernel is launched programmatically and starts in the tail of
 * the previous layer's objectness kernel.  It reads raw / labels and accumulates into loss4 / status before it waits for
 * that kernel, which is safe only because its predecessor is this library's own kernel; behind anything else (the
 * producer of raw or labels, the fill that zeroes loss4) use yl_loss_forward, a normal, fully ordered launch. */
int yl_loss_forward_chained(const float *raw, const float *labels, int B, int F, int K, int C, int layer_no,
                    const float *anchors_px, const int *anchor_mask3, float ignore_thre,
                    double *loss4, float *gobj, int *tcell_all, int *mcell, float *mgrad, int *status,
                    yl_stream_t stream);
int yl_loss_backward(const float *gobj, const int *mcell, const float *mgrad, const float *upstream,
                     int B, int F, int K, int C, float *grad_raw, yl_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Host-buffer entry (what a non-Python caller binds; also the bench's end-to-end leg): raw head tensors in
 * HOST memory -> detections in HOST memory.  The context owns device staging buffers, workspace, streams and
 * pinned bounce buffers; H2D copies, kernels and the D2H of rows/counts are all inside the call.
 *   raw_host[l] [B, 3*(5+C), F[l], F[l]] fp32 host (pinned or pageable)
 *   out_rows_host [B, cap_out, 7], counts_host [B]
 * Returns YL_OK; counts_host[b] > cap_out means image b was truncated.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct yl_context yl_context;
int yl_context_create(yl_context **ctx, int device, int B, const int *F, int n_layers, int C,
                      const float *anchors_px, const int *anchor_mask, int cap_seg, long cap_out);
int yl_context_destroy(yl_context *ctx);
int yl_detect_host(yl_context *ctx, const float *const *raw_host, float conf_thre, float nms_thre,
                   float *out_rows_host, int *counts_host);

/* ---------------------------------------------------------------------------------------------------------
 * (e) Multi-GPU: the final exchange of the image-sharded path            (no reference counterpart: the
 *     reference validates on rank 0 only, yolo/engine/build.py:110-190, main_amp.py:175,207)
 * One process per GPU; rank r owns B images and ends up with the detections of all world*B images (rank-major).  Each rank
 * owns one device window (rows / counts / flags) that the peers map through CUDA IPC over NVLink; yl_xchg_push copies
 * exactly counts[b] rows per image into every rank's window with plain peer stores (no collective, no staging, no host
 * synchronisation), yl_xchg_wait blocks the stream until every rank's rows of that slot have landed, yl_xchg_release
 * hands the slot back (a peer's next push into it waits for that).  All three are graph-capturable, so the exchange of
 * step i runs under the kernels of step i+1 (double-buffered `slots`).
 *   create   cap_out % 4 == 0 (row runs stay 16-byte aligned); slots in [1, 8]
 *   handle   yl_xchg_handle_bytes() opaque bytes of this rank's window, to be all-gathered by the host (torch.distributed)
 *   connect  handles = world * yl_xchg_handle_bytes() bytes, rank-major
 *   push     rows [B, cap_out, 7] fp32 and counts [B] i32 as written by yl_nms (device, rows 16-byte aligned)
 *   rows/counts  device pointers to THIS rank's gathered [world*B, cap_out, 7] / [world*B] of a slot
 *   status   device int32: 0 ok, 1 = a push timed out waiting for a reader, 2 = a wait timed out waiting for a source
 * --------------------------------------------------------------------------------------------------------- */
typedef struct yl_xchg yl_xchg;
int yl_xchg_create(yl_xchg **out, int device, int rank, int world, int B, long cap_out, int slots);
int yl_xchg_destroy(yl_xchg *x);
/* The same object over windows the CALLER owns: `windows[p]` = rank p's window as mapped into this process (all of
 * yl_xchg_window_bytes() bytes, 256-byte aligned; e.g. torch symmetric memory), `multicast` = an NVSwitch multicast mapping of the
 * same windows or NULL.  With a multicast mapping yl_xchg_push issues ONE multimem.st per 16 bytes and the switch replicates it into
 * every window (a rank sends 1/world of the peer-store forms' bytes).  No CUDA IPC; yl_xchg_destroy leaves the windows alone.
 * The caller synchronises the ranks between creation and the first push. */
size_t yl_xchg_window_bytes(int world, int B, long cap_out, int slots);
int yl_xchg_create_external(yl_xchg **out, int device, int rank, int world, int B, long cap_out, int slots,
                            void *const *windows, void *multicast);
size_t yl_xchg_handle_bytes(void);
int yl_xchg_local_handle(yl_xchg *x, void *handle);
int yl_xchg_connect(yl_xchg *x, const void *handles);
int yl_xchg_push(yl_xchg *x, const float *rows, const int *counts, int slot, yl_stream_t stream);
int yl_xchg_wait(yl_xchg *x, int slot, yl_stream_t stream);
int yl_xchg_release(yl_xchg *x, int slot, yl_stream_t stream);
void *yl_xchg_rows(yl_xchg *x, int slot);
void *yl_xchg_counts(yl_xchg *x, int slot);
void *yl_xchg_status(yl_xchg *x);

#ifdef __cplusplus
}
#endif
#endif /* YOLO_HEAD_H_ */
