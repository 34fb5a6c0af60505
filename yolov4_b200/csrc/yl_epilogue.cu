// yl_epilogue.cu -- SURVEY.md 8(f) row N1: the per-detection epilogue that follows postprocess() in the reference's callers.
//
//   mode 0  validate():  yolo/engine/build.py:146-164 + yolobox2xywh (yolo/util/utils.py:281-309)
//           row = (image_id, category_id = class_ids[int(cls)], x, y, w, h, score = obj_conf * cls_conf), all in float64
//           exactly as the Python loop computes them (float(fp32) -> double, then (v / dst) * src in double).
//   mode 1  detect.parse_info(): detect.py:171-179 + yolobox2yxyx (utils.py:312-340)
//           row = (image_id, category_id, y1, x1, y2, x2, cls_conf), box = (v * src) / dst in the row's own fp32 arithmetic
//           (NumPy float32 scalars times Python numbers stay float32 under NEP 50).
// One thread per detection row; 28 B in, 56 B out.  Replaces a Python loop with ~10 .item() calls per detection.
#include "yl_common.cuh"
#include "../../include/yolo_head.h"

namespace yl {

__device__ __forceinline__ void coco_row(const float *p, int b, const double *__restrict__ img_info,
                                         const long long *__restrict__ image_ids, const int *__restrict__ class_ids, int n_classes,
                                         int mode, double *o)
{
    const float x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3], obj = p[4], cls_conf = p[5];
    const int cls = (int)p[6];
    const double src_h = img_info[4 * b + 0], src_w = img_info[4 * b + 1], dst_h = img_info[4 * b + 2], dst_w = img_info[4 * b + 3];
    o[0] = (double)image_ids[b];
    o[1] = (cls >= 0 && cls < n_classes) ? (double)class_ids[cls] : -1.0;
    if (mode == 0) {
        const double dx1 = (double)x1, dy1 = (double)y1, dx2 = (double)x2, dy2 = (double)y2;
        o[2] = __dmul_rn(__ddiv_rn(dx1, dst_w), src_w);                                   // x1 / dst_w * src_w
        o[3] = __dmul_rn(__ddiv_rn(dy1, dst_h), src_h);
        o[4] = __dmul_rn(__ddiv_rn(__dsub_rn(dx2, dx1), dst_w), src_w);                   // (x2 - x1) / dst_w * src_w
        o[5] = __dmul_rn(__ddiv_rn(__dsub_rn(dy2, dy1), dst_h), src_h);
        o[6] = __dmul_rn((double)obj, (double)cls_conf);                                  // build.py:156
    } else {
        const float fsh = (float)src_h, fsw = (float)src_w, fdh = (float)dst_h, fdw = (float)dst_w;
        o[2] = (double)__fdiv_rn(__fmul_rn(y1, fsh), fdh);                                // y1 * src_h / dst_h
        o[3] = (double)__fdiv_rn(__fmul_rn(x1, fsw), fdw);
        o[4] = (double)__fdiv_rn(__fmul_rn(y2, fsh), fdh);
        o[5] = (double)__fdiv_rn(__fmul_rn(x2, fsw), fdw);
        o[6] = (double)cls_conf;
    }
}

__global__ void __launch_bounds__(256)
k_coco_rows(const float *__restrict__ rows, const int *__restrict__ row_image, long K, const double *__restrict__ img_info,
            const long long *__restrict__ image_ids, const int *__restrict__ class_ids, int n_classes, int mode,
            double *__restrict__ out)
{
    const long r = (long)blockIdx.x * 256 + threadIdx.x;
    if (r >= K) return;
    coco_row(rows + r * 7, row_image[r], img_info, image_ids, class_ids, n_classes, mode, out + r * 7);
}

// The padded form the device path produces (yl_nms: rows [B, cap_out, 7] + counts [B]): grid (row blocks, image); the output is
// compact, image b's rows start at the exclusive prefix of the counts (every CTA of an image adds up the b lower counts itself).
__global__ void __launch_bounds__(256)
k_coco_rows_padded(const float *__restrict__ rows, const int *__restrict__ counts, long cap_out, const double *__restrict__ img_info,
                   const long long *__restrict__ image_ids, const int *__restrict__ class_ids, int n_classes, int mode,
                   double *__restrict__ out)
{
    __shared__ long sh_pre[8];
    const int b = blockIdx.y;
    const long n = min((long)max(counts[b], 0), cap_out);
    if ((long)blockIdx.x * 256 >= n) return;                                              // CTA-uniform
    long pre = 0;
    for (int i = threadIdx.x; i < b; i += 256) pre += min((long)max(counts[i], 0), cap_out);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(0xFFFFFFFFu, pre, o);
    if ((threadIdx.x & 31) == 0) sh_pre[threadIdx.x >> 5] = pre;
    __syncthreads();
    pre = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) pre += sh_pre[w];
    const long r = (long)blockIdx.x * 256 + threadIdx.x;
    if (r >= n) return;
    coco_row(rows + ((size_t)b * cap_out + r) * 7, b, img_info, image_ids, class_ids, n_classes, mode, out + (pre + r) * 7);
}

}  // namespace yl

extern "C" int yl_coco_rows(const float *rows, const int *row_image, long K, const double *img_info, const long long *image_ids,
                            const int *class_ids, int n_classes, int mode, double *out, yl_stream_t stream)
{
    if (K < 0 || n_classes <= 0 || mode < 0 || mode > 1) return YL_ERR_ARG;
    if (K == 0) return YL_OK;
    if (!rows || !row_image || !img_info || !image_ids || !class_ids || !out) return YL_ERR_ARG;
    yl::k_coco_rows<<<(unsigned)((K + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, row_image, K, img_info, image_ids, class_ids,
                                                                               n_classes, mode, out);
    YL_LAUNCH_CHECK();
    return YL_OK;
}

extern "C" int yl_coco_rows_padded(const float *rows, const int *counts, int B, long cap_out, const double *img_info,
                                   const long long *image_ids, const int *class_ids, int n_classes, int mode, double *out,
                                   yl_stream_t stream)
{
    if (B <= 0 || cap_out <= 0 || n_classes <= 0 || mode < 0 || mode > 1) return YL_ERR_ARG;
    if (!rows || !counts || !img_info || !image_ids || !class_ids || !out) return YL_ERR_ARG;
    if (B > 65535) return YL_ERR_ARG;
    dim3 grid((unsigned)((cap_out + 255) / 256), (unsigned)B);
    yl::k_coco_rows_padded<<<grid, 256, 0, (cudaStream_t)stream>>>(rows, counts, cap_out, img_info, image_ids, class_ids, n_classes, mode,
                                                                  out);
    YL_LAUNCH_CHECK();
    return YL_OK;
}
