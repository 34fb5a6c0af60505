// yl_target.cu -- YOLOLoss.build_target (yolo/model/yololoss.py:118-371) and the bboxes_iou it calls (:16-91).
//
//   k_target_objmask   per cell: obj_mask = !(max_n IoU(pred, GT_n) > ignore_thresh)          (:276-294)
//   k_target_scatter   per image: GT <-> 9-anchor IoU argmax (:249-265) and the assignment of
//                      obj_mask / tgt_mask / tgt_scale / target at the matched cells (:304-369),
//                      resolving same-cell collisions the way the reference's sequential loop does:
//                      scalar fields last-writer-wins, class one-hots accumulate.
// The dense zero background of target / tgt_mask / tgt_scale (:156-167, 99.9 % of the 16 MB/image this path writes) is
// stored by k_target_objmask itself, 128 bits at a time, before it computes its cells' IoUs: the pass runs at the pace
// of the HBM writes and the IoU arithmetic hides underneath instead of following three memsets.
#include "yl_common.cuh"
#include "../../include/yolo_head.h"

namespace yl {

constexpr int TG_THREADS = 256;
constexpr int TG_MAXK = 256;       // labels per image held in shared memory (cfg DATA.MAX_NUM_LABELS is 60)

// Loads image b's labels, counts nlabel = #(row sum > 0) (:219) and converts the FIRST n rows to grid units (:196-202).
__device__ int load_truth(const float *__restrict__ labels, int b, int K, float stride, float (*tb)[4], float *cls, int *sh_n)
{
    if (threadIdx.x == 0) *sh_n = 0;
    __syncthreads();
    const float *lab = labels + (size_t)b * K * 5;
    for (int t = threadIdx.x; t < K; t += blockDim.x) {
        const float l0 = lab[5 * t], l1 = lab[5 * t + 1], l2 = lab[5 * t + 2], l3 = lab[5 * t + 3], l4 = lab[5 * t + 4];
        const float s = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(l0, l1), l2), l3), l4);
        if (s > 0.0f) atomicAdd(sh_n, 1);
        tb[t][0] = __fdiv_rn(l0, stride); tb[t][1] = __fdiv_rn(l1, stride);
        tb[t][2] = __fdiv_rn(l2, stride); tb[t][3] = __fdiv_rn(l3, stride);
        cls[t] = l4;
    }
    __syncthreads();
    return *sh_n;
}

// bboxes_iou(..., xyxy=False) for one pair, literal NaN-propagating form (:64-91)
__device__ __forceinline__ float iou_xywh(float ax, float ay, float aw, float ah, float bx, float by, float bw, float bh)
{
    const float ahw = __fmul_rn(aw, 0.5f), ahh = __fmul_rn(ah, 0.5f), bhw = __fmul_rn(bw, 0.5f), bhh = __fmul_rn(bh, 0.5f);
    const float tlx = nanmaxf(__fsub_rn(ax, ahw), __fsub_rn(bx, bhw)), tly = nanmaxf(__fsub_rn(ay, ahh), __fsub_rn(by, bhh));
    const float brx = nanminf(__fadd_rn(ax, ahw), __fadd_rn(bx, bhw)), bry = nanminf(__fadd_rn(ay, ahh), __fadd_rn(by, bhh));
    const float en = __fmul_rn((tlx < brx) ? 1.0f : 0.0f, (tly < bry) ? 1.0f : 0.0f);
    const float ai = __fmul_rn(__fmul_rn(__fsub_rn(brx, tlx), __fsub_rn(bry, tly)), en);
    return __fdiv_rn(ai, __fsub_rn(__fadd_rn(__fmul_rn(aw, ah), __fmul_rn(bw, bh)), ai));
}

// The CTA zeroes n floats at p (4-byte aligned): scalar head up to a 16-byte boundary, float4 body, scalar tail.
__device__ __forceinline__ void cta_zero(float *p, int n)
{
    int head = (int)(((16u - (unsigned)((uintptr_t)p & 15u)) & 15u) >> 2);
    if (head > n) head = n;
    if ((int)threadIdx.x < head) p[threadIdx.x] = 0.0f;
    float4 *p4 = reinterpret_cast<float4 *>(p + head);
    const int n4 = (n - head) >> 2;
#pragma unroll 4
    for (int i = threadIdx.x; i < n4; i += TG_THREADS) p4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int tail = head + 4 * n4 + (int)threadIdx.x;
    if (tail < n) p[tail] = 0.0f;
}

__device__ __forceinline__ bool bounded(float v) { return fabsf(v) <= 1e18f; }   // false for NaN / inf / huge

__global__ void __launch_bounds__(TG_THREADS)
k_target_objmask(const float *__restrict__ pred, long s0, long s1, long s2, long s3, long s4,
                 const float *__restrict__ labels, int F, int K, int C, float stride, float ignore_thre,
                 float *__restrict__ obj_mask, float *__restrict__ target, float *__restrict__ tgt_mask,
                 float *__restrict__ tgt_scale)
{
    __shared__ __align__(16) float tb[TG_MAXK][4];
    __shared__ __align__(16) float tc[TG_MAXK][4];       // GT corners (x1, y1, x2, y2) for the overlap pre-test
    __shared__ float tcls[TG_MAXK];
    __shared__ float tarea[TG_MAXK];
    __shared__ unsigned char tsimple[TG_MAXK];
    __shared__ int sh_n;
    const int b = blockIdx.y;
    const int cells = 3 * F * F;
    {
        // zero background of this CTA's cells (contiguous runs in all three tensors); k_target_scatter follows in stream order
        const int c0 = blockIdx.x * TG_THREADS;
        const int nc = min(TG_THREADS, cells - c0);
        const size_t g0 = (size_t)b * cells + c0;
        cta_zero(target + g0 * (5 + C), nc * (5 + C));
        cta_zero(tgt_mask + g0 * (4 + C), nc * (4 + C));
        cta_zero(tgt_scale + g0 * 2, nc * 2);
    }
    const int n = load_truth(labels, b, K, stride, tb, tcls, &sh_n);
    const int cell = blockIdx.x * TG_THREADS + threadIdx.x;
    if (n == 0) {                                                     // :225-227 obj_mask stays 1
        if (cell < cells) obj_mask[(size_t)b * cells + cell] = 1.0f;
        return;
    }
    for (int t = threadIdx.x; t < n; t += TG_THREADS) {
        tsimple[t] = bounded(tb[t][0]) && bounded(tb[t][1]) && bounded(tb[t][2]) && bounded(tb[t][3]) &&
                     (__fmul_rn(tb[t][2], tb[t][3]) >= 0.0f);    // union = area_a + area_b stays > 0
        const float bhw = __fmul_rn(tb[t][2], 0.5f), bhh = __fmul_rn(tb[t][3], 0.5f);
        tc[t][0] = __fsub_rn(tb[t][0], bhw); tc[t][1] = __fsub_rn(tb[t][1], bhh);
        tc[t][2] = __fadd_rn(tb[t][0], bhw); tc[t][3] = __fadd_rn(tb[t][1], bhh);
        tarea[t] = __fmul_rn(tb[t][2], tb[t][3]);
    }
    __syncthreads();
    if (cell >= cells) return;
    const int a = cell / (F * F);
    const int r = cell - a * F * F;
    const int j = r / F, i = r - j * F;
    const float *pp = pred + (size_t)b * s0 + (size_t)a * s1 + (size_t)j * s2 + (size_t)i * s3;
    const float ax = pp[0], ay = pp[s4], aw = pp[2 * s4], ah = pp[3 * s4];
    // Fast path: with all coordinates finite and bounded and a strictly positive pred area, a GT that does not
    // overlap the cell's box has IoU exactly +0 (no NaN, no overflow), so only overlapping pairs need the division.
    // (and a non-negative threshold: an IoU of exactly 0 must not count as "above")
    const bool csimple = bounded(ax) && bounded(ay) && bounded(aw) && bounded(ah) && (__fmul_rn(aw, ah) > 0.0f) &&
                         (ignore_thre >= 0.0f);
    const float ahw = __fmul_rn(aw, 0.5f), ahh = __fmul_rn(ah, 0.5f);
    const float ax1 = __fsub_rn(ax, ahw), ay1 = __fsub_rn(ay, ahh), ax2 = __fadd_rn(ax, ahw), ay2 = __fadd_rn(ay, ahh);
    const float area_a = __fmul_rn(aw, ah);
    // max_n IoU > thr (:283-286) == "some IoU > thr and no IoU is NaN" (torch.max propagates NaN, and NaN > thr is False)
    bool above = false, nan_seen = false;
    for (int t = 0; t < n; ++t) {
        if (csimple && tsimple[t]) {
            const float4 g = *reinterpret_cast<const float4 *>(tc[t]);
            const float tlx = fmaxf(ax1, g.x), brx = fminf(ax2, g.z);
            const float tly = fmaxf(ay1, g.y), bry = fminf(ay2, g.w);
            if (tlx < brx && tly < bry) {
                // same operations as iou_xywh with en == 1; the quotient is only formed when the comparison is within
                // 0.1 % of the threshold (its rounding error is 6e-8)
                const float ai = __fmul_rn(__fsub_rn(brx, tlx), __fsub_rn(bry, tly));
                const float uni = __fsub_rn(__fadd_rn(area_a, tarea[t]), ai);
                const float tu = __fmul_rn(ignore_thre, uni);
                bool d;
                if (uni > 0.0f && uni < 3.0e38f && ai > 1.001f * tu && ai < 3.0e38f) d = true;
                else if (uni > 0.0f && uni < 3.0e38f && ai < 0.999f * tu) d = false;
                else d = __fdiv_rn(ai, uni) > ignore_thre;
                above |= d;
            }
        } else {
            const float v = iou_xywh(ax, ay, aw, ah, tb[t][0], tb[t][1], tb[t][2], tb[t][3]);
            nan_seen |= (v != v);
            above |= (v > ignore_thre);
        }
    }
    const bool best_above = above && !nan_seen;
    obj_mask[(size_t)b * cells + cell] = best_above ? 0.0f : 1.0f;             // :286-294
}

struct AnchorSet { float w[9], h[9]; int mask[3]; };

__global__ void __launch_bounds__(TG_THREADS)
k_target_scatter(const float *__restrict__ labels, int F, int K, int C, float stride, AnchorSet an,
                 float *__restrict__ target, float *__restrict__ obj_mask, float *__restrict__ tgt_mask,
                 float *__restrict__ tgt_scale, int *__restrict__ status)
{
    __shared__ float tb[TG_MAXK][4];
    __shared__ float tcls[TG_MAXK];
    __shared__ int tcell[TG_MAXK];       // matched cell index within the image, or -1
    __shared__ int tanc[TG_MAXK];        // anchor slot a = best_n % 3
    __shared__ int sh_n;
    const int b = blockIdx.x;
    const int n = load_truth(labels, b, K, stride, tb, tcls, &sh_n);
    if (n == 0) return;
    const int nch = 5 + C;
    const int cells = 3 * F * F;
    for (int t = threadIdx.x; t < n; t += TG_THREADS) {
        // :249-254 IoU of (0,0,w,h) against the nine (0,0,aw,ah) reference anchors, xyxy=True form; first max wins,
        // NaN counts as the maximum (torch.argmax)
        const float w = tb[t][2], h = tb[t][3];
        const float area_a = __fmul_rn(__fsub_rn(w, 0.0f), __fsub_rn(h, 0.0f));
        int best_n = 0;
        float best = 0.0f;
        for (int q = 0; q < 9; ++q) {
            const float tlx = nanmaxf(0.0f, 0.0f), tly = tlx;
            const float brx = nanminf(w, an.w[q]), bry = nanminf(h, an.h[q]);
            const float en = __fmul_rn((tlx < brx) ? 1.0f : 0.0f, (tly < bry) ? 1.0f : 0.0f);
            const float ai = __fmul_rn(__fmul_rn(__fsub_rn(brx, tlx), __fsub_rn(bry, tly)), en);
            const float area_b = __fmul_rn(__fsub_rn(an.w[q], 0.0f), __fsub_rn(an.h[q], 0.0f));
            const float v = __fdiv_rn(ai, __fsub_rn(__fadd_rn(area_a, area_b), ai));
            if (q == 0) { best = v; best_n = 0; }
            else if (best == best && (v != v || v > best)) { best = v; best_n = q; }
        }
        int cell = -1;
        if (best_n == an.mask[0] || best_n == an.mask[1] || best_n == an.mask[2]) {          // :264-265
            int i = (int)(short)(int)tb[t][0];                                                // :206-207 int16 truncation
            int j = (int)(short)(int)tb[t][1];
            if (i < 0) i += F;                                                                // python negative indexing
            if (j < 0) j += F;
            if (i < 0 || i >= F || j < 0 || j >= F) { if (status) atomicExch(status, 1); }
            else cell = ((best_n % 3) * F + j) * F + i;                                       // :257
        }
        tcell[t] = cell;
        tanc[t] = best_n % 3;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = warp; t < n; t += TG_THREADS / 32) {
        const int cell = tcell[t];
        if (cell < 0) continue;
        const size_t gc = (size_t)b * cells + cell;
        bool overridden = false;                                       // a later matched GT lands on the same cell
        for (int u = t + 1 + lane; u < n; u += 32) overridden |= (tcell[u] == cell);
        overridden = __any_sync(0xFFFFFFFFu, overridden);
        for (int k = lane; k < 4 + C; k += 32) tgt_mask[gc * (4 + C) + k] = 1.0f;             // :333
        if (lane == 0) {
            obj_mask[gc] = 1.0f;                                                              // :330
            float *tg = target + gc * nch;
            if (!overridden) {
                const float tw = tb[t][2], th = tb[t][3];
                const float sc = __fsqrt_rn(__fsub_rn(2.0f, __fdiv_rn(__fdiv_rn(__fmul_rn(tw, th), (float)F), (float)F)));   // :337
                tgt_scale[gc * 2] = sc; tgt_scale[gc * 2 + 1] = sc;
                tg[0] = __fsub_rn(tb[t][0], (float)(short)(int)tb[t][0]);                     // :346
                tg[1] = __fsub_rn(tb[t][1], (float)(short)(int)tb[t][1]);                     // :349
                const int a = tanc[t];
                tg[2] = spec_logf(__fadd_rn(__fdiv_rn(tw, an.w[an.mask[a]]), 1e-16f));        // :362
                tg[3] = spec_logf(__fadd_rn(__fdiv_rn(th, an.h[an.mask[a]]), 1e-16f));        // :365
            }
            tg[4] = 1.0f;                                                                     // :367
            int kc = 5 + (int)(short)(int)tcls[t];                                            // :369
            if (kc < 0) kc += nch;
            if (kc < 0 || kc >= nch) { if (status) atomicExch(status, 1); }
            else tg[kc] = 1.0f;
        }
    }
}

}  // namespace yl

using namespace yl;

extern "C" int yl_build_target(const float *pred, const long *ps, const float *labels, int B, int F, int K,
                               int C, int layer_no, const float *anchors_px, const int *anchor_mask3, float ignore_thre,
                               float *target, float *obj_mask, float *tgt_mask, float *tgt_scale, int *status,
                               yl_stream_t stream)
{
    if (!pred || !ps || !labels || !anchors_px || !anchor_mask3 || !target || !obj_mask || !tgt_mask || !tgt_scale)
        return YL_ERR_ARG;
    if (B <= 0 || F <= 0 || K <= 0 || K > TG_MAXK || C <= 0 || layer_no < 0 || layer_no > 2) return YL_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const float stride = (float)(8 << layer_no);                                              // yololoss.py:99,136
    if (status) YL_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int), st));
    AnchorSet an;
    for (int q = 0; q < 9; ++q) {                                                             // :139-150
        an.w[q] = (float)((double)anchors_px[2 * q] / (double)stride);
        an.h[q] = (float)((double)anchors_px[2 * q + 1] / (double)stride);
    }
    for (int a = 0; a < 3; ++a) {
        if (anchor_mask3[a] < 0 || anchor_mask3[a] > 8) return YL_ERR_ARG;
        an.mask[a] = anchor_mask3[a];
    }
    dim3 grid((3 * F * F + TG_THREADS - 1) / TG_THREADS, B);
    k_target_objmask<<<grid, TG_THREADS, 0, st>>>(pred, ps[0], ps[1], ps[2], ps[3], ps[4], labels, F, K, C, stride, ignore_thre,
                                                  obj_mask, target, tgt_mask, tgt_scale);
    YL_LAUNCH_CHECK();
    k_target_scatter<<<B, TG_THREADS, 0, st>>>(labels, F, K, C, stride, an, target, obj_mask, tgt_mask, tgt_scale, status);
    YL_LAUNCH_CHECK();
    return YL_OK;
}
