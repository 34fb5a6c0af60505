// yl_loss.cu -- N2: YOLOLoss.forward (yolo/model/yololoss.py:373-443) for one layer, fused with the
// train-mode YOLOLayer (yololayer.py:122-145) and build_target (:118-371), straight from the raw head tensor.
//
// The reference materialises, per layer and step, the dense `output`, `pred`, `target`, `tgt_mask`, `tgt_scale` and `obj_mask`
// tensors and runs ~15 elementwise passes plus four reductions over them.  But outside the <= K matched cells of an
// image every masked term is exactly 0 (tgt_mask = 0 => BCE(0, 0) = 0, MSE(0, 0) = 0), and the objectness term of a cell
// needs only its five box / objectness logits.  So:
//
//   k_loss_match   per image: GT -> cell assignment (:249-265), then for every matched cell (last writer per cell, class
//                  one-hots accumulated over colliding GTs) the xy / wh / cls terms of :417-427 and their
//                  derivatives with respect to the raw logits, kept in a compact per-GT table.
//   k_loss_obj     per cell: decode the box from tx,ty,tw,th (5 of the 85 planes are read), ignore mask (:276-294),
//                  objectness BCE (:425) and d/d(raw objectness).
//   k_loss_bwd_*   grad_raw = upstream * (objectness plane from k_loss_obj, matched cells from the table); the rest of
//                  grad_raw is the zero background (cudaMemsetAsync).
//
// Per-term arithmetic is fp32 like the reference's (ATen binary_cross_entropy: (t-1) max(log(1-x), -100) - t max(log x, -100),
// backward (x-t)/max((1-x) x, 1e-12); mse_loss); sums are accumulated in fp64.  Parity: loss and gradient within 1e-5 /
// 2e-5 relative of the reference (tests/golden/loss.npz), not bit-exact (the reference's reduction order is unspecified).
#include "yl_target.cuh"
#include "../../include/yolo_head.h"

namespace yl {

constexpr int LS_THREADS = 256;

__device__ __forceinline__ float bce_term(float x, float t)
{
    const float lx = fmaxf(spec_logf(x), -100.0f), l1x = fmaxf(spec_logf(__fsub_rn(1.0f, x)), -100.0f);
    return __fsub_rn(__fmul_rn(__fsub_rn(t, 1.0f), l1x), __fmul_rn(t, lx));
}
// bce_term for t = 1 (one) or t = 0: the other logarithm is clamped to [-100, 0] and multiplied by zero, so the term is
// -max(log x, -100) resp. -max(log(1 - x), -100) -- the same value (NaN included) for one spec_logf instead of two.
__device__ __forceinline__ float bce_term01(float x, bool one)
{
    const float l = fmaxf(spec_logf(one ? x : __fsub_rn(1.0f, x)), -100.0f);
    return __fsub_rn(0.0f, l);
}
__device__ __forceinline__ float bce_grad(float x, float t)
{
    return __fdiv_rn(__fsub_rn(x, t), fmaxf(__fmul_rn(__fsub_rn(1.0f, x), x), 1e-12f));
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

struct LossAnchors { AnchorSet an; float maw[3], mah[3]; };          // all nine / the layer's three anchors, grid units

__global__ void __launch_bounds__(LS_THREADS)
k_loss_match(const float *__restrict__ raw, const float *__restrict__ labels, int F, int K, int C, float stride, LossAnchors A,
             double *__restrict__ loss4, int *__restrict__ tcell_all, int *__restrict__ mcell, float *__restrict__ mgrad,
             int *__restrict__ status)
{
    __shared__ float tb[TG_MAXK][4];
    __shared__ float tcls[TG_MAXK];
    __shared__ int tcell[TG_MAXK];
    __shared__ int tanc[TG_MAXK];
    __shared__ int sh_n;
    pdl_trigger();
    // grid (image, slice): every CTA of an image repeats the cheap assignment, then its warps take the GTs of its slice
    const int b = blockIdx.x;
    const int n = load_truth(labels, b, K, stride, tb, tcls, &sh_n);
    for (int t = threadIdx.x; t < K; t += LS_THREADS) {
        int cell = -1, anc = 0;
        if (t < n) cell = match_truth(tb[t], F, A.an, &anc, status);
        tcell[t] = cell;
        tanc[t] = anc;
        if (blockIdx.y == 0) tcell_all[(size_t)b * K + t] = cell;
        // the slice that owns GT slot t (see the loop below) initialises its entry of the table, before its own barrier
        if ((t % ((int)gridDim.y * (LS_THREADS / 32))) / (LS_THREADS / 32) == (int)blockIdx.y) mcell[(size_t)b * K + t] = -1;
    }
    __syncthreads();
    const int nch = 5 + C, F2 = F * F, nk = 4 + C;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc_xy = 0.0, acc_wh = 0.0, acc_cls = 0.0;
    for (int t = blockIdx.y * (LS_THREADS / 32) + warp; t < n; t += gridDim.y * (LS_THREADS / 32)) {
        const int cell = tcell[t];
        if (cell < 0) continue;
        bool overridden = false;                                       // a later matched GT lands on the same cell: it owns the cell
        for (int u = t + 1 + lane; u < n; u += 32) overridden |= (tcell[u] == cell);
        if (__any_sync(0xFFFFFFFFu, overridden)) continue;
        const int a = cell / F2, r = cell - a * F2;
        const float *cp = raw + ((size_t)(b * 3 + a) * nch) * F2 + r;   // channel k of this cell at cp[k * F2]
        const float tw = tb[t][2], th = tb[t][3];
        const float sc = __fsqrt_rn(__fsub_rn(2.0f, __fdiv_rn(__fdiv_rn(__fmul_rn(tw, th), (float)F), (float)F)));   // :337
        const float w2 = __fmul_rn(sc, sc);                                                                          // :417
        float *gout = mgrad + ((size_t)b * K + t) * nk;
        if (lane == 0) mcell[(size_t)b * K + t] = cell;
        // GTs that share the cell (usually just this one): bit u of same[u >> 5]
        unsigned same[TG_MAXK / 32];
#pragma unroll
        for (int w = 0; w < TG_MAXK / 32; ++w) {
            const int u = 32 * w + lane;
            same[w] = (32 * w < n) ? __ballot_sync(0xFFFFFFFFu, u < n && tcell[u] == cell) : 0u;
        }
        for (int kk = lane; kk < nk; kk += 32) {
            const int k = kk < 4 ? kk : kk + 1;
            const float v = cp[(size_t)k * F2];
            float g;
            if (kk < 2) {                                              // xy: weighted BCE (:421)
                const float x = spec_sigmoidf(v);
                const float tg = __fsub_rn(tb[t][kk], (float)(short)(int)tb[t][kk]);                                  // :346,349
                acc_xy += (double)__fmul_rn(bce_term(x, tg), w2);
                g = __fmul_rn(__fmul_rn(__fmul_rn(bce_grad(x, tg), w2), x), __fsub_rn(1.0f, x));
            } else if (kk < 4) {                                       // wh: MSE / 2 on the scaled raw logits (:423)
                const float anc = (kk == 2) ? A.maw[tanc[t]] : A.mah[tanc[t]];
                const float tg = spec_logf(__fadd_rn(__fdiv_rn(kk == 2 ? tw : th, anc), 1e-16f));                     // :362,365
                const float d = __fsub_rn(__fmul_rn(v, sc), __fmul_rn(tg, sc));
                acc_wh += 0.5 * (double)d * (double)d;
                g = __fmul_rn(d, sc);
            } else {                                                   // classes: BCE against the union of the cell's one-hots (:427)
                const int kc = kk - 4;
                float tg = 0.0f;
#pragma unroll
                for (int w = 0; w < TG_MAXK / 32; ++w) {
                    unsigned m = same[w];
                    while (m) {
                        const int u = 32 * w + __ffs(m) - 1;
                        m &= m - 1u;
                        const int cu = (int)(short)(int)tcls[u];                                                      // :369
                        if (cu == kc) tg = 1.0f;
                        if ((cu < 0 || cu >= C) && status) atomicExch(status, 1);   // a class id that would index outside 5..5+C
                    }
                }
                const float x = spec_sigmoidf(v);
                acc_cls += (double)bce_term(x, tg);
                g = __fmul_rn(__fmul_rn(bce_grad(x, tg), x), __fsub_rn(1.0f, x));
            }
            gout[kk] = g;
        }
    }
    acc_xy = warp_sum(acc_xy); acc_wh = warp_sum(acc_wh); acc_cls = warp_sum(acc_cls);
    if (lane == 0) {
        if (acc_xy != 0.0) atomicAdd(&loss4[0], acc_xy);
        if (acc_wh != 0.0) atomicAdd(&loss4[1], acc_wh);
        if (acc_cls != 0.0) atomicAdd(&loss4[3], acc_cls);
    }
    // Nothing here reads the output of the kernel before it on the stream (the previous scale's k_loss_obj), so the
    // two overlap; waiting at the end keeps completion ordered along the chain for whatever follows the last kernel.
    pdl_wait();
}

__global__ void __launch_bounds__(LS_THREADS)
k_loss_obj(const float *__restrict__ raw, const float *__restrict__ labels, int F, int K, int C, float stride, LossAnchors A,
           float ignore_thre, const int *__restrict__ tcell_all, double *__restrict__ loss4, float *__restrict__ gobj)
{
    __shared__ __align__(16) float tb[TG_MAXK][4];
    __shared__ __align__(16) float tc[TG_MAXK][4];
    __shared__ float tcls[TG_MAXK];
    __shared__ float tarea[TG_MAXK];
    __shared__ unsigned char tns[TG_MAXK];
    __shared__ float tkey[TG_MAXK];
    __shared__ int sh_n, sh_ns;
    __shared__ double sh_acc[LS_THREADS / 32];
    __shared__ unsigned sh_hit[LS_THREADS / 32];
    const int b = blockIdx.y;
    const int F2 = F * F, cells = 3 * F2, nch = 5 + C;
    pdl_trigger();
    const int n = load_truth(labels, b, K, stride, tb, tcls, &sh_n);
    if (threadIdx.x < LS_THREADS / 32) sh_hit[threadIdx.x] = 0u;
    if (n > 0) prep_truth(n, tb, tc, tarea, tns, tkey, &sh_ns);
    const int cell = blockIdx.x * LS_THREADS + threadIdx.x;
    double term = 0.0;
    float t4 = 0.0f;
    bool ignored = false;
    if (cell < cells) {
        const int a = cell / F2, r = cell - a * F2;
        const int j = r / F, i = r - j * F;
        const float *cp = raw + ((size_t)(b * 3 + a) * nch) * F2 + r;
        const float t0 = cp[0], t1 = cp[(size_t)F2], t2 = cp[2 * (size_t)F2], t3 = cp[3 * (size_t)F2];
        t4 = cp[4 * (size_t)F2];
        if (n > 0) {                                                   // :225-227: no labels => obj_mask stays 1
            // pred exactly as the train-mode YOLOLayer forms it (yololayer.py:126-134): grid units, no stride
            float sx, sy, ew, eh;
            spec_sigmoid2(t0, t1, sx, sy);                             // packed fp32x2 forms: the same bits as the scalar calls
            spec_exp2(t2, t3, ew, eh);
            const float ax = __fadd_rn(sx, (float)i), ay = __fadd_rn(sy, (float)j);
            const float aw = __fmul_rn(ew, A.maw[a]), ah = __fmul_rn(eh, A.mah[a]);
            ignored = iou_max_above(ax, ay, aw, ah, n, tb, tc, tarea, tns, sh_ns, ignore_thre);
        }
    }
    // Everything above is independent of k_loss_match (launched just before this kernel with programmatic dependent
    // launch): only the matched cells need its output.  Bitmap of the matched cells among this CTA's LS_THREADS cells.
    __syncthreads();
    pdl_wait();
    for (int t = threadIdx.x; t < n; t += LS_THREADS) {
        const int rel = tcell_all[(size_t)b * K + t] - (int)blockIdx.x * LS_THREADS;
        if (rel >= 0 && rel < LS_THREADS) atomicOr(&sh_hit[rel >> 5], 1u << (rel & 31));
    }
    __syncthreads();
    if (cell < cells) {
        const bool matched = (sh_hit[threadIdx.x >> 5] >> (threadIdx.x & 31)) & 1u;
        if (matched) ignored = false;                                  // matched cells are never ignored (:330)
        const float x = spec_sigmoidf(t4);
        float g = 0.0f;
        if (!ignored) {                                                // obj_mask = 1: BCE(x, t) with t = 1 on matched cells (:330,367,425)
            const float tg = matched ? 1.0f : 0.0f;
            term = (double)bce_term01(x, matched);
            g = __fmul_rn(__fmul_rn(bce_grad(x, tg), x), __fsub_rn(1.0f, x));
        }
        gobj[(size_t)b * cells + cell] = g;
    }
    term = warp_sum(term);
    if ((threadIdx.x & 31) == 0) sh_acc[threadIdx.x >> 5] = term;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < LS_THREADS / 32; ++w) s += sh_acc[w];
        if (s != 0.0) atomicAdd(&loss4[2], s);
    }
}

__global__ void __launch_bounds__(LS_THREADS)
k_loss_bwd_obj(const float *__restrict__ gobj, const float *__restrict__ upstream, int F2, int C, long n_cells,
               float *__restrict__ grad_raw)
{
    const long c = (long)blockIdx.x * LS_THREADS + threadIdx.x;       // cell index over [B*3, F2]
    if (c >= n_cells) return;
    const long ba = c / F2;
    const int r = (int)(c - ba * F2);
    grad_raw[((size_t)ba * (5 + C) + 4) * F2 + r] = __fmul_rn(gobj[c], upstream[0]);
}

__global__ void __launch_bounds__(LS_THREADS)
k_loss_bwd_sparse(const int *__restrict__ mcell, const float *__restrict__ mgrad, const float *__restrict__ upstream,
                  int F2, int K, int C, int n_gt, float *__restrict__ grad_raw)
{
    const int gt = blockIdx.x * (LS_THREADS / 32) + (threadIdx.x >> 5);   // one warp per (image, GT slot)
    if (gt >= n_gt) return;
    const int cell = mcell[gt];
    if (cell < 0) return;
    const int lane = threadIdx.x & 31, b = gt / K, nch = 5 + C, nk = 4 + C;
    const int a = cell / F2, r = cell - a * F2;
    float *gp = grad_raw + ((size_t)(b * 3 + a) * nch) * F2 + r;
    const float up = upstream[0];
    for (int kk = lane; kk < nk; kk += 32) gp[(size_t)(kk < 4 ? kk : kk + 1) * F2] = __fmul_rn(mgrad[(size_t)gt * nk + kk], up);
}

static int fill_anchors(LossAnchors &A, const float *anchors_px, const int *anchor_mask3, float stride)
{
    for (int q = 0; q < 9; ++q) {                                                             // yololoss.py:139-150
        A.an.w[q] = (float)((double)anchors_px[2 * q] / (double)stride);
        A.an.h[q] = (float)((double)anchors_px[2 * q + 1] / (double)stride);
    }
    for (int a = 0; a < 3; ++a) {
        if (anchor_mask3[a] < 0 || anchor_mask3[a] > 8) return YL_ERR_ARG;
        A.an.mask[a] = anchor_mask3[a];
        A.maw[a] = A.an.w[anchor_mask3[a]];
        A.mah[a] = A.an.h[anchor_mask3[a]];
    }
    return YL_OK;
}

}  // namespace yl

using namespace yl;

static int loss_forward_impl(const float *raw, const float *labels, int B, int F, int K, int C, int layer_no,
                             const float *anchors_px, const int *anchor_mask3, float ignore_thre,
                             double *loss4, float *gobj, int *tcell_all, int *mcell, float *mgrad, int *status,
                             int chained, yl_stream_t stream)
{
    if (!raw || !labels || !anchors_px || !anchor_mask3 || !loss4 || !gobj || !tcell_all || !mcell || !mgrad) return YL_ERR_ARG;
    if (B <= 0 || F <= 0 || K <= 0 || K > TG_MAXK || C <= 0 || layer_no < 0 || layer_no > 2) return YL_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const float stride = (float)(8 << layer_no);
    LossAnchors A;
    const int rc = fill_anchors(A, anchors_px, anchor_mask3, stride);
    if (rc != YL_OK) return rc;
    // programmatic dependent launch along match(l) -> obj(l) -> match(l+1) -> ...: the latency-bound match kernels and the
    // small scales' obj kernels fill the tails of their neighbours (see the waits in the kernels).  k_loss_match reads raw /
    // labels and adds into loss4 / status BEFORE it waits, so it may only be launched programmatically behind this library's
    // own k_loss_obj (chained != 0): behind a foreign kernel (the producer of raw or labels, the fill of loss4) it is a
    // normal launch, fully ordered after everything earlier on the stream.
    const bool pdl = pdl_enabled();
    YL_CUDA_TRY(launch_after(k_loss_match, dim3(B, 8), dim3(LS_THREADS), 0, st, pdl && chained != 0, raw, labels, F, K, C, stride, A,
                             loss4, tcell_all, mcell, mgrad, status));
    dim3 grid((3 * F * F + LS_THREADS - 1) / LS_THREADS, B);
    YL_CUDA_TRY(launch_after(k_loss_obj, grid, dim3(LS_THREADS), 0, st, pdl, raw, labels, F, K, C, stride, A, ignore_thre,
                             (const int *)tcell_all, loss4, gobj));
    return YL_OK;
}

extern "C" int yl_loss_forward(const float *raw, const float *labels, int B, int F, int K, int C, int layer_no,
                               const float *anchors_px, const int *anchor_mask3, float ignore_thre,
                               double *loss4, float *gobj, int *tcell_all, int *mcell, float *mgrad, int *status,
                               yl_stream_t stream)
{
    return loss_forward_impl(raw, labels, B, F, K, C, layer_no, anchors_px, anchor_mask3, ignore_thre, loss4, gobj, tcell_all,
                             mcell, mgrad, status, 0, stream);
}

extern "C" int yl_loss_forward_chained(const float *raw, const float *labels, int B, int F, int K, int C, int layer_no,
                                       const float *anchors_px, const int *anchor_mask3, float ignore_thre,
                                       double *loss4, float *gobj, int *tcell_all, int *mcell, float *mgrad, int *status,
                                       yl_stream_t stream)
{
    return loss_forward_impl(raw, labels, B, F, K, C, layer_no, anchors_px, anchor_mask3, ignore_thre, loss4, gobj, tcell_all,
                             mcell, mgrad, status, 1, stream);
}

extern "C" int yl_loss_backward(const float *gobj, const int *mcell, const float *mgrad, const float *upstream,
                                int B, int F, int K, int C, float *grad_raw, yl_stream_t stream)
{
    if (!gobj || !mcell || !mgrad || !upstream || !grad_raw || B <= 0 || F <= 0 || K <= 0 || C <= 0) return YL_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int F2 = F * F;
    const long n_cells = (long)B * 3 * F2;
    YL_CUDA_TRY(cudaMemsetAsync(grad_raw, 0, sizeof(float) * (size_t)n_cells * (5 + C), st));
    k_loss_bwd_obj<<<(unsigned)((n_cells + LS_THREADS - 1) / LS_THREADS), LS_THREADS, 0, st>>>(gobj, upstream, F2, C, n_cells, grad_raw);
    YL_LAUNCH_CHECK();
    const int n_gt = B * K;
    k_loss_bwd_sparse<<<(n_gt + LS_THREADS / 32 - 1) / (LS_THREADS / 32), LS_THREADS, 0, st>>>(mcell, mgrad, upstream, F2, K, C, n_gt,
                                                                                             grad_raw);
    YL_LAUNCH_CHECK();
    return YL_OK;
}
