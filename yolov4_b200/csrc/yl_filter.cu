// yl_filter.cu -- front ends of the postprocess pipeline: confidence filter -> candidate segments.
//
//   k_filter_raw    fused YOLOLayer eval decode (yololayer.py:88-166) + xywh->xyxy (utils.py:117-126) + conf filter
//                   with multi-label expansion (utils.py:139-184), single pass over the raw head tensor.
//   k_filter_dense  the same filter over an already decoded [B,M,5+C] tensor (the literal postprocess() input).
//
// Both append one record per surviving (box,class) pair to the (image,class) segment cand[(b*C+c)*cap_seg + slot]
// and store the box once in boxtab/objtab[b*M + row].  Slot order inside a segment is arbitrary (atomics); the NMS
// stage sorts by the unique key (score, row), so final results are deterministic.
#include "yl_common.cuh"
#include "../../include/yolo_head.h"

namespace yl {

constexpr int K1_THREADS = 128;

// Conservative logit-domain bound: every class logit t with fl(obj * sigmoid(t)) >= thr satisfies t >= bound, so
// the streaming loop needs one compare per element and the exact spec-math test runs only on ~1% of them.
//   pass  =>  sigmoid_spec(t) >= (thr/obj)(1 - 2^-23)  =>  sigmoid(t) >= q' := (thr/obj)(1 - 1e-5)   (spec error < 4e-7)
//         =>  t >= logit(min(q', 0.9));  0.01 + 1e-3|L| absorbs the error of the fast __logf / division.
__device__ __forceinline__ float class_logit_bound(float obj, float thr)
{
    if (!(obj >= thr)) return kInf;            // sigmoid <= 1, so fl(obj*cls) <= obj < thr: no class can pass (also NaN obj)
    if (thr <= 0.0f) return -kInf;             // everything non-NaN passes
    float q = fminf((thr / obj) * (1.0f - 1e-5f), 0.9f);
    const float L = __logf(q / (1.0f - q));
    return L - 0.01f - 1e-3f * fabsf(L);
}

template <int VEC, int NW>
__global__ void __launch_bounds__(K1_THREADS)
k_filter_raw(const float *__restrict__ raw, int Fw, int F2, int C, float stride,
             float aw0, float ah0, float aw1, float ah1, float aw2, float ah2,
             int row_off, long M, float thr, int cap_seg, int img_first,
             uint4 *__restrict__ cand, unsigned *__restrict__ seg_count,
             float4 *__restrict__ boxtab, float *__restrict__ objtab)
{
    const int p0 = (blockIdx.x * K1_THREADS + threadIdx.x) * VEC;
    if (p0 >= F2) return;
    const int ba = img_first * 3 + blockIdx.y;
    const int b = ba / 3, a = ba - 3 * b;
    const int nch = 5 + C;
    const float *base = raw + ((size_t)ba * nch) * F2 + p0;

    Vec<VEC> tob;
    tob.load(base + 4 * (size_t)F2);
    float obj[VEC], lth[VEC];
    bool any_alive = false;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        obj[v] = spec_sigmoidf(tob.v[v]);
        lth[v] = class_logit_bound(obj[v], thr);
        any_alive |= (lth[v] != kInf);
    }
    if (!any_alive) return;

    // ---- streaming pass: one compare per class logit, result bits kept in registers -----------------------
    unsigned bits[VEC][NW];
#pragma unroll
    for (int v = 0; v < VEC; ++v)
#pragma unroll
        for (int w = 0; w < NW; ++w) bits[v][w] = 0u;

    const float *cp = base + 5 * (size_t)F2;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const int kn = min(32, C - 32 * w);
#pragma unroll 1
        for (int kk = 0; kk < kn; kk += 8) {
            Vec<VEC> t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (kk + u < kn) t[u].load(cp + (size_t)(32 * w + kk + u) * F2);
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (kk + u < kn) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v)
                        bits[v][w] |= (t[u].v[v] < lth[v]) ? 0u : (1u << (kk + u));   // NaN logits set the bit too
                }
        }
    }

    // ---- exact pass over the flagged pairs (rare) ----------------------------------------------------------
    const float aw = (a == 0) ? aw0 : ((a == 1) ? aw1 : aw2);
    const float ah = (a == 0) ? ah0 : ((a == 1) ? ah1 : ah2);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        if (lth[v] == kInf) continue;
        unsigned any = 0u;
#pragma unroll
        for (int w = 0; w < NW; ++w) any |= bits[v][w];
        if (!any) continue;
        // A: exact test; a NaN class logit makes torch.max (utils.py:139) NaN and drops the whole row (:145)
        bool nan_row = false;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            unsigned m = bits[v][w];
            while (m) {
                const int bit = __ffs(m) - 1;
                m &= m - 1;
                const float t = cp[(size_t)(32 * w + bit) * F2 + v];
                nan_row |= (t != t);
                const float s = __fmul_rn(obj[v], spec_sigmoidf(t));
                if (!(s >= thr)) bits[v][w] &= ~(1u << bit);
            }
        }
        if (nan_row) continue;
        any = 0u;
#pragma unroll
        for (int w = 0; w < NW; ++w) any |= bits[v][w];
        if (!any) continue;
        // B: decode the box once (yololayer.py:150-162) and convert to corners (utils.py:117-126)
        const int p = p0 + v;
        const int gy = p / Fw, gx = p - gy * Fw;
        const float bx = __fmul_rn(__fadd_rn(spec_sigmoidf(base[v]), (float)gx), stride);
        const float by = __fmul_rn(__fadd_rn(spec_sigmoidf(base[(size_t)F2 + v]), (float)gy), stride);
        const float bw = __fmul_rn(__fmul_rn(spec_expf(base[2 * (size_t)F2 + v]), aw), stride);
        const float bh = __fmul_rn(__fmul_rn(spec_expf(base[3 * (size_t)F2 + v]), ah), stride);
        const float hw = __fmul_rn(bw, 0.5f), hh = __fmul_rn(bh, 0.5f);
        const unsigned row = (unsigned)(row_off + a * F2 + p);
        const size_t brow = (size_t)b * M + row;
        boxtab[brow] = make_float4(__fsub_rn(bx, hw), __fsub_rn(by, hh), __fadd_rn(bx, hw), __fadd_rn(by, hh));
        objtab[brow] = obj[v];
        // C: emit one record per surviving (box,class) pair
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            unsigned m = bits[v][w];
            while (m) {
                const int bit = __ffs(m) - 1;
                m &= m - 1;
                const int k = 32 * w + bit;
                const float cls = spec_sigmoidf(cp[(size_t)k * F2 + v]);
                const float s = __fadd_rn(__fmul_rn(obj[v], cls), 0.0f);      // +0 canonicalises -0
                const unsigned seg = (unsigned)(b * C + k);
                const unsigned slot = atomicAdd(&seg_count[seg], 1u);
                if (slot < (unsigned)cap_seg)
                    cand[(size_t)seg * cap_seg + slot] = make_uint4(__float_as_uint(s), row, __float_as_uint(cls), 0u);
            }
        }
    }
}

// One warp per decoded row (5+C contiguous floats, <= 133): coalesced 128-byte reads, ballot-free emission.
constexpr int KD_THREADS = 256;
constexpr int KD_MAXJ = (5 + YL_MAX_CLASSES + 31) / 32;

__global__ void __launch_bounds__(KD_THREADS)
k_filter_dense(const float *__restrict__ pred, long M, int C, int num_classes, float thr, int cap_seg,
               int img_first, long n_rows,
               uint4 *__restrict__ cand, unsigned *__restrict__ seg_count,
               float4 *__restrict__ boxtab, float *__restrict__ objtab)
{
    const int lane = threadIdx.x & 31;
    const long wid = ((long)blockIdx.x * KD_THREADS + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * KD_THREADS) >> 5;
    const int nch = 5 + C;
    const int nj = (nch + 31) >> 5;
    for (long r = wid; r < n_rows; r += nwarps) {
        const int b = img_first + (int)(r / M);
        const unsigned row = (unsigned)(r % M);
        const float *p = pred + ((size_t)b * M + row) * nch;
        float e[KD_MAXJ];
#pragma unroll
        for (int j = 0; j < KD_MAXJ; ++j) {
            const int idx = 32 * j + lane;
            e[j] = (j < nj && idx < nch) ? ldg_stream1(p + idx) : 0.0f;
        }
        const float obj = __shfl_sync(0xFFFFFFFFu, e[0], 4);
        // row pre-filter obj * max_c cls >= thr with torch.max NaN propagation (utils.py:139-148)
        float mx = -kInf;
        bool has_nan = false;
#pragma unroll
        for (int j = 0; j < KD_MAXJ; ++j) {
            const int idx = 32 * j + lane;
            if (j < nj && idx >= 5 && idx < 5 + num_classes) { has_nan |= (e[j] != e[j]); mx = fmaxf(mx, e[j]); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
        has_nan = __any_sync(0xFFFFFFFFu, has_nan);
        if (has_nan || !(__fmul_rn(obj, mx) >= thr)) continue;
        bool pass[KD_MAXJ];
        bool any = false;
#pragma unroll
        for (int j = 0; j < KD_MAXJ; ++j) {
            const int idx = 32 * j + lane;
            pass[j] = (j < nj && idx >= 5 && idx < nch) && (__fmul_rn(e[j], obj) >= thr);    // utils.py:170
            any |= pass[j];
        }
        if (!__any_sync(0xFFFFFFFFu, any)) continue;
        const size_t brow = (size_t)b * M + row;
        const float cx = __shfl_sync(0xFFFFFFFFu, e[0], 0), cy = __shfl_sync(0xFFFFFFFFu, e[0], 1);
        const float w = __shfl_sync(0xFFFFFFFFu, e[0], 2), h = __shfl_sync(0xFFFFFFFFu, e[0], 3);
        if (lane == 0) {
            const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);      // utils.py:117-126
            boxtab[brow] = make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
            objtab[brow] = obj;
        }
#pragma unroll
        for (int j = 0; j < KD_MAXJ; ++j)
            if (pass[j]) {
                const int k = 32 * j + lane - 5;
                const float s = __fadd_rn(__fmul_rn(obj, e[j]), 0.0f);        // nms score = obj*cls (utils.py:209)
                const unsigned seg = (unsigned)(b * C + k);
                const unsigned slot = atomicAdd(&seg_count[seg], 1u);
                if (slot < (unsigned)cap_seg)
                    cand[(size_t)seg * cap_seg + slot] = make_uint4(__float_as_uint(s), row, __float_as_uint(e[j]), 0u);
            }
    }
}

template <int VEC>
static int launch_filter_raw(int NW, dim3 grid, cudaStream_t st, const float *raw, int Fw, int F2, int C, float stride,
                             const float *ag, int row_off, long M, float thr, int cap_seg, int img_first,
                             uint4 *cand, unsigned *seg_count, float4 *boxtab, float *objtab)
{
#define YL_K1_CASE(NW_)                                                                                           \
    case NW_:                                                                                                     \
        k_filter_raw<VEC, NW_><<<grid, K1_THREADS, 0, st>>>(raw, Fw, F2, C, stride, ag[0], ag[1], ag[2], ag[3],  \
                                                            ag[4], ag[5], row_off, M, thr, cap_seg, img_first,  \
                                                            cand, seg_count, boxtab, objtab);                  \
        break;
    switch (NW) {
        YL_K1_CASE(1)
        YL_K1_CASE(2)
        YL_K1_CASE(3)
        YL_K1_CASE(4)
    default:
        return YL_ERR_CLASSES;
    }
#undef YL_K1_CASE
    return YL_OK;
}

}  // namespace yl

using namespace yl;

extern "C" size_t yl_post_workspace_bytes(int B, long M, int C, int cap_seg)
{
    if (B <= 0 || M <= 0 || C <= 0 || cap_seg <= 0) return 0;
    return post_layout(B, M, C, cap_seg).total;
}

extern "C" int yl_post_reset(void *ws, size_t ws_bytes, int B, long M, int C, int cap_seg, yl_stream_t stream)
{
    if (!ws || B <= 0 || M <= 0 || C <= 0 || cap_seg <= 0) return YL_ERR_ARG;
    const PostLayout L = post_layout(B, M, C, cap_seg);
    if (ws_bytes < L.total) return YL_ERR_WORKSPACE;
    YL_CUDA_TRY(cudaMemsetAsync(ws, 0, L.counters_bytes, (cudaStream_t)stream));
    return YL_OK;
}

extern "C" int yl_filter_raw(const float *const *raw, const int *F, int n_layers, int B, int C,
                             const float *anchors_px, const int *anchor_mask, float conf_thre,
                             void *ws, size_t ws_bytes, long M, int cap_seg, int img_first, int img_count,
                             yl_stream_t stream)
{
    if (!raw || !F || !anchors_px || !anchor_mask || !ws) return YL_ERR_ARG;
    if (n_layers < 1 || n_layers > 3 || B <= 0 || C <= 0 || cap_seg <= 0) return YL_ERR_ARG;
    if (img_first < 0 || img_count < 0 || img_first + img_count > B) return YL_ERR_ARG;
    if (C > YL_MAX_CLASSES) return YL_ERR_CLASSES;
    long m_sum = 0;
    for (int l = 0; l < n_layers; ++l) { if (F[l] <= 0 || !raw[l]) return YL_ERR_ARG; m_sum += 3L * F[l] * F[l]; }
    if (m_sum != M) return YL_ERR_ARG;
    const PostLayout L = post_layout(B, M, C, cap_seg);
    if (ws_bytes < L.total) return YL_ERR_WORKSPACE;
    if (img_count == 0) return YL_OK;
    char *w = (char *)ws;
    unsigned *seg_count = (unsigned *)(w + L.off_seg_count);
    uint4 *cand = (uint4 *)(w + L.off_cand);
    float4 *boxtab = (float4 *)(w + L.off_box);
    float *objtab = (float *)(w + L.off_obj);
    const int NW = (C + 31) / 32;
    int row_off = 0;
    for (int l = 0; l < n_layers; ++l) {
        const int Fw = F[l], F2 = Fw * Fw;
        const float stride = (float)(8 << l);                               // yololayer.py:54
        float ag[6];
        for (int a = 0; a < 3; ++a) {                                       // yololayer.py:73-76 (doubles, then fp32)
            const int q = anchor_mask[3 * l + a];
            ag[2 * a] = (float)((double)anchors_px[2 * q] / (double)stride);
            ag[2 * a + 1] = (float)((double)anchors_px[2 * q + 1] / (double)stride);
        }
        // 128-bit loads need 16-byte aligned planes: F^2 % 4 == 0 and an aligned base (19x19 / 13x13 grids fall back)
        const bool vec4 = (F2 % 4 == 0) && (((uintptr_t)raw[l]) % 16 == 0);
        int rc;
        if (vec4) {
            dim3 grid((F2 / 4 + K1_THREADS - 1) / K1_THREADS, img_count * 3);
            rc = launch_filter_raw<4>(NW, grid, (cudaStream_t)stream, raw[l], Fw, F2, C, stride, ag, row_off, M, conf_thre,
                                      cap_seg, img_first, cand, seg_count, boxtab, objtab);
        } else {
            dim3 grid((F2 + K1_THREADS - 1) / K1_THREADS, img_count * 3);
            rc = launch_filter_raw<1>(NW, grid, (cudaStream_t)stream, raw[l], Fw, F2, C, stride, ag, row_off, M, conf_thre,
                                      cap_seg, img_first, cand, seg_count, boxtab, objtab);
        }
        if (rc != YL_OK) return rc;
        YL_LAUNCH_CHECK();
        row_off += 3 * F2;
    }
    return YL_OK;
}

extern "C" int yl_filter_dense(const float *pred, int B, long M, int C, int num_classes, float conf_thre,
                               void *ws, size_t ws_bytes, int cap_seg, int img_first, int img_count, yl_stream_t stream)
{
    if (!pred || !ws || B <= 0 || M <= 0 || C <= 0 || cap_seg <= 0) return YL_ERR_ARG;
    if (num_classes < 1 || num_classes > C) return YL_ERR_ARG;
    if (img_first < 0 || img_count < 0 || img_first + img_count > B) return YL_ERR_ARG;
    if (C > YL_MAX_CLASSES) return YL_ERR_CLASSES;
    const PostLayout L = post_layout(B, M, C, cap_seg);
    if (ws_bytes < L.total) return YL_ERR_WORKSPACE;
    if (img_count == 0) return YL_OK;
    char *w = (char *)ws;
    const long n_rows = (long)img_count * M;
    const long blocks_needed = (n_rows * 32 + KD_THREADS - 1) / KD_THREADS;
    const int grid = (int)(blocks_needed < 148L * 64 ? blocks_needed : 148L * 64);
    k_filter_dense<<<grid, KD_THREADS, 0, (cudaStream_t)stream>>>(
        pred, M, C, num_classes, conf_thre, cap_seg, img_first, n_rows, (uint4 *)(w + L.off_cand),
        (unsigned *)(w + L.off_seg_count), (float4 *)(w + L.off_box), (float *)(w + L.off_obj));
    YL_LAUNCH_CHECK();
    return YL_OK;
}
