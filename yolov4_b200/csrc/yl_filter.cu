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

// One scale of the head as the kernel sees it.
struct RawLayer {
    const float *raw;      // [B, 3, 5+C, F, F]
    int Fw, F2;
    int row_off;           // rows of the lower scales in the concatenated [M] axis (yolov4.py:324)
    int tiles;             // CTAs along x for this scale
    int vec;               // 4 = 128-bit loads, 1 = scalar loads (planes not 16-byte aligned, e.g. 19x19)
    float stride;
    float aw[3], ah[3];    // masked anchors in grid units (yololayer.py:73-76)
};
struct RawParams {
    RawLayer layer[3];
    int n_layers, C, cap_seg, img_first;
    long M;
    float thr;
    uint4 *cand;
    unsigned *seg_count;
    float4 *boxtab;
    float *objtab;
};

constexpr int K1_WARPS = K1_THREADS / 32;
constexpr int K1_QCAP = 512;           // flagged (box,class) pairs a warp resolves cooperatively per tile

struct K1Smem {
    unsigned short ent[K1_WARPS][K1_QCAP];   // bit15 pass | box slot (7b) << 8 | class (7b)
    unsigned cls[K1_WARPS][K1_QCAP];         // sigmoid(class logit) bits of passing entries
    float obj[K1_WARPS][128];
    unsigned any[K1_WARPS][4], nan[K1_WARPS][4];
};

// Decode one box (yololayer.py:150-162) and convert to corners (utils.py:117-126).
__device__ __forceinline__ float4 decode_box(const float *bp, int F2, int Fw, int p, float aw, float ah, float stride)
{
    const float tx = bp[0], ty = bp[(size_t)F2], tw = bp[2 * (size_t)F2], th = bp[3 * (size_t)F2];
    const int gy = p / Fw, gx = p - gy * Fw;
    const float bx = __fmul_rn(__fadd_rn(spec_sigmoidf(tx), (float)gx), stride);
    const float by = __fmul_rn(__fadd_rn(spec_sigmoidf(ty), (float)gy), stride);
    const float bw = __fmul_rn(__fmul_rn(spec_expf(tw), aw), stride);
    const float bh = __fmul_rn(__fmul_rn(spec_expf(th), ah), stride);
    const float hw = __fmul_rn(bw, 0.5f), hh = __fmul_rn(bh, 0.5f);
    return make_float4(__fsub_rn(bx, hw), __fsub_rn(by, hh), __fadd_rn(bx, hw), __fadd_rn(by, hh));
}

// One tile = K1_THREADS*VEC consecutive boxes of one (image, anchor).  Phase 1 streams the class planes with one
// compare per logit; phase 2 resolves the ~1% flagged pairs warp-cooperatively (all lanes, all loads in flight
// together) instead of serially in the lane that owns the box.
template <int VEC, int NW>
__device__ __forceinline__ void filter_tile(const RawParams &P, const RawLayer &Ly, int tile, int ba, K1Smem &sm)
{
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = P.C, F2 = Ly.F2;
    const float thr = P.thr;
    const int p0 = (tile * K1_THREADS + threadIdx.x) * VEC;
    const bool inb = p0 < F2;                                        // F2 % VEC == 0, so a vector is all in or all out
    const int b = ba / 3, a = ba - 3 * b;
    const int nch = 5 + C;
    const float *base = Ly.raw + ((size_t)ba * nch) * F2 + (inb ? p0 : 0);
    const float *cp = base + 5 * (size_t)F2;

    float obj[VEC], lth[VEC];
    bool any_alive = false;
    if (inb) {
        Vec<VEC> tob;
        tob.load(base + 4 * (size_t)F2);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            obj[v] = spec_sigmoidf(tob.v[v]);
            lth[v] = class_logit_bound(obj[v], thr);
            any_alive |= (lth[v] != kInf);
        }
    } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) { obj[v] = 0.0f; lth[v] = kInf; }
    }

    // ---- phase 1: streaming pass, result bits kept in registers -------------------------------------------------
    unsigned bits[VEC][NW];
#pragma unroll
    for (int v = 0; v < VEC; ++v)
#pragma unroll
        for (int w = 0; w < NW; ++w) bits[v][w] = 0u;
    if (any_alive) {
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
    }
    int cnt = 0;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        if (lth[v] == kInf) {
#pragma unroll
            for (int w = 0; w < NW; ++w) bits[v][w] = 0u;
        }
#pragma unroll
        for (int w = 0; w < NW; ++w) cnt += __popc(bits[v][w]);
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += n;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    if (total == 0) return;                                          // warp-uniform

    const float aw = Ly.aw[a], ah = Ly.ah[a];
    const int row_base = Ly.row_off + a * F2;                        // + p = row inside the image
    if (total <= K1_QCAP) {
        // ---- phase 2: cooperative exact pass ------------------------------------------------------------------
        unsigned short *ent = sm.ent[warp];
        unsigned *ecls = sm.cls[warp];
        int e = incl - cnt;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            sm.obj[warp][lane * VEC + v] = obj[v];
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                unsigned m = bits[v][w];
                while (m) {
                    const int bit = __ffs(m) - 1;
                    m &= m - 1;
                    ent[e++] = (unsigned short)(((lane * VEC + v) << 8) | (32 * w + bit));
                }
            }
        }
        if (lane < 4) { sm.any[warp][lane] = 0u; sm.nan[warp][lane] = 0u; }
        __syncwarp();
        const int wp0 = __shfl_sync(FULL, p0, 0);                    // first box of the warp (lane 0 is in bounds when total > 0)
        const float *wbase = Ly.raw + ((size_t)ba * nch) * F2 + wp0;
        const float *wcp = wbase + 5 * (size_t)F2;
        for (int e0 = 0; e0 < total; e0 += 128) {
            float t[4];
            unsigned en[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int q = e0 + 32 * u + lane;
                en[u] = (q < total) ? ent[q] : 0xFFFFu;
                t[u] = (q < total) ? wcp[(size_t)(en[u] & 0x7F) * F2 + (en[u] >> 8)] : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int q = e0 + 32 * u + lane;
                if (q < total) {
                    const int bs = en[u] >> 8;
                    // a NaN class logit makes torch.max (utils.py:139) NaN and drops the whole row (:145)
                    if (t[u] != t[u]) atomicOr(&sm.nan[warp][bs >> 5], 1u << (bs & 31));
                    const float cls = spec_sigmoidf(t[u]);
                    if (__fmul_rn(sm.obj[warp][bs], cls) >= thr) {
                        atomicOr(&sm.any[warp][bs >> 5], 1u << (bs & 31));
                        ecls[q] = __float_as_uint(cls);
                        ent[q] = (unsigned short)(en[u] | 0x8000u);
                    }
                }
            }
        }
        __syncwarp();
        // boxes with at least one surviving pair: decode once, store corners + objectness
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const int bs = lane + 32 * j;
            const unsigned live = (sm.any[warp][bs >> 5] & ~sm.nan[warp][bs >> 5]) >> (bs & 31) & 1u;
            if (live) {
                const int p = wp0 + bs;
                const size_t brow = (size_t)b * P.M + (row_base + p);
                P.boxtab[brow] = decode_box(wbase + bs, F2, Ly.Fw, p, aw, ah, Ly.stride);
                P.objtab[brow] = sm.obj[warp][bs];
            }
        }
        // one record per surviving (box,class) pair
        for (int q = lane; q < total; q += 32) {
            const unsigned en = ent[q];
            const int bs = (en >> 8) & 0x7F;
            if ((en & 0x8000u) && !((sm.nan[warp][bs >> 5] >> (bs & 31)) & 1u)) {
                const int k = en & 0x7F;
                const float cls = __uint_as_float(ecls[q]);
                const float s = __fadd_rn(__fmul_rn(sm.obj[warp][bs], cls), 0.0f);      // +0 canonicalises -0
                const unsigned seg = (unsigned)(b * C + k);
                const unsigned slot = atomicAdd(&P.seg_count[seg], 1u);
                if (slot < (unsigned)P.cap_seg)
                    P.cand[(size_t)seg * P.cap_seg + slot] =
                        make_uint4(__float_as_uint(s), (unsigned)(row_base + wp0 + bs), __float_as_uint(cls), 0u);
            }
        }
        __syncwarp();
        return;
    }

    // ---- dense fallback (more than K1_QCAP flagged pairs in the warp: degenerate inputs): per-lane serial pass ------
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        unsigned any = 0u;
#pragma unroll
        for (int w = 0; w < NW; ++w) any |= bits[v][w];
        if (!any) continue;
        bool nan_row = false;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            unsigned m = bits[v][w];
            while (m) {
                const int bit = __ffs(m) - 1;
                m &= m - 1;
                const float t = cp[(size_t)(32 * w + bit) * F2 + v];
                nan_row |= (t != t);
                if (!(__fmul_rn(obj[v], spec_sigmoidf(t)) >= thr)) bits[v][w] &= ~(1u << bit);
            }
        }
        if (nan_row) continue;
        any = 0u;
#pragma unroll
        for (int w = 0; w < NW; ++w) any |= bits[v][w];
        if (!any) continue;
        const int p = p0 + v;
        const unsigned row = (unsigned)(row_base + p);
        const size_t brow = (size_t)b * P.M + row;
        P.boxtab[brow] = decode_box(base + v, F2, Ly.Fw, p, aw, ah, Ly.stride);
        P.objtab[brow] = obj[v];
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            unsigned m = bits[v][w];
            while (m) {
                const int bit = __ffs(m) - 1;
                m &= m - 1;
                const int k = 32 * w + bit;
                const float cls = spec_sigmoidf(cp[(size_t)k * F2 + v]);
                const float s = __fadd_rn(__fmul_rn(obj[v], cls), 0.0f);
                const unsigned seg = (unsigned)(b * C + k);
                const unsigned slot = atomicAdd(&P.seg_count[seg], 1u);
                if (slot < (unsigned)P.cap_seg)
                    P.cand[(size_t)seg * P.cap_seg + slot] = make_uint4(__float_as_uint(s), row, __float_as_uint(cls), 0u);
            }
        }
    }
}

// grid = (sum of tiles over the scales, img_count*3): one launch covers all scales of an image group.
template <int NW>
__global__ void __launch_bounds__(K1_THREADS)
k_filter_raw(const __grid_constant__ RawParams P)
{
    __shared__ K1Smem sm;
    const int ba = P.img_first * 3 + blockIdx.y;
    int tile = blockIdx.x;
    int l = 0;
    while (l < P.n_layers - 1 && tile >= P.layer[l].tiles) { tile -= P.layer[l].tiles; ++l; }
    if (P.layer[l].vec == 4) filter_tile<4, NW>(P, P.layer[l], tile, ba, sm);
    else filter_tile<1, NW>(P, P.layer[l], tile, ba, sm);
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
    RawParams P;
    P.n_layers = n_layers; P.C = C; P.cap_seg = cap_seg; P.img_first = img_first; P.M = M; P.thr = conf_thre;
    P.cand = cand; P.seg_count = seg_count; P.boxtab = boxtab; P.objtab = objtab;
    int row_off = 0, tiles_total = 0;
    for (int l = 0; l < 3; ++l) {
        RawLayer &Ly = P.layer[l];
        if (l >= n_layers) { Ly = P.layer[0]; Ly.tiles = 0; continue; }
        Ly.raw = raw[l]; Ly.Fw = F[l]; Ly.F2 = F[l] * F[l]; Ly.row_off = row_off;
        Ly.stride = (float)(8 << l);                                        // yololayer.py:54
        for (int a = 0; a < 3; ++a) {                                       // yololayer.py:73-76 (doubles, then fp32)
            const int q = anchor_mask[3 * l + a];
            if (q < 0 || q > 8) return YL_ERR_ARG;
            Ly.aw[a] = (float)((double)anchors_px[2 * q] / (double)Ly.stride);
            Ly.ah[a] = (float)((double)anchors_px[2 * q + 1] / (double)Ly.stride);
        }
        // 128-bit loads need 16-byte aligned planes: F^2 % 4 == 0 and an aligned base (19x19 / 13x13 grids fall back)
        Ly.vec = ((Ly.F2 % 4 == 0) && (((uintptr_t)raw[l]) % 16 == 0)) ? 4 : 1;
        Ly.tiles = (Ly.F2 / Ly.vec + K1_THREADS - 1) / K1_THREADS;
        tiles_total += Ly.tiles;
        row_off += 3 * Ly.F2;
    }
    dim3 grid(tiles_total, img_count * 3);
    cudaStream_t st = (cudaStream_t)stream;
    switch ((C + 31) / 32) {
    case 1: k_filter_raw<1><<<grid, K1_THREADS, 0, st>>>(P); break;
    case 2: k_filter_raw<2><<<grid, K1_THREADS, 0, st>>>(P); break;
    case 3: k_filter_raw<3><<<grid, K1_THREADS, 0, st>>>(P); break;
    case 4: k_filter_raw<4><<<grid, K1_THREADS, 0, st>>>(P); break;
    default: return YL_ERR_CLASSES;
    }
    YL_LAUNCH_CHECK();
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
