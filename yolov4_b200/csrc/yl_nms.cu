// yl_nms.cu -- back end of the postprocess pipeline.
//
//   k_segment_nms   one CTA per (image,class) candidate segment: sort by (score desc, box row desc)
//                   [utils.py:58 with stable tie order], greedy NMS dropping a box when IoU >= thr with any
//                   already kept box [utils.py:67-84], kept records compacted to the front of the segment.
//   k_gather_rows   class-ascending concatenation of the kept rows of an image [utils.py:191-220] into
//                   out_rows[b] = (x1,y1,x2,y2,obj,cls_conf,cls) and the per-image counts.
//
// Segments up to SMEM_R candidates are processed entirely in shared memory; larger ones (degenerate inputs
// such as the all-ties random-init case) run the same algorithm in place in global memory.
#include "yl_common.cuh"
#include "../../include/yolo_head.h"

namespace yl {

constexpr int NMS_THREADS = 128;
constexpr int SMEM_R = kSmemR;
constexpr int CHUNK = 64;

// IoU >= thr decision, bit-equal to utils.py:71-77 (fp32, separate roundings, (area_a + area_b) - inter, IEEE divide).
// Fast path (thr > 0, boxes sanitised so that NaN coordinates can never overlap): a pair whose intersection
// is empty has iou = 0 or NaN, neither of which is >= thr, so only overlapping pairs pay for the division.
__device__ __forceinline__ bool suppresses_pos(const float4 &a, float area_a, const float4 &b, float area_b, float thr)
{
    const float tlx = fmaxf(a.x, b.x), tly = fmaxf(a.y, b.y);
    const float brx = fminf(a.z, b.z), bry = fminf(a.w, b.w);
    if (!(tlx < brx && tly < bry)) return false;
    const float inter = __fmul_rn(__fsub_rn(brx, tlx), __fsub_rn(bry, tly));
    const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
    return iou >= thr;
}

// General path (thr <= 0 or unsanitised boxes): literal NaN-propagating formula.
__device__ __forceinline__ bool suppresses_any(const float4 &a, float area_a, const float4 &b, float area_b, float thr)
{
    const float tlx = nanmaxf(a.x, b.x), tly = nanmaxf(a.y, b.y);
    const float brx = nanminf(a.z, b.z), bry = nanminf(a.w, b.w);
    const float en = (tlx < brx && tly < bry) ? 1.0f : 0.0f;
    const float inter = __fmul_rn(__fmul_rn(__fsub_rn(brx, tlx), __fsub_rn(bry, tly)), en);
    const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
    return iou >= thr;
}

__device__ __forceinline__ float box_area(const float4 &b) { return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y)); }

__device__ __forceinline__ float4 sanitise(const float4 &b)
{
    // A box with a NaN coordinate never suppresses and is never suppressed (np.maximum/np.minimum propagate the
    // NaN, (tl < br) is False, iou is NaN).  Give it an empty extent so the fmaxf/fminf fast path agrees.
    if (b.x != b.x || b.y != b.y || b.z != b.z || b.w != b.w) return make_float4(kInf, kInf, -kInf, -kInf);
    return b;
}

// ---- storage back ends ------------------------------------------------------------------------------------
struct SmemStore {
    unsigned long long *key;   // ascending sort key: (~score order bits) << 32 | ~row
    unsigned *conf;
    float4 *box;               // sanitised when POS
    float *area;
    __device__ __forceinline__ unsigned long long get_key(int i) const { return key[i]; }
    __device__ __forceinline__ void cswap(int i, int j)
    {
        const unsigned long long a = key[i], b = key[j];
        if (b < a) { key[i] = b; key[j] = a; const unsigned t = conf[i]; conf[i] = conf[j]; conf[j] = t; }
    }
    __device__ __forceinline__ float4 get_box(int i) const { return box[i]; }
    __device__ __forceinline__ float get_area(int i) const { return area[i]; }
};

struct GlobalStore {
    uint4 *rec;                // {key hi, key lo, conf, 0}
    const float4 *boxes;       // boxtab + b*M
    bool pos;
    __device__ __forceinline__ unsigned long long get_key(int i) const
    {
        const uint4 r = rec[i];
        return ((unsigned long long)r.x << 32) | r.y;
    }
    __device__ __forceinline__ void cswap(int i, int j)
    {
        const uint4 a = rec[i], b = rec[j];
        const unsigned long long ka = ((unsigned long long)a.x << 32) | a.y, kb = ((unsigned long long)b.x << 32) | b.y;
        if (kb < ka) { rec[i] = b; rec[j] = a; }
    }
    __device__ __forceinline__ float4 get_box(int i) const
    {
        const float4 b = boxes[~rec[i].y];
        return pos ? sanitise(b) : b;
    }
    __device__ __forceinline__ float get_area(int i) const { return box_area(boxes[~rec[i].y]); }
};

// Bitonic network in the "flip" formulation: every compare-exchange orders ascending, so virtual +inf padding
// beyond n never moves and non-power-of-two n works by skipping out-of-range partners.
template <class Store>
__device__ void bitonic_sort(Store &s, int n)
{
    int P = 1;
    while (P < n) P <<= 1;
    for (int k = 2; k <= P; k <<= 1) {
        for (int i = threadIdx.x; i < n; i += NMS_THREADS) {
            const int l = i ^ (k - 1);
            if (l > i && l < n) s.cswap(i, l);
        }
        __syncthreads();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += NMS_THREADS) {
                const int l = i ^ j;
                if (l > i && l < n) s.cswap(i, l);
            }
            __syncthreads();
        }
    }
}

// Greedy NMS over the sorted segment, 64 candidates at a time:
//   (1) each chunk member is tested against every box kept so far (all threads, pairs in parallel)
//   (2) the 64x64 intra-chunk suppression matrix is built as one 64-bit row per member
//   (3) warp 0 resolves the chunk serially with the rows in registers
//   (4) survivors are appended to the kept list (compacted in place: a kept box never moves to a higher index)
// kept_of[q] = sorted position of the q-th kept box.  Returns the number kept.
template <bool POS, class Store>
__device__ int greedy_nms(Store &s, int n, float thr, unsigned short *kept_s, unsigned *kept_g,
                          unsigned *sh_supp /*[2]*/, unsigned long long *sh_rows /*[64]*/, int *sh_nk)
{
    const int tid = threadIdx.x;
    if (tid == 0) *sh_nk = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += CHUNK) {
        const int cn = min(CHUNK, n - c0);
        const int nk = *sh_nk;
        if (tid < 2) sh_supp[tid] = 0u;
        if (tid < CHUNK) sh_rows[tid] = 0ull;
        __syncthreads();
        // (1) chunk x kept
        {
            const int i = tid & (CHUNK - 1);
            if (i < cn) {
                const float4 bi = s.get_box(c0 + i);
                const float ai = s.get_area(c0 + i);
                bool dead = false;
                for (int q = tid / CHUNK; q < nk && !dead; q += NMS_THREADS / CHUNK) {
                    const int kq = kept_s ? (int)kept_s[q] : (int)kept_g[q];
                    const float4 bk = s.get_box(kq);
                    const float ak = s.get_area(kq);
                    dead = POS ? suppresses_pos(bi, ai, bk, ak, thr) : suppresses_any(bi, ai, bk, ak, thr);
                }
                if (dead) atomicOr(&sh_supp[i >> 5], 1u << (i & 31));
            }
        }
        // (2) intra-chunk rows: row i has bit j set (j > i) when i suppresses j
        for (int pr = tid; pr < CHUNK * CHUNK; pr += NMS_THREADS) {
            const int i = pr / CHUNK, j = pr - i * CHUNK;
            if (j > i && j < cn) {
                const float4 bi = s.get_box(c0 + i), bj = s.get_box(c0 + j);
                const bool sup = POS ? suppresses_pos(bj, s.get_area(c0 + j), bi, s.get_area(c0 + i), thr)
                                     : suppresses_any(bj, s.get_area(c0 + j), bi, s.get_area(c0 + i), thr);
                if (sup) atomicOr(&sh_rows[i], 1ull << j);
            }
        }
        __syncthreads();
        // (3) serial resolve by warp 0
        if (tid < 32) {
            const unsigned long long r0 = sh_rows[tid], r1 = sh_rows[tid + 32];
            unsigned long long removed = ((unsigned long long)sh_supp[1] << 32) | sh_supp[0];
            unsigned long long keepbits = 0ull;
            for (int i = 0; i < cn; ++i) {
                const unsigned long long src = (i < 32) ? r0 : r1;
                const unsigned long long row = __shfl_sync(0xFFFFFFFFu, src, i & 31);
                if (!((removed >> i) & 1ull)) { keepbits |= 1ull << i; removed |= row; }
            }
            // (4) append survivors
            for (int i = tid; i < cn; i += 32)
                if ((keepbits >> i) & 1ull) {
                    const int q = nk + __popcll(keepbits & ((1ull << i) - 1ull));
                    if (kept_s) kept_s[q] = (unsigned short)(c0 + i); else kept_g[q] = (unsigned)(c0 + i);
                }
            if (tid == 0) *sh_nk = nk + __popcll(keepbits);
        }
        __syncthreads();
    }
    return *sh_nk;
}

// ---------------------------------------------------------------------------------------------------------------
// Small tier (the common case: ~150 candidates per (image,class) at conf 1e-4): everything in ~18 KB of shared memory.
//   1. rank sort on the unique 64-bit key (each thread counts the keys below its own; broadcast reads, no barriers)
//   2. transposed suppression matrix T[j] = { i < j : IoU(i,j) >= thr }, one row per thread, all pairs independent
//   3. warp 0 walks j in score order: kept[j] = (T[j] & kept) == 0        (utils.py:67-84)
// Segments with more than SMALL_R candidates are queued for the big tier.
// ---------------------------------------------------------------------------------------------------------------
constexpr int SMALL_R = 256;
constexpr int SMALL_W = SMALL_R / 32;
constexpr int SMALL_THREADS = 128;

template <bool POS>
__device__ __forceinline__ void small_pairs(const float4 *sh_box, const float *sh_area, unsigned (*sh_T)[SMALL_W + 1], int n, float thr)
{
    for (int j = threadIdx.x; j < n; j += SMALL_THREADS) {
        const float4 bj = sh_box[j];
        const float aj = sh_area[j];
#pragma unroll
        for (int w = 0; w < SMALL_W; ++w) {
            unsigned word = 0u;
            const int lim = min(32, j - 32 * w);
            for (int ii = 0; ii < lim; ++ii) {
                const int i = 32 * w + ii;
                const bool sup = POS ? suppresses_pos(bj, aj, sh_box[i], sh_area[i], thr) : suppresses_any(bj, aj, sh_box[i], sh_area[i], thr);
                word |= sup ? (1u << ii) : 0u;
            }
            sh_T[j][w] = word;
        }
    }
}

__global__ void __launch_bounds__(SMALL_THREADS)
k_segment_nms_small(uint4 *__restrict__ cand, const unsigned *__restrict__ seg_count, unsigned *__restrict__ kept_count,
                    const float4 *__restrict__ boxtab, long M, int C, int cap_seg, float thr, int seg_first,
                    unsigned *__restrict__ big_count, unsigned *__restrict__ big_list)
{
    __shared__ unsigned long long sh_key[SMALL_R];
    __shared__ unsigned sh_row[SMALL_R], sh_conf[SMALL_R];      // sorted
    __shared__ float4 sh_box[SMALL_R];
    __shared__ float sh_area[SMALL_R];
    __shared__ unsigned sh_T[SMALL_R][SMALL_W + 1];             // +1: conflict-free row writes
    __shared__ unsigned sh_kw[SMALL_W], sh_pref[SMALL_W + 1];

    const int seg = seg_first + blockIdx.x;
    const unsigned cnt = seg_count[seg];
    if (cnt == 0u) return;                                      // kept_count was zeroed by yl_post_reset
    if (cnt > (unsigned)cap_seg) return;                        // overflow: reported through meta[], caller re-runs
    const int tid = threadIdx.x;
    if (cnt > (unsigned)SMALL_R) {
        if (tid == 0) big_list[atomicAdd(big_count, 1u)] = (unsigned)seg;
        return;
    }
    const int n = (int)cnt;
    const int b = seg / C;
    uint4 *rec = cand + (size_t)seg * cap_seg;
    const float4 *boxes = boxtab + (size_t)b * M;
    const bool pos = thr > 0.0f;

    // load (at most two records per thread), start the box gathers, then rank
    unsigned long long key[2];
    unsigned conf[2], row[2];
    float4 bx[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int i = tid + u * SMALL_THREADS;
        key[u] = ~0ull; conf[u] = 0u; row[u] = 0u;
        if (i < n) {
            const uint4 r = rec[i];
            key[u] = ((unsigned long long)score_desc_bits(r.x) << 32) | (unsigned)(~r.y);
            conf[u] = r.z; row[u] = r.y;
            sh_key[i] = key[u];
        }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
        if (tid + u * SMALL_THREADS < n) bx[u] = boxes[row[u]];
    __syncthreads();
    int rank[2] = {0, 0};
    for (int j = 0; j < n; ++j) {
        const unsigned long long k = sh_key[j];
        rank[0] += (k < key[0]) ? 1 : 0;
        rank[1] += (k < key[1]) ? 1 : 0;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
        if (tid + u * SMALL_THREADS < n) {
            const int r = rank[u];
            sh_row[r] = row[u]; sh_conf[r] = conf[u];
            sh_area[r] = box_area(bx[u]);
            sh_box[r] = pos ? sanitise(bx[u]) : bx[u];
        }
    __syncthreads();
    if (pos) small_pairs<true>(sh_box, sh_area, sh_T, n, thr);
    else small_pairs<false>(sh_box, sh_area, sh_T, n, thr);
    __syncthreads();
    if (tid < 32) {
        unsigned keptw = 0u;                                    // lane w owns kept word w
        const int lw = tid < SMALL_W ? tid : 0;
        int j = 0;
        for (; j + 4 <= n; j += 4) {
            unsigned t0 = sh_T[j][lw], t1 = sh_T[j + 1][lw], t2 = sh_T[j + 2][lw], t3 = sh_T[j + 3][lw];
            if (tid >= SMALL_W) { t0 = t1 = t2 = t3 = 0u; }
            if (!__any_sync(0xFFFFFFFFu, t0 & keptw) && tid == (j >> 5)) keptw |= 1u << (j & 31);
            if (!__any_sync(0xFFFFFFFFu, t1 & keptw) && tid == ((j + 1) >> 5)) keptw |= 1u << ((j + 1) & 31);
            if (!__any_sync(0xFFFFFFFFu, t2 & keptw) && tid == ((j + 2) >> 5)) keptw |= 1u << ((j + 2) & 31);
            if (!__any_sync(0xFFFFFFFFu, t3 & keptw) && tid == ((j + 3) >> 5)) keptw |= 1u << ((j + 3) & 31);
        }
        for (; j < n; ++j) {
            const unsigned t0 = (tid < SMALL_W) ? sh_T[j][lw] : 0u;
            if (!__any_sync(0xFFFFFFFFu, t0 & keptw) && tid == (j >> 5)) keptw |= 1u << (j & 31);
        }
        int c = (tid < SMALL_W) ? __popc(keptw) : 0, incl = c;
#pragma unroll
        for (int o = 1; o < SMALL_W; o <<= 1) {
            const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (tid >= o) incl += v;
        }
        if (tid < SMALL_W) { sh_kw[tid] = keptw; sh_pref[tid] = (unsigned)(incl - c); }
        if (tid == SMALL_W - 1) { sh_pref[SMALL_W] = (unsigned)incl; kept_count[seg] = (unsigned)incl; }
    }
    __syncthreads();
    for (int j2 = tid; j2 < n; j2 += SMALL_THREADS) {
        const unsigned kw = sh_kw[j2 >> 5];
        if ((kw >> (j2 & 31)) & 1u) {
            const unsigned q = sh_pref[j2 >> 5] + __popc(kw & ((1u << (j2 & 31)) - 1u));
            rec[q] = make_uint4(sh_row[j2], sh_conf[j2], 0u, 0u);        // {box row, cls_conf bits}
        }
    }
}

// Big tier: persistent CTAs walk the list of segments the small tier could not take (more than SMALL_R candidates).
__global__ void __launch_bounds__(NMS_THREADS)
k_segment_nms_big(uint4 *__restrict__ cand, const unsigned *__restrict__ seg_count, unsigned *__restrict__ kept_count,
                  const float4 *__restrict__ boxtab, long M, int C, int cap_seg, float thr,
                  const unsigned *__restrict__ big_count, const unsigned *__restrict__ big_list,
                  unsigned *__restrict__ kept_scratch /* [B*C*cap_seg] u32, only when cap_seg > SMEM_R; else null */)
{
    __shared__ unsigned long long sh_key[SMEM_R];
    __shared__ unsigned sh_conf[SMEM_R];
    __shared__ float4 sh_box[SMEM_R];
    __shared__ float sh_area[SMEM_R];
    __shared__ unsigned short sh_kept[SMEM_R];
    __shared__ unsigned long long sh_rows[CHUNK];
    __shared__ unsigned sh_supp[2];
    __shared__ int sh_nk;

    const unsigned nbig = *big_count;
    for (unsigned it = blockIdx.x; it < nbig; it += gridDim.x) {
    const int seg = (int)big_list[it];
    const int b = seg / C;
    const unsigned cnt = seg_count[seg];
    const int n = (int)cnt;
    uint4 *rec = cand + (size_t)seg * cap_seg;
    const float4 *boxes = boxtab + (size_t)b * M;
    const int tid = threadIdx.x;
    const bool pos = thr > 0.0f;

    if (n <= SMEM_R) {
        for (int i = tid; i < n; i += NMS_THREADS) {
            const uint4 r = rec[i];
            sh_key[i] = ((unsigned long long)score_desc_bits(r.x) << 32) | (unsigned)(~r.y);
            sh_conf[i] = r.z;
        }
        __syncthreads();
        SmemStore s{sh_key, sh_conf, sh_box, sh_area};
        bitonic_sort(s, n);
        for (int i = tid; i < n; i += NMS_THREADS) {
            const float4 bx = boxes[~(unsigned)sh_key[i]];
            sh_area[i] = box_area(bx);
            sh_box[i] = pos ? sanitise(bx) : bx;
        }
        __syncthreads();
        const int nk = pos ? greedy_nms<true>(s, n, thr, sh_kept, nullptr, sh_supp, sh_rows, &sh_nk)
                           : greedy_nms<false>(s, n, thr, sh_kept, nullptr, sh_supp, sh_rows, &sh_nk);
        for (int q = tid; q < nk; q += NMS_THREADS) {
            const int i = sh_kept[q];
            rec[q] = make_uint4(~(unsigned)sh_key[i], sh_conf[i], 0u, 0u);      // {box row, cls_conf bits}
        }
        if (tid == 0) kept_count[seg] = (unsigned)nk;
    } else {
        // oversized segment: same algorithm in place in global memory (slow path, correctness only)
        for (int i = tid; i < n; i += NMS_THREADS) {
            const uint4 r = rec[i];
            rec[i] = make_uint4(score_desc_bits(r.x), ~r.y, r.z, 0u);
        }
        __syncthreads();
        GlobalStore s{rec, boxes, pos};
        bitonic_sort(s, n);
        unsigned *kept_g = kept_scratch + (size_t)seg * cap_seg;
        const int nk = pos ? greedy_nms<true>(s, n, thr, nullptr, kept_g, sh_supp, sh_rows, &sh_nk)
                           : greedy_nms<false>(s, n, thr, nullptr, kept_g, sh_supp, sh_rows, &sh_nk);
        // compact in place: kept_g[q] >= q and strictly increasing, so go through registers chunk by chunk
        for (int q0 = 0; q0 < nk; q0 += NMS_THREADS) {
            const int q = q0 + tid;
            uint4 r = make_uint4(0, 0, 0, 0);
            if (q < nk) r = rec[kept_g[q]];
            __syncthreads();
            if (q < nk) rec[q] = make_uint4(~r.y, r.z, 0u, 0u);
            __syncthreads();
        }
        if (tid == 0) kept_count[seg] = (unsigned)nk;
    }
    __syncthreads();
    }
}

constexpr int GATHER_THREADS = 128;

__global__ void __launch_bounds__(GATHER_THREADS)
k_gather_rows(const uint4 *__restrict__ cand, const unsigned *__restrict__ seg_count, const unsigned *__restrict__ kept_count,
              const float4 *__restrict__ boxtab, const float *__restrict__ objtab, long M, int C, int cap_seg, int B,
              float *__restrict__ out_rows, long cap_out, int *__restrict__ meta, int seg_first)
{
    __shared__ unsigned sh_red[3][GATHER_THREADS / 32];
    const int seg = seg_first + blockIdx.x;
    const int b = seg / C, c = seg - b * C;
    const int tid = threadIdx.x;
    // exclusive prefix of kept counts over lower classes; class 0 also reduces the candidate statistics
    unsigned pre = 0u, mx = 0u, tot = 0u;
    for (int k = tid; k < C; k += GATHER_THREADS) {
        if (k < c) pre += kept_count[b * C + k];
        if (c == 0) { const unsigned s = seg_count[b * C + k]; mx = max(mx, s); tot += s; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        pre += __shfl_xor_sync(0xFFFFFFFFu, pre, o);
        tot += __shfl_xor_sync(0xFFFFFFFFu, tot, o);
        mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    }
    if ((tid & 31) == 0) { sh_red[0][tid >> 5] = pre; sh_red[1][tid >> 5] = tot; sh_red[2][tid >> 5] = mx; }
    __syncthreads();
    pre = 0u; tot = 0u; mx = 0u;
#pragma unroll
    for (int w = 0; w < GATHER_THREADS / 32; ++w) { pre += sh_red[0][w]; tot += sh_red[1][w]; mx = max(mx, sh_red[2][w]); }
    const unsigned nk = kept_count[seg];
    if (tid == 0) {
        if (c == C - 1) meta[b] = (int)(pre + nk);
        if (c == 0) { meta[B + b] = (int)mx; meta[2 * B + b] = (int)tot; }
    }
    const uint4 *rec = cand + (size_t)seg * cap_seg;
    const float fc = (float)c;                                             // utils.py:183 class id stored as float
    for (unsigned q = tid; q < nk; q += GATHER_THREADS) {
        const long r = (long)pre + q;
        if (r >= cap_out) break;
        const uint4 e = rec[q];
        const size_t brow = (size_t)b * M + e.x;
        const float4 bx = boxtab[brow];
        float *o = out_rows + ((size_t)b * cap_out + r) * 7;
        o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
        o[4] = objtab[brow]; o[5] = __uint_as_float(e.y); o[6] = fc;
    }
}

}  // namespace yl

using namespace yl;

extern "C" int yl_nms(void *ws, size_t ws_bytes, int B, long M, int C, int cap_seg, float nms_thre,
                      float *out_rows, long cap_out, int *meta, int img_first, int img_count, yl_stream_t stream)
{
    if (!ws || !out_rows || !meta || B <= 0 || M <= 0 || C <= 0 || cap_seg <= 0 || cap_out <= 0) return YL_ERR_ARG;
    if (img_first < 0 || img_count < 0 || img_first + img_count > B) return YL_ERR_ARG;
    if (C > YL_MAX_CLASSES) return YL_ERR_CLASSES;
    const PostLayout L = post_layout(B, M, C, cap_seg);
    if (ws_bytes < L.total) return YL_ERR_WORKSPACE;
    if (img_count == 0) return YL_OK;
    char *w = (char *)ws;
    uint4 *cand = (uint4 *)(w + L.off_cand);
    unsigned *seg_count = (unsigned *)(w + L.off_seg_count);
    unsigned *kept_count = (unsigned *)(w + L.off_kept_count);
    const float4 *boxtab = (const float4 *)(w + L.off_box);
    const float *objtab = (const float *)(w + L.off_obj);
    unsigned *kept_scratch = (cap_seg > SMEM_R) ? (unsigned *)(w + L.off_kept_scratch) : nullptr;
    const int nseg = img_count * C, seg_first = img_first * C;
    unsigned *big_count = (unsigned *)(w + L.off_big_count) + img_first;
    unsigned *big_list = (unsigned *)(w + L.off_big_list) + seg_first;
    k_segment_nms_small<<<nseg, SMALL_THREADS, 0, (cudaStream_t)stream>>>(cand, seg_count, kept_count, boxtab, M, C, cap_seg,
                                                                        nms_thre, seg_first, big_count, big_list);
    YL_LAUNCH_CHECK();
    if (cap_seg > SMALL_R) {
        const int grid_big = nseg < 148 * 2 ? nseg : 148 * 2;
        k_segment_nms_big<<<grid_big, NMS_THREADS, 0, (cudaStream_t)stream>>>(cand, seg_count, kept_count, boxtab, M, C, cap_seg,
                                                                             nms_thre, big_count, big_list, kept_scratch);
        YL_LAUNCH_CHECK();
    }
    k_gather_rows<<<nseg, GATHER_THREADS, 0, (cudaStream_t)stream>>>(cand, seg_count, kept_count, boxtab, objtab, M, C,
                                                                   cap_seg, B, out_rows, cap_out, meta, seg_first);
    YL_LAUNCH_CHECK();
    return YL_OK;
}
