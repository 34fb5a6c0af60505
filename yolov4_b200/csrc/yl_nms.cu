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
#include <stdlib.h>

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
    // Overlapping boxes have positive extents, so fl(inter) <= min(area) and union >= max(area): boxes whose areas
    // differ by more than the threshold ratio cannot reach it (common across anchor scales), no matter the overlap.
    {
        const float amin = fminf(area_a, area_b), amax = fmaxf(area_a, area_b);
        if (amax < 3.0e38f && amin < 0.999f * thr * amax) return false;
    }
    const float inter = __fmul_rn(__fsub_rn(brx, tlx), __fsub_rn(bry, tly));
    const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    // fl(inter/uni) >= thr is decided without the division when inter is not within 0.1% of thr*uni (the quotient's
    // rounding error is 6e-8): only the rare near-threshold pair pays for the IEEE divide.
    if (uni > 0.0f && uni < 3.0e38f) {
        const float tu = thr * uni;
        if (inter < 0.999f * tu) return false;
        if (inter > 1.001f * tu && inter < 3.0e38f) return true;
    }
    return __fdiv_rn(inter, uni) >= thr;
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
#ifndef YL_SMALL_THREADS
#define YL_SMALL_THREADS 128
#endif
constexpr int SMALL_THREADS = YL_SMALL_THREADS;
constexpr int SMALL_EPT = SMALL_R / SMALL_THREADS;      // records per thread, at most

// Conservative 16-bit image of a (sanitised) box for the pair prefilter: corners are rounded outwards (floor for x1,y1,
// ceil for x2,y2), clamped to +-16384 px and biased to 15-bit unsigned, two per 32-bit word.  Rounding and clamping are
// monotone, so  true overlap (tl < br on both axes)  =>  q(tl) <= q(br) on both axes : the integer test never rejects
// a pair the exact fp32 test would accept.  A NaN box (sanitised to +inf/-inf) maps to lo=32767 > hi=0: never overlaps.
__device__ __forceinline__ uint2 quantise_box(const float4 &b)
{
    const int x1 = min(max(__float2int_rd(fminf(fmaxf(b.x, -16384.0f), 16383.0f)), -16384), 16383) + 16384;
    const int y1 = min(max(__float2int_rd(fminf(fmaxf(b.y, -16384.0f), 16383.0f)), -16384), 16383) + 16384;
    const int x2 = min(max(__float2int_ru(fminf(fmaxf(b.z, -16384.0f), 16383.0f)), -16384), 16383) + 16384;
    const int y2 = min(max(__float2int_ru(fminf(fmaxf(b.w, -16384.0f), 16383.0f)), -16384), 16383) + 16384;
    return make_uint2((unsigned)x1 | ((unsigned)y1 << 16), (unsigned)x2 | ((unsigned)y2 << 16));
}

// (br + 0x8000 - tl) per 16-bit half never borrows or overflows for 15-bit operands; bit 15 of a half is set iff br >= tl.
__device__ __forceinline__ bool may_overlap(const uint2 &a, const uint2 &b)
{
    const unsigned tl = __vmaxu2(a.x, b.x), br = __vminu2(a.y, b.y);
    return ((br + 0x80008000u - tl) & 0x80008000u) == 0x80008000u;
}

// Row j of the transposed suppression matrix: bit i (i < j) set when box i suppresses box j.
// Number of keys below `key` (keys are unique): LDS.128 per two keys, one 64-bit compare + predicated add per key.
__device__ __forceinline__ int rank_of(const unsigned long long *sh_key, int n, unsigned long long key)
{
    int r = 0;
    const ulonglong2 *k2 = reinterpret_cast<const ulonglong2 *>(sh_key);
    const int n2 = (n + 1) >> 1;
#pragma unroll 4
    for (int j = 0; j < n2; ++j) {
        const ulonglong2 k = k2[j];
        asm("{\n\t.reg .pred p, q;\n\tsetp.lt.u64 p, %1, %3;\n\tsetp.lt.u64 q, %2, %3;\n\t@p add.s32 %0, %0, 1;\n\t@q add.s32 %0, %0, 1;\n\t}"
            : "+r"(r) : "l"(k.x), "l"(k.y), "l"(key));
    }
    return r;
}

template <bool POS>
__device__ __forceinline__ void small_pairs(const float4 *sh_box, const float *sh_area, const uint2 *sh_q,
                                            unsigned (*sh_T)[SMALL_W + 1], int n, float thr)
{
    for (int j0 = 0; j0 < n; j0 += SMALL_THREADS) {
        const int j = j0 + (int)threadIdx.x;
        const int jw = (j0 + ((int)threadIdx.x & ~31)) >> 5;        // word of this warp's rows (warp-uniform)
        if (32 * jw >= n) break;                                     // the whole warp is past the segment
        const bool have = j < n;
        const float4 bj = sh_box[have ? j : 0];
        const float aj = sh_area[have ? j : 0];
        const uint2 qj = sh_q[have ? j : 0];
        for (int w = 0; w <= jw; ++w) {
            unsigned word = 0u;
            const unsigned lim_mask = (w < jw) ? 0xFFFFFFFFu : ((1u << (j & 31)) - 1u);      // diagonal block: only i < j
            if (POS) {
                // branch-free integer prefilter over the 32 boxes of the block, then the exact test on the few survivors
                unsigned cand = 0u;
#pragma unroll
                for (int ii = 0; ii < 32; ++ii) cand |= may_overlap(qj, sh_q[32 * w + ii]) ? (1u << ii) : 0u;
                cand &= lim_mask;
                while (cand) {
                    const int ii = __ffs(cand) - 1;
                    cand &= cand - 1;
                    const int i = 32 * w + ii;
                    if (suppresses_pos(bj, aj, sh_box[i], sh_area[i], thr)) word |= 1u << ii;
                }
            } else {
#pragma unroll 4
                for (int ii = 0; ii < 32; ++ii) {
                    const int i = 32 * w + ii;
                    if (((lim_mask >> ii) & 1u) && suppresses_any(bj, aj, sh_box[i], sh_area[i], thr)) word |= 1u << ii;
                }
            }
            if (have) sh_T[j][w] = word;
        }
    }
}

__global__ void __launch_bounds__(SMALL_THREADS)
k_segment_nms_small(uint4 *__restrict__ cand, const unsigned *__restrict__ seg_count, unsigned *__restrict__ kept_count,
                    const float4 *__restrict__ boxtab, long M, int C, int cap_seg, float thr, int seg_first,
                    unsigned *__restrict__ big_count, unsigned *__restrict__ big_list)
{
    __shared__ __align__(16) unsigned long long sh_key[SMALL_R + 2];
    __shared__ unsigned sh_row[SMALL_R], sh_conf[SMALL_R];      // sorted
    __shared__ float4 sh_box[SMALL_R];
    __shared__ float sh_area[SMALL_R];
    __shared__ uint2 sh_q[SMALL_R];                             // 16-bit conservative corners for the pair prefilter
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
    unsigned long long key[SMALL_EPT];
    unsigned conf[SMALL_EPT], row[SMALL_EPT];
    float4 bx[SMALL_EPT];
#pragma unroll
    for (int u = 0; u < SMALL_EPT; ++u) {
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
    for (int u = 0; u < SMALL_EPT; ++u)
        if (tid + u * SMALL_THREADS < n) bx[u] = boxes[row[u]];
    __syncthreads();
    int rank[SMALL_EPT];
#pragma unroll
    for (int u = 0; u < SMALL_EPT; ++u) rank[u] = 0;
    const int wbase = tid & ~31;                                // first element index of this warp (warp-uniform bounds below)
    // keys beyond n were never written: pad the tail of the last 16-byte pair so the 2-wide loop may read it
    if (tid == 0 && (n & 1)) sh_key[n] = ~0ull;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < SMALL_EPT; ++u)
        if (wbase + u * SMALL_THREADS < n) rank[u] = rank_of(sh_key, n, key[u]);          // warp-uniform bound
#pragma unroll
    for (int u = 0; u < SMALL_EPT; ++u)
        if (tid + u * SMALL_THREADS < n) {
            const int r = rank[u];
            sh_row[r] = row[u]; sh_conf[r] = conf[u];
            sh_area[r] = box_area(bx[u]);
            const float4 sb = pos ? sanitise(bx[u]) : bx[u];
            sh_box[r] = sb;
            sh_q[r] = quantise_box(sb);
        }
    __syncthreads();
    if (pos) small_pairs<true>(sh_box, sh_area, sh_q, sh_T, n, thr);
    else small_pairs<false>(sh_box, sh_area, sh_q, sh_T, n, thr);
    __syncthreads();
    if (tid < 32) {
        unsigned keptw = 0u;                                    // lane w owns kept word w
        const int lw = tid < SMALL_W ? tid : 0;
        int j = 0;
        // row j only has words 0..j/32 (the rest was never written): lanes above the diagonal contribute nothing
        for (; j + 4 <= n; j += 4) {
            const bool rd = tid <= (j >> 5);                    // j..j+3 share j>>5 (j is a multiple of 4)
            const unsigned t0 = rd ? sh_T[j][lw] : 0u, t1 = rd ? sh_T[j + 1][lw] : 0u;
            const unsigned t2 = rd ? sh_T[j + 2][lw] : 0u, t3 = rd ? sh_T[j + 3][lw] : 0u;
            if (!__any_sync(0xFFFFFFFFu, t0 & keptw) && tid == (j >> 5)) keptw |= 1u << (j & 31);
            if (!__any_sync(0xFFFFFFFFu, t1 & keptw) && tid == (j >> 5)) keptw |= 1u << ((j + 1) & 31);
            if (!__any_sync(0xFFFFFFFFu, t2 & keptw) && tid == (j >> 5)) keptw |= 1u << ((j + 2) & 31);
            if (!__any_sync(0xFFFFFFFFu, t3 & keptw) && tid == (j >> 5)) keptw |= 1u << ((j + 3) & 31);
        }
        for (; j < n; ++j) {
            const unsigned t0 = (tid <= (j >> 5)) ? sh_T[j][lw] : 0u;
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

// ---------------------------------------------------------------------------------------------------------------
// Small tier, warp-per-segment form (default): one warp owns one (image,class) segment end to end, so there are no
// block barriers and no warp ever idles while another resolves; a CTA is just WS_WARPS independent segments.
//   1. records -> shared memory (key, row, conf) + box gather; 2. rank sort on the unique key;
//   3. sorted conservative corners q[] + permutation;  4. per block of 32 sorted candidates: integer prefilter against
//   every kept earlier box (exact fp32 test only on prefilter hits), 32x32 diagonal block, serial resolve by ballot,
//   kept records appended in score order.
// ---------------------------------------------------------------------------------------------------------------
constexpr int WS_WARPS = 4;
constexpr int WS_U = SMALL_R / 32;         // records per lane, at most

struct alignas(16) WsSeg {
    union {
        unsigned long long key[SMALL_R + 2];   // while ranking
        uint2 q[SMALL_R];                      // afterwards: conservative 16-bit corners in sorted order
    };
    float4 ubox[SMALL_R];                      // record order; sanitised when thr > 0
    unsigned urow[SMALL_R], uconf[SMALL_R];    // record order
    unsigned short perm[SMALL_R];              // sorted position -> record index
};

template <bool POS>
__device__ __forceinline__ void warp_segment_nms(WsSeg &S, uint4 *__restrict__ rec, const float4 *__restrict__ boxes, int n,
                                                 float thr, unsigned *kept_count_out)
{
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int nu = (n + 31) >> 5;
    unsigned long long mykey[WS_U];
    // 1. records, then boxes (all loads of a phase in flight together)
    {
        uint4 r[WS_U];
#pragma unroll
        for (int u = 0; u < WS_U; ++u)
            if (u < nu) { const int e = lane + 32 * u; r[u] = (e < n) ? rec[e] : make_uint4(0u, 0u, 0u, 0u); }
        float4 bx[WS_U];
#pragma unroll
        for (int u = 0; u < WS_U; ++u)
            if (u < nu) { const int e = lane + 32 * u; if (e < n) bx[u] = boxes[r[u].y]; }
#pragma unroll
        for (int u = 0; u < WS_U; ++u) {
            mykey[u] = ~0ull;
            if (u < nu) {
                const int e = lane + 32 * u;
                if (e < n) {
                    mykey[u] = ((unsigned long long)score_desc_bits(r[u].x) << 32) | (unsigned)(~r[u].y);
                    S.key[e] = mykey[u];
                    S.urow[e] = r[u].y; S.uconf[e] = r[u].z;
                    S.ubox[e] = POS ? sanitise(bx[u]) : bx[u];
                }
            }
        }
        if (lane == 0 && (n & 1)) S.key[n] = ~0ull;              // rank_of reads keys two at a time
    }
    __syncwarp();
    // 2. rank
    int rank[WS_U];
#pragma unroll
    for (int u = 0; u < WS_U; ++u) rank[u] = (u < nu) ? rank_of(S.key, n, mykey[u]) : 0;
    __syncwarp();                                                // key storage is reused below
    // 3. sorted prefilter corners + permutation
#pragma unroll
    for (int u = 0; u < WS_U; ++u)
        if (u < nu) {
            const int e = lane + 32 * u;
            if (e < n) { S.q[rank[u]] = quantise_box(S.ubox[e]); S.perm[rank[u]] = (unsigned short)e; }
        }
    __syncwarp();
    // 4. blocks of 32 in score order
    unsigned keptw = 0u;                                         // lane w owns the kept bits of block w
    int nk = 0;
    for (int jb = 0; jb < nu; ++jb) {
        const int j = 32 * jb + lane;
        const bool have = j < n;
        const int ej = have ? (int)S.perm[j] : 0;
        const uint2 qj = have ? S.q[j] : make_uint2(0x7FFF7FFFu, 0u);
        const float4 bj = S.ubox[ej];
        const float aj = box_area(bj);
        bool supp = false;
        for (int ib = 0; ib < jb; ++ib) {
            const unsigned kw = __shfl_sync(FULL, keptw, ib);
            if (kw == 0u) continue;                              // warp-uniform
            unsigned cand;
            if (POS) {
                cand = 0u;
#pragma unroll
                for (int ii = 0; ii < 32; ++ii) cand |= may_overlap(qj, S.q[32 * ib + ii]) ? (1u << ii) : 0u;
                cand &= kw;
            } else {
                cand = kw;
            }
            if (!have || supp) cand = 0u;
            while (cand) {
                const int ii = __ffs(cand) - 1;
                cand &= cand - 1;
                const float4 bi = S.ubox[S.perm[32 * ib + ii]];
                const bool sup = POS ? suppresses_pos(bj, aj, bi, box_area(bi), thr) : suppresses_any(bj, aj, bi, box_area(bi), thr);
                if (sup) { supp = true; cand = 0u; }
            }
        }
        // diagonal block: candidates i < j of the same block
        const int nvalid = min(32, n - 32 * jb);
        unsigned cand;
        if (POS) {
            cand = 0u;
#pragma unroll
            for (int ii = 0; ii < 32; ++ii) cand |= may_overlap(qj, S.q[32 * jb + ii]) ? (1u << ii) : 0u;
        } else {
            cand = FULL;
        }
        cand &= ((1u << lane) - 1u) & (nvalid == 32 ? FULL : ((1u << nvalid) - 1u));
        if (!have || supp) cand = 0u;
        unsigned word = 0u;
        while (cand) {
            const int ii = __ffs(cand) - 1;
            cand &= cand - 1;
            const float4 bi = S.ubox[S.perm[32 * jb + ii]];
            const bool sup = POS ? suppresses_pos(bj, aj, bi, box_area(bi), thr) : suppresses_any(bj, aj, bi, box_area(bi), thr);
            word |= sup ? (1u << ii) : 0u;
        }
        unsigned removed = __ballot_sync(FULL, supp || !have);
        unsigned keep;
        if (!__any_sync(FULL, word != 0u)) {
            keep = ~removed;                                     // nothing inside the block suppresses anything
        } else {
            keep = 0u;
            for (int i = 0; i < nvalid; ++i) {                   // warp-uniform serial resolve (utils.py:67-84)
                const unsigned hit = __ballot_sync(FULL, (word >> i) & 1u);   // lanes that box i suppresses
                if (!((removed >> i) & 1u)) { keep |= 1u << i; removed |= hit; }
            }
        }
        if (lane == jb) keptw = keep;
        if ((keep >> lane) & 1u) rec[nk + __popc(keep & ((1u << lane) - 1u))] = make_uint4(S.urow[ej], S.uconf[ej], 0u, 0u);
        nk += __popc(keep);
    }
    if (lane == 0) *kept_count_out = (unsigned)nk;
    __syncwarp();
}

__global__ void __launch_bounds__(WS_WARPS * 32)
k_segment_nms_warp(uint4 *__restrict__ cand, const unsigned *__restrict__ seg_count, unsigned *__restrict__ kept_count,
                   const float4 *__restrict__ boxtab, long M, int C, int cap_seg, float thr, int seg_first, int nseg,
                   unsigned *__restrict__ big_count, unsigned *__restrict__ big_list)
{
    __shared__ WsSeg sh[WS_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sl = blockIdx.x * WS_WARPS + warp;
    if (sl >= nseg) return;
    const int seg = seg_first + sl;
    const unsigned cnt = seg_count[seg];
    if (cnt == 0u) return;                                      // kept_count was zeroed by yl_post_reset
    if (cnt > (unsigned)cap_seg) return;                        // overflow: reported through meta[], caller re-runs
    if (cnt > (unsigned)SMALL_R) {
        if (lane == 0) big_list[atomicAdd(big_count, 1u)] = (unsigned)seg;
        return;
    }
    uint4 *rec = cand + (size_t)seg * cap_seg;
    const float4 *boxes = boxtab + (size_t)(seg / C) * M;
    if (thr > 0.0f) warp_segment_nms<true>(sh[warp], rec, boxes, (int)cnt, thr, &kept_count[seg]);
    else warp_segment_nms<false>(sh[warp], rec, boxes, (int)cnt, thr, &kept_count[seg]);
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
    // rows of a segment are consecutive in the output: gather 128 rows into shared memory, then store them as one
    // contiguous run of 896 floats (coalesced) instead of seven 28-byte-strided scalar stores per thread
    __shared__ float sh_rows[GATHER_THREADS * 7];
    const unsigned nwr = (unsigned)min((long)nk, max(0L, cap_out - (long)pre));      // rows that fit the output capacity
    for (unsigned q0 = 0; q0 < nwr; q0 += GATHER_THREADS) {
        const unsigned q = q0 + tid;
        if (q < nwr) {
            const uint4 e = rec[q];
            const size_t brow = (size_t)b * M + e.x;
            const float4 bx = boxtab[brow];
            float *o = sh_rows + tid * 7;
            o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
            o[4] = objtab[brow]; o[5] = __uint_as_float(e.y); o[6] = fc;
        }
        __syncthreads();
        const unsigned nrun = min((unsigned)GATHER_THREADS, nwr - q0) * 7u;
        float *dst = out_rows + ((size_t)b * cap_out + pre + q0) * 7;
        for (unsigned i = tid; i < nrun; i += GATHER_THREADS) dst[i] = sh_rows[i];
        __syncthreads();
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
    // block-per-segment form by default (118 us vs 131 us at B=64, conf 1e-4); YL_NMS_WARP=1 selects warp-per-segment
    static const bool cta_tier = !(getenv("YL_NMS_WARP") && getenv("YL_NMS_WARP")[0] == '1');
    if (cta_tier)
        k_segment_nms_small<<<nseg, SMALL_THREADS, 0, (cudaStream_t)stream>>>(cand, seg_count, kept_count, boxtab, L.M4, C, cap_seg,
                                                                            nms_thre, seg_first, big_count, big_list);
    else
        k_segment_nms_warp<<<(nseg + WS_WARPS - 1) / WS_WARPS, WS_WARPS * 32, 0, (cudaStream_t)stream>>>(
            cand, seg_count, kept_count, boxtab, L.M4, C, cap_seg, nms_thre, seg_first, nseg, big_count, big_list);
    YL_LAUNCH_CHECK();
    if (cap_seg > SMALL_R) {
        const int grid_big = nseg < 148 * 2 ? nseg : 148 * 2;
        k_segment_nms_big<<<grid_big, NMS_THREADS, 0, (cudaStream_t)stream>>>(cand, seg_count, kept_count, boxtab, L.M4, C, cap_seg,
                                                                             nms_thre, big_count, big_list, kept_scratch);
        YL_LAUNCH_CHECK();
    }
    k_gather_rows<<<nseg, GATHER_THREADS, 0, (cudaStream_t)stream>>>(cand, seg_count, kept_count, boxtab, objtab, L.M4, C,
                                                                   cap_seg, B, out_rows, cap_out, meta, seg_first);
    YL_LAUNCH_CHECK();
    return YL_OK;
}
