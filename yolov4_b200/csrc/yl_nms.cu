// yl_nms.cu -- back end of the postprocess pipeline.
//
//   k_segment_nms_bins  one CTA per (image,class) candidate segment of up to SMALL_R records: sort by
//                       (score desc, box row desc) [utils.py:58 with stable tie order], greedy NMS dropping a box when
//                       IoU >= thr with any already kept box [utils.py:67-84]; the kept detections are written as
//                       finished 7-float rows at the front of the segment.
//   k_segment_nms_big   persistent CTAs over the queue of segments the small tier passes on (more than SMALL_R
//                       candidates, dense overlap graphs, thr <= 0): bitonic sort + 64-wide chunked bitmask NMS in
//                       shared memory (<= SMEM_R) or in place in global memory (degenerate all-ties inputs).
//   k_gather_rows       class-ascending concatenation of the kept rows of an image [utils.py:191-220] into
//                       out_rows[b] = (x1,y1,x2,y2,obj,cls_conf,cls) and the per-image counts.
//
// Candidate record (written by the filter kernels, yl_filter.cu): two uint4,
//   A = {score bits, box row, cls_conf bits, obj_conf bits}    B = {x1, y1, x2, y2}
// so a segment is self-contained and the back end reads nothing but its own 32-byte records.
#include <math.h>
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

__device__ __forceinline__ float4 as_float4(const uint4 &u)
{
    return make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
}

// One finished output row (utils.py:177-184): (x1, y1, x2, y2, obj_conf, cls_conf, cls_idx as float).
__device__ __forceinline__ void store_row(float *o, const float4 &box, unsigned conf_bits, unsigned obj_bits, float fc)
{
    o[0] = box.x; o[1] = box.y; o[2] = box.z; o[3] = box.w;
    o[4] = __uint_as_float(obj_bits); o[5] = __uint_as_float(conf_bits); o[6] = fc;
}

// ---------------------------------------------------------------------------------------------------------------
// Small tier (the common case: ~150 candidates per (image,class) at conf 1e-4).  Everything in ~21 KB of shared memory.
//   1. rank sort on the 32-bit score key (LDS.128 per four keys, one compare + predicated add per key); if two records
//      tie on the score the ranks collide, which is detected, and the segment is ranked again on the unique 64-bit key
//      (score, row);
//   2. candidate pairs by table lookup instead of an all-pairs loop: boxes are binned on x1, x2, y1, y2 (NBIN bins over
//      the segment's coordinate range) and on log2(area) (4 bins per octave); six prefix tables hold, per bin, the set
//      (bitmask over sorted positions) of boxes that CANNOT overlap / reach the IoU threshold with a box of that bin:
//         x1_i > x2_j | x2_i < x1_j | y1_i > y2_j | y2_i < y1_j | area_i >> area_j | area_i << area_j
//      so the candidate set of row j is one 6-way OR per 64 earlier boxes.  Binning is monotone, hence conservative: a
//      pair that overlaps (tl < br on both axes) and whose area ratio allows IoU >= thr is never excluded;
//   3. the surviving (i, j) pairs (~2 % of all pairs) are queued, then tested exactly by all threads (no divergence);
//   4. pairs with IoU >= thr set a bit in the transposed suppression matrix T[j]; one warp walks only the rows that
//      have such a bit, in score order: kept[j] = (T[j] & kept) == 0                       (utils.py:67-84);
//   5. the kept detections are staged in shared memory and stored as one contiguous run of 7-float rows.
// Segments with more than SMALL_R candidates, with more queued pairs than the queue holds, or with thr <= 0 (where
// every pair must go through the literal NaN-propagating formula) are passed on to the big tier.
// ---------------------------------------------------------------------------------------------------------------
constexpr int SMALL_R = 256;
constexpr int SMALL_W = SMALL_R / 32;
#ifndef YL_NMS_THREADS
#define YL_NMS_THREADS 128
#endif
constexpr int SMALL_THREADS = YL_NMS_THREADS;          // a multiple of 32
constexpr int SMALL_WARPS = SMALL_THREADS / 32;
constexpr int SMALL_EPT = (SMALL_R + SMALL_THREADS - 1) / SMALL_THREADS;      // records per thread, at most
#ifndef YL_NMS_NBIN
#define YL_NMS_NBIN 48
#endif
constexpr int NBIN = YL_NMS_NBIN;                     // bins per table (at most 64: bin numbers are packed in 6 bits)
#ifndef YL_NMS_QCAP
#define YL_NMS_QCAP 512
#endif
constexpr int QCAP_W = YL_NMS_QCAP;                     // queued pairs per warp
#ifndef YL_NMS_MINB
#define YL_NMS_MINB 10
#endif
#ifndef YL_NMS_SPLIT_RANK
#define YL_NMS_SPLIT_RANK 1
#endif
constexpr int AREA_BIN_OFF = (127 + 4) << 2;            // float bits >> 21 of 16.0f: areas below 16 px^2 share bin 0

struct alignas(16) BinsSmem {
    union {
        struct {
            unsigned key[SMALL_R + 4];                  // score key per record (record order), padded for the 4-wide loop
            unsigned pad_[4];
            unsigned long long k64[SMALL_R + 2];        // tie fallback: (score key, ~row)
        } s;
        unsigned short queue[SMALL_WARPS][QCAP_W];      // (j << 8) | i candidate pairs, one queue per warp
    } a;
    float4 box[SMALL_R];                                // sorted
    uint2 co[SMALL_R];                                  // sorted {cls_conf bits, obj_conf bits}
    unsigned bins[SMALL_R];                             // sorted: bx1 | bx2<<6 | by1<<12 | by2<<18 | barea<<24 | invalid<<31
    union {
        unsigned long long tab[6][SMALL_W / 2][NBIN + 1];   // exclusion sets per bin, two 32-box words per entry (+1: conflict-free scan)
        unsigned T[SMALL_R][SMALL_W];                   // transposed suppression matrix
        float stage[SMALL_R * 7];                       // finished rows
    } b;
    unsigned char owner[SMALL_R];
    float red[2][SMALL_WARPS];
    unsigned qn[SMALL_WARPS];
    unsigned dirty[SMALL_W], keptw[SMALL_W], pref[SMALL_W + 1];
};

// Number of keys below `key`; keys beyond n are padded with 0xFFFFFFFF up to a multiple of four.
__device__ __forceinline__ int rank32(const unsigned *sh_key, int n, unsigned key)
{
    int r0 = 0, r1 = 0, r2 = 0, r3 = 0;                        // four independent counters: no serial chain through the adds
    const uint4 *k4 = reinterpret_cast<const uint4 *>(sh_key);
    const int n4 = (n + 3) >> 2;
#pragma unroll 4
    for (int j = 0; j < n4; ++j) {
        const uint4 k = k4[j];
        asm("{\n\t.reg .pred p, q, r, s;\n\t"
            "setp.lt.u32 p, %4, %8;\n\tsetp.lt.u32 q, %5, %8;\n\tsetp.lt.u32 r, %6, %8;\n\tsetp.lt.u32 s, %7, %8;\n\t"
            "@p add.s32 %0, %0, 1;\n\t@q add.s32 %1, %1, 1;\n\t@r add.s32 %2, %2, 1;\n\t@s add.s32 %3, %3, 1;\n\t}"
            : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3) : "r"(k.x), "r"(k.y), "r"(k.z), "r"(k.w), "r"(key));
    }
    return (r0 + r1) + (r2 + r3);
}

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

__device__ __forceinline__ int coord_bin(float v, float lo, float scale)
{
    // monotone in v: fp subtraction, multiplication by a non-negative constant, floor and clamping all are
    const int q = __float2int_rd(__fmul_rn(__fsub_rn(v, lo), scale));      // saturates on +-inf, 0 on NaN
    return min(max(q, 0), NBIN - 1);
}

__global__ void __launch_bounds__(SMALL_THREADS, YL_NMS_MINB)
k_segment_nms_bins(uint4 *__restrict__ cand, const unsigned *__restrict__ seg_count, unsigned *__restrict__ kept_count,
                   int C, int cap_seg, float thr, int area_k, int seg_first,
                   unsigned *__restrict__ big_count, unsigned *__restrict__ big_list)
{
    __shared__ BinsSmem S;
    const unsigned FULL = 0xFFFFFFFFu;
    const int seg = seg_first + blockIdx.x;
    pdl_trigger();
    {
        // zero the exclusion tables: independent of the input, so a CTA launched early (PDL) does it while the filter
        // kernel is still draining, and otherwise while the record loads are in flight
        uint4 *z = reinterpret_cast<uint4 *>(&S.b.tab[0][0][0]);
        constexpr int NZ = (int)(sizeof(S.b.tab) / sizeof(uint4));
        for (int i = threadIdx.x; i < NZ; i += SMALL_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
        if (threadIdx.x < SMALL_WARPS) S.qn[threadIdx.x] = 0u;
        if (threadIdx.x < SMALL_W) S.dirty[threadIdx.x] = 0u;
    }
    pdl_wait();                                                 // segment counts and records of the filter kernel
    const unsigned cnt = seg_count[seg];
    if (cnt == 0u) return;                                      // kept_count was zeroed by yl_post_reset
    if (cnt > (unsigned)cap_seg) return;                        // overflow: reported through meta[], caller re-runs
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (cnt > (unsigned)SMALL_R || !(thr > 0.0f)) {
        if (tid == 0) big_list[atomicAdd(big_count, 1u)] = (unsigned)seg;
        return;
    }
    const int n = (int)cnt;
    const int nw = (n + 31) >> 5, nwp = (nw + 1) >> 1;
    uint4 *rec = cand + (size_t)seg * cap_seg * 2;

    // ---- 1. records -> registers; score keys -> shared memory; coordinate range of the segment ----
    uint4 ra[SMALL_EPT];
    float4 rb[SMALL_EPT];
#pragma unroll
    for (int u = 0; u < SMALL_EPT; ++u) {
        const int e = tid + u * SMALL_THREADS;
        ra[u] = make_uint4(0u, 0u, 0u, 0u);
        rb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < n) { ra[u] = rec[2 * e]; rb[u] = as_float4(rec[2 * e + 1]); }
    }
    unsigned key[SMALL_EPT];
    bool valid[SMALL_EPT];
    float lo = kInf, hi = -kInf;
#pragma unroll
    for (int u = 0; u < SMALL_EPT; ++u) {
        const int e = tid + u * SMALL_THREADS;
        key[u] = 0xFFFFFFFFu;
        valid[u] = false;
        if (e < n) {
            key[u] = score_desc_bits(ra[u].x);
            S.a.s.key[e] = key[u];
            // only boxes with positive extents (hence no NaN) can overlap anything (suppresses_pos: tl < br on both axes)
            valid[u] = rb[u].x < rb[u].z && rb[u].y < rb[u].w;
            if (valid[u]) { lo = fminf(lo, fminf(rb[u].x, rb[u].y)); hi = fmaxf(hi, fmaxf(rb[u].z, rb[u].w)); }
        }
    }
    if (tid < 4) S.a.s.key[n + tid] = 0xFFFFFFFFu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(FULL, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(FULL, hi, o));
    }
    if (lane == 0) { S.red[0][warp] = lo; S.red[1][warp] = hi; }
    __syncthreads();

    // ---- 2. rank ----
    const int wbase = tid & ~31;                                // warp-uniform bounds below
    int rank[SMALL_EPT];
#pragma unroll
    for (int u = 0; u < SMALL_EPT; ++u) rank[u] = 0;
    if (wbase < n) rank[0] = rank32(S.a.s.key, n, key[0]);
#if YL_NMS_SPLIT_RANK
    static_assert(SMALL_EPT == 2, "the split second pass assumes two records per thread");
    if (n > SMALL_THREADS) {                                    // CTA-uniform
        // Records SMALL_THREADS.. (a few dozen in the typical segment of ~144) would all sit in warp 0 and make it walk
        // every key a second time while the other warps wait at the barrier.  Instead every warp ranks them against its own
        // quarter of the keys and the partial counts are added up in shared memory.
        const int x = n - SMALL_THREADS;
        S.bins[tid] = 0u;                                       // S.bins is free until the sorted arrays are built
        __syncthreads();
        const int n4 = (n + 3) >> 2;
        const int g0 = (n4 * warp) / SMALL_WARPS, g1 = (n4 * (warp + 1)) / SMALL_WARPS;
        for (int c = 0; 32 * c < x; ++c) {
            const int i = 32 * c + lane;
            if (i < x) {
                const int part = rank32(S.a.s.key + 4 * g0, 4 * (g1 - g0), S.a.s.key[SMALL_THREADS + i]);
                atomicAdd(&S.bins[i], (unsigned)part);
            }
        }
        __syncthreads();
        if (tid < x) rank[1] = (int)S.bins[tid];
    }
#else
#pragma unroll
    for (int u = 1; u < SMALL_EPT; ++u)
        if (wbase + u * SMALL_THREADS < n) rank[u] = rank32(S.a.s.key, n, key[u]);
#endif
#pragma unroll
    for (int u = 0; u < SMALL_EPT; ++u) {
        const int e = tid + u * SMALL_THREADS;
        if (e < n) S.owner[rank[u]] = (unsigned char)e;
    }
    __syncthreads();
    {
        bool coll = false;
#pragma unroll
        for (int u = 0; u < SMALL_EPT; ++u) {
            const int e = tid + u * SMALL_THREADS;
            if (e < n) coll |= (S.owner[rank[u]] != (unsigned char)e);
        }
        if (__syncthreads_or(coll)) {
            // equal scores inside the segment: the order is (score desc, row desc), decided on the unique 64-bit key
            unsigned long long k64[SMALL_EPT];
#pragma unroll
            for (int u = 0; u < SMALL_EPT; ++u) {
                const int e = tid + u * SMALL_THREADS;
                k64[u] = ~0ull;
                if (e < n) { k64[u] = ((unsigned long long)key[u] << 32) | (unsigned)(~ra[u].y); S.a.s.k64[e] = k64[u]; }
            }
            if (tid == 0 && (n & 1)) S.a.s.k64[n] = ~0ull;
            __syncthreads();
#pragma unroll
            for (int u = 0; u < SMALL_EPT; ++u)
                if (wbase + u * SMALL_THREADS < n) rank[u] = rank_of(S.a.s.k64, n, k64[u]);
        }
    }

    // ---- 3. sorted arrays, bins, table inserts ----
    {
        lo = S.red[0][0]; hi = S.red[1][0];
#pragma unroll
        for (int wq = 1; wq < SMALL_WARPS; ++wq) { lo = fminf(lo, S.red[0][wq]); hi = fmaxf(hi, S.red[1][wq]); }
        lo = fmaxf(lo, -1.0e6f);
        hi = fminf(hi, 1.0e6f);
        float scale = __fdiv_rn((float)NBIN, __fsub_rn(hi, lo));
        if (!(scale > 0.0f && scale < 3.0e38f)) scale = 0.0f;  // empty / degenerate range: one bin, nothing is excluded
#pragma unroll
        for (int u = 0; u < SMALL_EPT; ++u) {
            const int e = tid + u * SMALL_THREADS;
            if (e < n) {
                const int r = rank[u];
                S.box[r] = rb[u];
                S.co[r] = make_uint2(ra[u].z, ra[u].w);
                const unsigned bit = 1u << (r & 31);
                const int wp = r >> 6, half = (r >> 5) & 1;
                unsigned bn = 0x80000000u;
                if (valid[u]) {
                    const int bx1 = coord_bin(rb[u].x, lo, scale), bx2 = coord_bin(rb[u].z, lo, scale);
                    const int by1 = coord_bin(rb[u].y, lo, scale), by2 = coord_bin(rb[u].w, lo, scale);
                    const int ab = min(max((int)(__float_as_uint(box_area(rb[u])) >> 21) - AREA_BIN_OFF, 0), NBIN - 1);
                    bn = (unsigned)bx1 | ((unsigned)bx2 << 6) | ((unsigned)by1 << 12) | ((unsigned)by2 << 18) | ((unsigned)ab << 24);
                    // tab[0][.][b] (suffix-OR) = { i : bx1_i - 1 >= b }  -> lookup at bx2_j gives x1_i > x2_j (in bins)
                    // tab[1][.][b] (prefix-OR) = { i : bx2_i + 1 <= b }  -> lookup at bx1_j gives x2_i < x1_j
                    if (bx1 >= 1) atomicOr(reinterpret_cast<unsigned *>(&S.b.tab[0][wp][bx1 - 1]) + half, bit);
                    if (bx2 + 1 < NBIN) atomicOr(reinterpret_cast<unsigned *>(&S.b.tab[1][wp][bx2 + 1]) + half, bit);
                    if (by1 >= 1) atomicOr(reinterpret_cast<unsigned *>(&S.b.tab[2][wp][by1 - 1]) + half, bit);
                    if (by2 + 1 < NBIN) atomicOr(reinterpret_cast<unsigned *>(&S.b.tab[3][wp][by2 + 1]) + half, bit);
                    // tab[4][.][b] (suffix-OR) = { i : ab_i - K >= b }   -> lookup at ab_j gives area_i too large for IoU >= thr
                    // tab[5][.][b] (prefix-OR) = { i : ab_i + K <= b }   -> lookup at ab_j gives area_i too small
                    if (ab - area_k >= 0) atomicOr(reinterpret_cast<unsigned *>(&S.b.tab[4][wp][ab - area_k]) + half, bit);
                    if (ab + area_k < NBIN) atomicOr(reinterpret_cast<unsigned *>(&S.b.tab[5][wp][ab + area_k]) + half, bit);
                } else {
                    atomicOr(reinterpret_cast<unsigned *>(&S.b.tab[1][wp][0]) + half, bit);     // in every prefix set: never a candidate
                }
                S.bins[r] = bn;
            }
        }
    }
    __syncthreads();
    // prefix / suffix OR over the bins of 6 tables x nwp word pairs.  Eight lanes share a chain: each ORs its own run of
    // bins (independent loads), a 3-step shuffle scan over the eight run totals gives every lane its carry, and the runs
    // are stored back -- about 15 dependent steps where one lane per chain walked NBIN of them while three warps waited.
    {
        constexpr int BPL = (NBIN + 7) / 8;                     // bins per lane
        const int nchain = 6 * nwp;
        const int sub = tid & 7;
        for (int c0 = 0; c0 < nchain; c0 += SMALL_THREADS / 8) {    // CTA-uniform trip count: the shuffles run with a full mask
            const int c = c0 + (tid >> 3);
            const bool valid = c < nchain;
            const int t = valid ? c / nwp : 0, wp = valid ? c - t * nwp : 0;
            unsigned long long *p = &S.b.tab[t][wp][0];
            const bool fwd = (t & 1) != 0;                      // odd tables are prefix sets, even ones suffix sets
            unsigned long long v[BPL];
#pragma unroll
            for (int k = 0; k < BPL; ++k) {
                const int i = BPL * sub + k;                    // position along the scan direction
                v[k] = (valid && i < NBIN) ? p[fwd ? i : NBIN - 1 - i] : 0ull;
            }
#pragma unroll
            for (int k = 1; k < BPL; ++k) v[k] |= v[k - 1];
            unsigned long long incl = v[BPL - 1];
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                const unsigned long long y = __shfl_up_sync(FULL, incl, o, 8);
                if (sub >= o) incl |= y;
            }
            unsigned long long carry = __shfl_up_sync(FULL, incl, 1, 8);
            if (sub == 0) carry = 0ull;
#pragma unroll
            for (int k = 0; k < BPL; ++k) {
                const int i = BPL * sub + k;
                if (valid && i < NBIN) p[fwd ? i : NBIN - 1 - i] = v[k] | carry;
            }
        }
    }
    __syncthreads();

    // ---- 4. candidate pairs (i < j) -> this warp's queue ----
    unsigned qcount = 0u;
    bool overflow = false;
    {
        unsigned short *myq = S.a.queue[warp];
        const unsigned lt = (1u << lane) - 1u;
#pragma unroll
        for (int u = 0; u < SMALL_EPT; ++u) {
            const int j0 = wbase + u * SMALL_THREADS;           // first row of this warp's block of 32 (warp-uniform)
            if (j0 < n) {
                const int j = j0 + lane;
                const unsigned bn = (j < n) ? S.bins[j] : 0x80000000u;
                const bool act = !(bn >> 31);
                const int bx1 = bn & 63, bx2 = (bn >> 6) & 63, by1 = (bn >> 12) & 63, by2 = (bn >> 18) & 63, ab = (bn >> 24) & 63;
                const int wpj = j0 >> 6;
                for (int wp = 0; wp <= wpj; ++wp) {
                    const unsigned long long ex = S.b.tab[0][wp][bx2] | S.b.tab[1][wp][bx1] | S.b.tab[2][wp][by2] |
                                                  S.b.tab[3][wp][by1] | S.b.tab[4][wp][ab] | S.b.tab[5][wp][ab];
                    const int rel = j - 64 * wp;                 // rows of this warp lie in one 64-block: rel >= 0
                    const unsigned long long lim = (rel >= 64) ? ~0ull : ((1ull << rel) - 1ull);
                    const unsigned long long cw = act ? (~ex & lim) : 0ull;
                    unsigned c0 = (unsigned)cw, c1 = (unsigned)(cw >> 32);
                    const int ibase = 64 * wp;
                    while (__any_sync(FULL, (c0 | c1) != 0u)) {  // one pair per lane and round, low word first
                        const bool has = (c0 | c1) != 0u;
                        const unsigned m = __ballot_sync(FULL, has);
                        if (has) {
                            int ii;
                            if (c0) { ii = __ffs(c0) - 1; c0 &= c0 - 1u; }
                            else { ii = 32 + __ffs(c1) - 1; c1 &= c1 - 1u; }
                            const unsigned slot = qcount + __popc(m & lt);
                            if (slot < (unsigned)QCAP_W) myq[slot] = (unsigned short)((j << 8) | (ibase + ii));
                            else overflow = true;
                        }
                        qcount += __popc(m);
                    }
                }
            }
        }
        if (lane == 0) S.qn[warp] = min(qcount, (unsigned)QCAP_W);
    }
    if (__syncthreads_or(overflow)) {
        // dense overlap graph: more candidate pairs than the queue holds -- the chunked big tier takes the segment
        if (tid == 0) big_list[atomicAdd(big_count, 1u)] = (unsigned)seg;
        return;
    }
    unsigned qtotal = 0u;
#pragma unroll
    for (int wq = 0; wq < SMALL_WARPS; ++wq) qtotal += S.qn[wq];

    // ---- 5. exact tests on the queued pairs, suppression matrix, resolve ----
    bool any_edges = false;
    if (qtotal != 0u) {                                          // CTA-uniform
        uint4 *z = reinterpret_cast<uint4 *>(&S.b.T[0][0]);
        for (int i = tid; i < n * (SMALL_W / 4); i += SMALL_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        bool found = false;
#pragma unroll 1
        for (int wq = 0; wq < SMALL_WARPS; ++wq) {
            const unsigned cq = S.qn[wq];
            for (unsigned q = tid; q < cq; q += SMALL_THREADS) {
                const unsigned e = S.a.queue[wq][q];
                const int j = e >> 8, i = e & 255;
                const float4 bj = S.box[j], bi = S.box[i];
                if (suppresses_pos(bj, box_area(bj), bi, box_area(bi), thr)) {
                    atomicOr(&S.b.T[j][i >> 5], 1u << (i & 31));
                    atomicOr(&S.dirty[j >> 5], 1u << (j & 31));
                    found = true;
                }
            }
        }
        any_edges = __syncthreads_or(found) != 0;
    }
    if (tid < 32) {
        // lane w owns kept word w; only rows with a suppressor are visited, in score order (utils.py:67-84)
        unsigned keptw = 0u;
        if (lane < nw) keptw = (lane == nw - 1 && (n & 31)) ? ((1u << (n & 31)) - 1u) : FULL;
        if (any_edges) {
            for (int w = 0; w < nw; ++w) {
                unsigned d = S.dirty[w];
                while (d) {
                    const int jj = __ffs(d) - 1;
                    d &= d - 1u;
                    const unsigned t = (lane <= w) ? S.b.T[32 * w + jj][lane & (SMALL_W - 1)] : 0u;
                    if (__any_sync(FULL, (t & keptw) != 0u) && lane == w) keptw &= ~(1u << jj);
                }
            }
        }
        int c = __popc(keptw), incl = c;
#pragma unroll
        for (int o = 1; o < SMALL_W; o <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane < SMALL_W) { S.keptw[lane] = keptw; S.pref[lane] = (unsigned)(incl - c); }
        if (lane == SMALL_W - 1) { S.pref[SMALL_W] = (unsigned)incl; kept_count[seg] = (unsigned)incl; }
    }
    __syncthreads();

    // ---- 6. finished rows: stage in shared memory (the tables / T are dead), store as one contiguous run ----
    const float fc = (float)(seg % C);                           // utils.py:183 class id stored as float
    for (int j = tid; j < n; j += SMALL_THREADS) {
        const unsigned kw = S.keptw[j >> 5];
        if ((kw >> (j & 31)) & 1u) {
            const unsigned q = S.pref[j >> 5] + __popc(kw & ((1u << (j & 31)) - 1u));
            const uint2 co = S.co[j];
            store_row(S.b.stage + q * 7, S.box[j], co.x, co.y, fc);
        }
    }
    __syncthreads();
    {
        const int nf = (int)S.pref[SMALL_W] * 7;
        float4 *dst4 = reinterpret_cast<float4 *>(rec);
        const float4 *src4 = reinterpret_cast<const float4 *>(S.b.stage);
        for (int i = tid; i < (nf >> 2); i += SMALL_THREADS) dst4[i] = src4[i];
        const int tail = nf & ~3;
        if (tid < (nf & 3)) reinterpret_cast<float *>(rec)[tail + tid] = S.b.stage[tail + tid];
    }
}

// ---- big tier ---------------------------------------------------------------------------------------------
struct SmemStore {
    unsigned long long *key;   // ascending sort key: (~score order bits) << 32 | ~row
    unsigned short *idx;       // sorted position -> record index
    const float4 *ubox;        // record order, original boxes
    bool pos;
    __device__ __forceinline__ void cswap(int i, int j)
    {
        const unsigned long long a = key[i], b = key[j];
        if (b < a) { key[i] = b; key[j] = a; const unsigned short t = idx[i]; idx[i] = idx[j]; idx[j] = t; }
    }
    __device__ __forceinline__ float4 get_box(int i) const { const float4 b = ubox[idx[i]]; return pos ? sanitise(b) : b; }
    __device__ __forceinline__ float get_area(int i) const { return box_area(ubox[idx[i]]); }
};

struct GlobalStore {
    uint4 *rec;                // record i: rec[2i] = {key hi, key lo, conf, obj}, rec[2i+1] = box
    bool pos;
    __device__ __forceinline__ void cswap(int i, int j)
    {
        const uint4 a = rec[2 * i], b = rec[2 * j];
        const unsigned long long ka = ((unsigned long long)a.x << 32) | a.y, kb = ((unsigned long long)b.x << 32) | b.y;
        if (kb < ka) {
            const uint4 ab = rec[2 * i + 1], bb = rec[2 * j + 1];
            rec[2 * i] = b; rec[2 * j] = a; rec[2 * i + 1] = bb; rec[2 * j + 1] = ab;
        }
    }
    __device__ __forceinline__ float4 get_box(int i) const
    {
        const float4 b = as_float4(rec[2 * i + 1]);
        return pos ? sanitise(b) : b;
    }
    __device__ __forceinline__ float get_area(int i) const { return box_area(as_float4(rec[2 * i + 1])); }
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

// Persistent CTAs walk the list of segments the small tier passed on.
__global__ void __launch_bounds__(NMS_THREADS)
k_segment_nms_big(uint4 *__restrict__ cand, const unsigned *__restrict__ seg_count, unsigned *__restrict__ kept_count,
                  int C, int cap_seg, float thr, const unsigned *__restrict__ big_count, const unsigned *__restrict__ big_list,
                  unsigned *__restrict__ kept_scratch /* [B*C*cap_seg] u32, only when cap_seg > SMEM_R; else null */)
{
    __shared__ unsigned long long sh_key[SMEM_R];
    __shared__ unsigned short sh_idx[SMEM_R];
    __shared__ float4 sh_ubox[SMEM_R];
    __shared__ uint2 sh_uco[SMEM_R];
    __shared__ unsigned short sh_kept[SMEM_R];
    __shared__ unsigned long long sh_rows[CHUNK];
    __shared__ unsigned sh_supp[2];
    __shared__ int sh_nk;

    pdl_trigger();
    pdl_wait();                                                 // the queue filled by k_segment_nms_bins
    const unsigned nbig = *big_count;
    for (unsigned it = blockIdx.x; it < nbig; it += gridDim.x) {
        const int seg = (int)big_list[it];
        const int n = (int)seg_count[seg];
        uint4 *rec = cand + (size_t)seg * cap_seg * 2;
        float *rows = reinterpret_cast<float *>(rec);
        const float fc = (float)(seg % C);
        const int tid = threadIdx.x;
        const bool pos = thr > 0.0f;

        if (n <= SMEM_R) {
            for (int i = tid; i < n; i += NMS_THREADS) {
                const uint4 a = rec[2 * i];
                sh_key[i] = ((unsigned long long)score_desc_bits(a.x) << 32) | (unsigned)(~a.y);
                sh_idx[i] = (unsigned short)i;
                sh_uco[i] = make_uint2(a.z, a.w);
                sh_ubox[i] = as_float4(rec[2 * i + 1]);
            }
            __syncthreads();
            SmemStore s{sh_key, sh_idx, sh_ubox, pos};
            bitonic_sort(s, n);
            const int nk = pos ? greedy_nms<true>(s, n, thr, sh_kept, nullptr, sh_supp, sh_rows, &sh_nk)
                               : greedy_nms<false>(s, n, thr, sh_kept, nullptr, sh_supp, sh_rows, &sh_nk);
            for (int q = tid; q < nk; q += NMS_THREADS) {
                const int e = sh_idx[sh_kept[q]];
                const uint2 co = sh_uco[e];
                store_row(rows + (size_t)q * 7, sh_ubox[e], co.x, co.y, fc);
            }
            if (tid == 0) kept_count[seg] = (unsigned)nk;
        } else {
            // oversized segment: same algorithm in place in global memory (slow path, correctness only)
            for (int i = tid; i < n; i += NMS_THREADS) {
                const uint4 a = rec[2 * i];
                rec[2 * i] = make_uint4(score_desc_bits(a.x), ~a.y, a.z, a.w);
            }
            __syncthreads();
            GlobalStore s{rec, pos};
            bitonic_sort(s, n);
            unsigned *kept_g = kept_scratch + (size_t)seg * cap_seg;
            const int nk = pos ? greedy_nms<true>(s, n, thr, nullptr, kept_g, sh_supp, sh_rows, &sh_nk)
                               : greedy_nms<false>(s, n, thr, nullptr, kept_g, sh_supp, sh_rows, &sh_nk);
            // rows replace records in place: kept_g[q] >= q and a row (28 B) is shorter than a record (32 B), so the
            // rows of a chunk never reach a record a later chunk still has to read; go through registers chunk by chunk
            for (int q0 = 0; q0 < nk; q0 += NMS_THREADS) {
                const int q = q0 + tid;
                uint4 a = make_uint4(0u, 0u, 0u, 0u), bb = make_uint4(0u, 0u, 0u, 0u);
                if (q < nk) { a = rec[2 * (size_t)kept_g[q]]; bb = rec[2 * (size_t)kept_g[q] + 1]; }
                __syncthreads();
                if (q < nk) store_row(rows + (size_t)q * 7, as_float4(bb), a.z, a.w, fc);
                __syncthreads();
            }
            if (tid == 0) kept_count[seg] = (unsigned)nk;
        }
        __syncthreads();
    }
}

constexpr int GATHER_THREADS = 128;

// The kept rows of a segment are finished and contiguous at the front of the segment: concatenation is a copy.
__global__ void __launch_bounds__(GATHER_THREADS)
k_gather_rows(const uint4 *__restrict__ cand, const unsigned *__restrict__ seg_count, const unsigned *__restrict__ kept_count,
              int C, int cap_seg, int B, float *__restrict__ out_rows, long cap_out, int *__restrict__ meta, int seg_first)
{
    __shared__ unsigned sh_red[3][GATHER_THREADS / 32];
    pdl_wait();                                                 // kept counts and finished rows of both NMS tiers
    const int seg = seg_first + blockIdx.x;
    const int b = seg / C, c = seg - b * C;
    const int tid = threadIdx.x;
    // the first 2 x 128 float4s of the segment's finished rows (4 KB, the typical segment) are requested before the kept
    // counts are known, so the copy's round trip overlaps the prefix's instead of following it
    const float4 *src4 = reinterpret_cast<const float4 *>(cand + (size_t)seg * cap_seg * 2);
    const unsigned lim4 = (unsigned)cap_seg * 2u;
    float4 spec0 = make_float4(0.f, 0.f, 0.f, 0.f), spec1 = spec0;
    if ((unsigned)tid < lim4) spec0 = src4[tid];
    if ((unsigned)tid + GATHER_THREADS < lim4) spec1 = src4[tid + GATHER_THREADS];
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
    const unsigned nwr = (unsigned)min((long)nk, max(0L, cap_out - (long)pre));      // rows that fit the output capacity
    const float *src = reinterpret_cast<const float *>(cand + (size_t)seg * cap_seg * 2);
    float *dst = out_rows + ((size_t)b * cap_out + pre) * 7;
    // the source run is 16-byte aligned, the destination only 4-byte: 128-bit loads, scalar stores
    const unsigned nf = nwr * 7u, nf4 = nf >> 2;
    for (unsigned i = tid; i < nf4; i += GATHER_THREADS) {
        const float4 v = (i == (unsigned)tid) ? spec0 : ((i == (unsigned)tid + GATHER_THREADS) ? spec1 : src4[i]);
        float *d = dst + 4 * i;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    if (tid < (nf & 3u)) dst[4 * nf4 + tid] = src[4 * nf4 + tid];
}

// Smallest K such that two positive areas whose bins (float bits >> 21: four per octave) differ by at least K have a
// ratio above 1/(0.999 thr) * (1 + 1e-3): then IoU <= amin/amax < thr with a margin far above fp32 rounding, which is
// the condition under which suppresses_pos returns false on the area test alone.
static int area_bin_gap(float thr)
{
    if (!(thr > 0.0f)) return NBIN + 1;
    const double need = 1.0 / (0.999 * (double)thr) * 1.001;
    for (int K = 1; K <= NBIN; ++K) {
        double worst = 1e300;
        for (int p = 0; p < 4; ++p) {
            // bins p+1 .. : lower edge of bin (p + K) over upper edge of bin p (= lower edge of bin p + 1)
            const int hi = p + K, lo = p + 1;
            const double Lh = ldexp(1.0 + (hi & 3) * 0.25, hi >> 2), Ll = ldexp(1.0 + (lo & 3) * 0.25, lo >> 2);
            worst = fmin(worst, Lh / Ll);
        }
        if (worst >= need) return K;
    }
    return NBIN + 1;
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
    unsigned *kept_scratch = (cap_seg > SMEM_R) ? (unsigned *)(w + L.off_kept_scratch) : nullptr;
    const int nseg = img_count * C, seg_first = img_first * C;
    unsigned *big_count = (unsigned *)(w + L.off_big_count) + img_first;
    unsigned *big_list = (unsigned *)(w + L.off_big_list) + seg_first;
    cudaStream_t st = (cudaStream_t)stream;
    // each kernel may become resident while its predecessor drains (PDL); it waits before reading the predecessor's output.
    // The first one follows the emit / dense filter kernel of yl_filter_* on the same stream.
    const bool pdl = pdl_enabled();
    YL_CUDA_TRY(launch_after(k_segment_nms_bins, dim3(nseg), dim3(SMALL_THREADS), 0, st, pdl, cand, seg_count, kept_count, C, cap_seg,
                             nms_thre, area_bin_gap(nms_thre), seg_first, big_count, big_list));
    const int grid_big = nseg < 148 * 2 ? nseg : 148 * 2;
    YL_CUDA_TRY(launch_after(k_segment_nms_big, dim3(grid_big), dim3(NMS_THREADS), 0, st, pdl, cand, (const unsigned *)seg_count, kept_count,
                             C, cap_seg, nms_thre, (const unsigned *)big_count, (const unsigned *)big_list, kept_scratch));
    YL_CUDA_TRY(launch_after(k_gather_rows, dim3(nseg), dim3(GATHER_THREADS), 0, st, pdl, (const uint4 *)cand, (const unsigned *)seg_count,
                             (const unsigned *)kept_count, C, cap_seg, B, out_rows, cap_out, meta, seg_first));
    return YL_OK;
}
