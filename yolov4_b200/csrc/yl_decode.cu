// yl_decode.cu -- the contract-literal YOLOLayer.forward outputs.
//
//   k_decode_dense        eval branch (yololayer.py:146-166): planar raw [B,3,5+C,F,F] -> AoS rows [.., 5+C], written
//                         directly at the layer's row offset of the concatenated [B, M, 5+C] tensor (yolov4.py:324).
//                         Coalesced 256-byte plane reads, shared-memory transpose, linear coalesced row writes.
//   k_decode_train        train branch (yololayer.py:122-145): planar `output` (sigmoid on xy/obj/cls) + planar `pred`.
//   k_decode_train_bwd    d(output)/d(raw) for the autograd wrapper.
#include "yl_common.cuh"
#include "../../include/yolo_head.h"

namespace yl {

constexpr int DD_TP = 64;          // boxes per tile
constexpr int DD_THREADS = 256;    // 4 channel rows x 64 boxes per pass

__global__ void __launch_bounds__(DD_THREADS)
k_decode_dense(const float *__restrict__ raw, int Fw, int F2, int C, float stride,
               float aw0, float ah0, float aw1, float ah1, float aw2, float ah2,
               float *__restrict__ out, long rows_per_image, long row_offset)
{
    extern __shared__ float tile[];            // [DD_TP][nchp], nchp odd -> conflict-free transposed stores
    const int nch = 5 + C;
    const int nchp = nch | 1;
    const int ba = blockIdx.y;
    const int b = ba / 3, a = ba - 3 * b;
    const int p0 = blockIdx.x * DD_TP;
    const int np = min(DD_TP, F2 - p0);
    const int pl = threadIdx.x & (DD_TP - 1);
    const int kr = threadIdx.x / DD_TP;
    const float aw = (a == 0) ? aw0 : ((a == 1) ? aw1 : aw2);
    const float ah = (a == 0) ? ah0 : ((a == 1) ? ah1 : ah2);
    if (pl < np) {
        const int p = p0 + pl;
        const int gy = p / Fw, gx = p - gy * Fw;
        const float *src = raw + ((size_t)ba * nch) * F2 + p;
        for (int k = kr; k < nch; k += DD_THREADS / DD_TP) {
            const float t = ldg_stream1(src + (size_t)k * F2);
            float v;
            if (k == 0) v = __fmul_rn(__fadd_rn(spec_sigmoidf(t), (float)gx), stride);
            else if (k == 1) v = __fmul_rn(__fadd_rn(spec_sigmoidf(t), (float)gy), stride);
            else if (k == 2) v = __fmul_rn(__fmul_rn(spec_expf(t), aw), stride);
            else if (k == 3) v = __fmul_rn(__fmul_rn(spec_expf(t), ah), stride);
            else v = spec_sigmoidf(t);
            tile[pl * nchp + k] = v;
        }
    }
    __syncthreads();
    float *dst = out + ((size_t)b * rows_per_image + row_offset + (size_t)a * F2 + p0) * nch;
    const int n = np * nch;
    for (int e = threadIdx.x; e < n; e += DD_THREADS) {
        const int box = e / nch, k = e - box * nch;
        dst[e] = tile[box * nchp + k];
    }
}

constexpr int DT_THREADS = 256;

__global__ void __launch_bounds__(DT_THREADS)
k_decode_train(const float *__restrict__ raw, int Fw, int F2, int C, long total,
               float aw0, float ah0, float aw1, float ah1, float aw2, float ah2,
               float *__restrict__ output_planar, float *__restrict__ pred_planar)
{
    const int nch = 5 + C;
    for (long idx = (long)blockIdx.x * DT_THREADS + threadIdx.x; idx < total; idx += (long)gridDim.x * DT_THREADS) {
        const long plane = idx / F2;
        const int p = (int)(idx - plane * F2);
        const int ba = (int)(plane / nch);
        const int k = (int)(plane - (long)ba * nch);
        const float t = raw[idx];
        float o = t;
        if (k != 2 && k != 3) o = spec_sigmoidf(t);                          // yololayer.py:105
        output_planar[idx] = o;
        if (k < 4) {
            const int a = ba % 3;
            float v;
            if (k == 0) v = __fadd_rn(o, (float)(p % Fw));                   // :126
            else if (k == 1) v = __fadd_rn(o, (float)(p / Fw));              // :129
            else if (k == 2) v = __fmul_rn(spec_expf(t), (a == 0) ? aw0 : ((a == 1) ? aw1 : aw2));   // :132
            else v = __fmul_rn(spec_expf(t), (a == 0) ? ah0 : ((a == 1) ? ah1 : ah2));               // :134
            pred_planar[((size_t)ba * 4 + k) * F2 + p] = v;
        }
    }
}

__global__ void __launch_bounds__(DT_THREADS)
k_decode_train_bwd(const float *__restrict__ output_planar, const float *__restrict__ grad_out, int F2, int C, long total,
                   float *__restrict__ grad_raw)
{
    const int nch = 5 + C;
    for (long idx = (long)blockIdx.x * DT_THREADS + threadIdx.x; idx < total; idx += (long)gridDim.x * DT_THREADS) {
        const int k = (int)((idx / F2) % nch);
        const float g = grad_out[idx];
        const float o = output_planar[idx];
        grad_raw[idx] = (k == 2 || k == 3) ? g : (g * (1.0f - o)) * o;     // ATen sigmoid_backward order
    }
}

}  // namespace yl

using namespace yl;

extern "C" int yl_decode_dense(const float *raw, int B, int F, int C, const float *ag, float stride,
                               float *out, long rows_per_image, long row_offset, yl_stream_t stream)
{
    if (!raw || !ag || !out || B <= 0 || F <= 0 || C <= 0 || rows_per_image < 3L * F * F + row_offset || row_offset < 0)
        return YL_ERR_ARG;
    const int F2 = F * F;
    const int nchp = (5 + C) | 1;
    const size_t smem = sizeof(float) * DD_TP * nchp;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_decode_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return YL_ERR_CUDA_BASE + (int)e;
    }
    dim3 grid((F2 + DD_TP - 1) / DD_TP, B * 3);
    k_decode_dense<<<grid, DD_THREADS, smem, (cudaStream_t)stream>>>(raw, F, F2, C, stride, ag[0], ag[1], ag[2], ag[3],
                                                                     ag[4], ag[5], out, rows_per_image, row_offset);
    YL_LAUNCH_CHECK();
    return YL_OK;
}

extern "C" int yl_decode_train(const float *raw, int B, int F, int C, const float *ag,
                               float *output_planar, float *pred_planar, yl_stream_t stream)
{
    if (!raw || !ag || !output_planar || !pred_planar || B <= 0 || F <= 0 || C <= 0) return YL_ERR_ARG;
    const int F2 = F * F;
    const long total = (long)B * 3 * (5 + C) * F2;
    const long blocks = (total + DT_THREADS - 1) / DT_THREADS;
    const int grid = (int)(blocks < 148L * 32 ? blocks : 148L * 32);
    k_decode_train<<<grid, DT_THREADS, 0, (cudaStream_t)stream>>>(raw, F, F2, C, total, ag[0], ag[1], ag[2], ag[3], ag[4],
                                                                  ag[5], output_planar, pred_planar);
    YL_LAUNCH_CHECK();
    return YL_OK;
}

extern "C" int yl_decode_train_backward(const float *output_planar, const float *grad_out_planar, int B, int F, int C,
                                        float *grad_raw, yl_stream_t stream)
{
    if (!output_planar || !grad_out_planar || !grad_raw || B <= 0 || F <= 0 || C <= 0) return YL_ERR_ARG;
    const int F2 = F * F;
    const long total = (long)B * 3 * (5 + C) * F2;
    const long blocks = (total + DT_THREADS - 1) / DT_THREADS;
    const int grid = (int)(blocks < 148L * 32 ? blocks : 148L * 32);
    k_decode_train_bwd<<<grid, DT_THREADS, 0, (cudaStream_t)stream>>>(output_planar, grad_out_planar, F2, C, total, grad_raw);
    YL_LAUNCH_CHECK();
    return YL_OK;
}
