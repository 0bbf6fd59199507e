// canonicalvoting_b200/csrc/sparse_conv.cu -- sparse convolution as an output-stationary implicit GEMM.
//
// Replaces MinkowskiEngine's convolution kernels (external package; reference call sites
// utils/minkunet.py:53-114 and MinkowskiEngine.modules.resnet_block.BasicBlock).  ME runs, per kernel
// offset, gather -> GEMM -> scatter-add through global memory [ME-recall]; here ONE kernel walks all K^3
// offsets of an output tile, gathers the neighbour rows named by the neighbour table straight into shared
// memory and keeps the accumulators on chip, so every output row is written exactly once and there is no
// scatter and no atomic in the forward / input-gradient passes:
//     out[o, :] = sum_k  in[nbr[o, k], :] @ W[k]        (rows with nbr = -1 contribute nothing)
// The same kernel serves stride-1 3^3/5^3 convolutions, the stride-2 2^3 convolution (children table), the
// transposed 2^3 convolution (parent table) and all input gradients (same tables, transposed weights --
// see canonicalvoting_b200/sparse/functional.py).  The weight gradient is a split-K reduction over table rows.
//
// This file holds the fp32 CUDA-core path (exact fp32 accumulation: the parity mode and the backward
// pass); the tcgen05 tensor-core forward lives in sparse_conv_tc.cu.
#include "common.cuh"

namespace cvb200 {

constexpr int kBM = 64, kBN = 64, kBK = 16, kConvThreads = 256;

// One CTA: kBM output rows x kBN output channels; thread (ty, tx) owns a 4x4 micro-tile.
__global__ void __launch_bounds__(kConvThreads)
sc_conv_table_kernel(const float *__restrict__ in, int cin, const float *__restrict__ w, int cout,
                     const int *__restrict__ nbr, int n_out, int k3, const float *__restrict__ bias,
                     float *__restrict__ out) {
    __shared__ float As[kBK][kBM + 4];
    __shared__ __align__(16) float Bs[kBK][kBN];
    __shared__ int s_idx[kBM];
    const int row0 = blockIdx.x * kBM, n0 = blockIdx.y * kBN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    for (int k = 0; k < k3; k++) {
        int mine = -1;
        if (threadIdx.x < kBM) {
            const int r = row0 + threadIdx.x;
            mine = r < n_out ? __ldg(nbr + (size_t)r * k3 + k) : -1;
            s_idx[threadIdx.x] = mine;
        }
        if (!__syncthreads_or(mine >= 0)) continue;   // no row of this tile has a neighbour at offset k
        const float *wk = w + (size_t)k * cin * cout;
        for (int c0 = 0; c0 < cin; c0 += kBK) {
            // A tile: 64 rows x 16 channels, consecutive threads read consecutive channels of a row
#pragma unroll
            for (int e = threadIdx.x; e < kBM * kBK; e += kConvThreads) {
                const int r = e / kBK, kk = e % kBK;
                const int idx = s_idx[r];
                As[kk][r] = (idx >= 0 && c0 + kk < cin) ? __ldg(in + (size_t)idx * cin + c0 + kk) : 0.f;
            }
#pragma unroll
            for (int e = threadIdx.x; e < kBK * kBN; e += kConvThreads) {
                const int kk = e / kBN, nn = e % kBN;
                Bs[kk][nn] = (c0 + kk < cin && n0 + nn < cout) ? __ldg(wk + (size_t)(c0 + kk) * cout + n0 + nn) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < kBK; kk++) {
                float a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; i++) a[i] = As[kk][ty * 4 + i];
                const float4 b4 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
                b[0] = b4.x; b[1] = b4.y; b[2] = b4.z; b[3] = b4.w;
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int r = row0 + ty * 4 + i;
        if (r >= n_out) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = n0 + tx * 4 + j;
            if (c < cout) out[(size_t)r * cout + c] = acc[i][j] + (bias ? __ldg(bias + c) : 0.f);
        }
    }
}

// dW[k] (ca x cb) += sum over table rows r of  A[ia(r,k), :]^T (x) B[ib(r,k), :]
//   table_on_b == 0:  ia = table[r][k], ib = r      (convolutions driven by a neighbour / children table)
//   table_on_b == 1:  ia = r,           ib = table[r][k]
// grid = (row chunks, ca tiles * cb tiles, K^3); split-K partial tiles are combined with float atomics.
constexpr int kWgRows = 1024;   // table rows per CTA

__global__ void __launch_bounds__(kConvThreads)
sc_wgrad_kernel(const float *__restrict__ a, int ca, const float *__restrict__ b, int cb, const int *__restrict__ table,
                int n_rows, int k3, int table_on_b, float *__restrict__ dw) {
    __shared__ float As[kBK][kBM + 4];
    __shared__ __align__(16) float Bs[kBK][kBN];
    __shared__ int s_idx[kBK];
    const int k = blockIdx.z;
    const int tiles_b = (cb + kBN - 1) / kBN;
    const int a0 = (blockIdx.y / tiles_b) * kBM, b0 = (blockIdx.y % tiles_b) * kBN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int r_begin = blockIdx.x * kWgRows, r_end = min(n_rows, r_begin + kWgRows);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    bool any_work = false;
    for (int r0 = r_begin; r0 < r_end; r0 += kBK) {
        int mine = -1;
        if (threadIdx.x < kBK) {
            const int r = r0 + threadIdx.x;
            mine = r < r_end ? __ldg(table + (size_t)r * k3 + k) : -1;
            s_idx[threadIdx.x] = mine;
        }
        if (!__syncthreads_or(mine >= 0)) continue;
        any_work = true;
#pragma unroll
        for (int e = threadIdx.x; e < kBK * kBM; e += kConvThreads) {
            const int rr = e / kBM, cc = e % kBM;
            const int t = s_idx[rr];
            const int row = table_on_b ? r0 + rr : t;
            As[rr][cc] = (t >= 0 && a0 + cc < ca) ? __ldg(a + (size_t)row * ca + a0 + cc) : 0.f;
        }
#pragma unroll
        for (int e = threadIdx.x; e < kBK * kBN; e += kConvThreads) {
            const int rr = e / kBN, cc = e % kBN;
            const int t = s_idx[rr];
            const int row = table_on_b ? t : r0 + rr;
            Bs[rr][cc] = (t >= 0 && b0 + cc < cb) ? __ldg(b + (size_t)row * cb + b0 + cc) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < kBK; rr++) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; i++) av[i] = As[rr][ty * 4 + i];
            const float4 b4 = *reinterpret_cast<const float4 *>(&Bs[rr][tx * 4]);
            bv[0] = b4.x; bv[1] = b4.y; bv[2] = b4.z; bv[3] = b4.w;
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    if (!any_work) return;
    float *dwk = dw + (size_t)k * ca * cb;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int ra = a0 + ty * 4 + i;
        if (ra >= ca) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = b0 + tx * 4 + j;
            if (c < cb && acc[i][j] != 0.f) atomicAdd(dwk + (size_t)ra * cb + c, acc[i][j]);
        }
    }
}

}  // namespace cvb200

using namespace cvb200;

extern "C" int cvb200_sc_conv_forward(const float *d_in, int32_t cin, const float *d_w, int32_t cout, const int32_t *d_nbr,
                                      int64_t n_out, int32_t k3, const float *d_bias, float *d_out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVB_REQUIRE(cin > 0 && cout > 0 && k3 > 0 && n_out >= 0 && n_out < (1LL << 31), CVB200_EINVAL, "sc_conv_forward: bad sizes");
    if (n_out == 0) return 0;
    CVB_REQUIRE(d_in && d_w && d_nbr && d_out, CVB200_EINVAL, "sc_conv_forward: NULL argument");
    dim3 grid((unsigned)ceil_div(n_out, kBM), (unsigned)ceil_div(cout, kBN));
    sc_conv_table_kernel<<<grid, kConvThreads, 0, stream>>>(d_in, cin, d_w, cout, d_nbr, (int)n_out, k3, d_bias, d_out);
    CVB_LAUNCH_CHECK("sc_conv_table_kernel");
    return 0;
}

extern "C" int cvb200_sc_conv_wgrad(const float *d_a, int32_t ca, const float *d_b, int32_t cb, const int32_t *d_table,
                                    int64_t n_rows, int32_t k3, int32_t table_on_b, float *d_dw, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVB_REQUIRE(ca > 0 && cb > 0 && k3 > 0 && k3 <= 65535 && n_rows >= 0 && n_rows < (1LL << 31), CVB200_EINVAL,
                "sc_conv_wgrad: bad sizes");
    CVB_REQUIRE(d_dw, CVB200_EINVAL, "sc_conv_wgrad: NULL dw");
    CVB_CUDA(cudaMemsetAsync(d_dw, 0, sizeof(float) * (size_t)k3 * ca * cb, stream));
    if (n_rows == 0) return 0;
    CVB_REQUIRE(d_a && d_b && d_table, CVB200_EINVAL, "sc_conv_wgrad: NULL argument");
    const int tiles = (int)(ceil_div(ca, kBM) * ceil_div(cb, kBN));
    dim3 grid((unsigned)ceil_div(n_rows, kWgRows), (unsigned)tiles, (unsigned)k3);
    sc_wgrad_kernel<<<grid, kConvThreads, 0, stream>>>(d_a, ca, d_b, cb, d_table, (int)n_rows, k3, table_on_b, d_dw);
    CVB_LAUNCH_CHECK("sc_wgrad_kernel");
    return 0;
}
