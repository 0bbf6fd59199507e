// canonicalvoting_b200/csrc/sparse_conv_tc.cu -- tensor-core (tcgen05 / TMEM) forward of the sparse convolution.
//
//     out[o, :] = sum_k  in[nbr[o, k], :] @ W[k]
// as an output-stationary implicit GEMM on the 5th-generation tensor cores of sm_100a:
//   * one CTA owns a tile of 128 output rows; its accumulator (128 x Cout fp32) lives in TENSOR MEMORY for the
//     whole walk over the K^3 kernel offsets and is read back exactly once (tcgen05.ld) by the epilogue;
//   * per (offset, 32-channel block) the CTA gathers the neighbour rows named by the neighbour table straight
//     from the feature matrix into a 128B-swizzled K-major shared-memory tile with 16-byte cp.async copies
//     (rows without a neighbour are zero-filled by a zero-length copy, so only existing pairs cost bandwidth)
//     and stages the matching 32 x Cout weight block next to it; a 3-4 stage ring overlaps the gathers with
//     the MMAs;
//   * ONE elected thread issues tcgen05.mma.kind::tf32 (M=128, N=Cout, K=8; fp32 features and weights are
//     consumed as TF32, accumulation is fp32) and tcgen05.commit releases each stage through an mbarrier;
//   * offsets for which no row of the tile has a neighbour are skipped.
// Weights arrive pre-transposed as Wt[k][cout][cin] (K-major B operand).  Requirements: cin % 32 == 0,
// cout % 16 == 0, 16 <= cout <= 256, K^3 <= 32.  Everything else (cin = 3 stem, gradients) uses the fp32
// CUDA-core path in sparse_conv.cu.
#include "common.cuh"

namespace cvb200 {

constexpr int kTcM = 128;          // output rows per CTA = UMMA M
constexpr int kTcKB = 32;          // channels per k-block = 128 bytes of tf32 = one swizzle-128B row
constexpr int kTcThreads = 256;
constexpr int kTcMaxK3 = 32;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// 16-byte async copy global -> shared; src_bytes = 0 zero-fills the destination without reading
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// shared-memory matrix descriptor, K-major, SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffff) >> 4);        // start address            bits [0,14)
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major) [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset: next 8-row group            [32,46)
    d |= (uint64_t)1 << 46;                         // descriptor version (sm_100)                      [46,48)
    d |= (uint64_t)2 << 61;                         // layout type SWIZZLE_128B                         [61,64)
    return d;
}

// D[tmem] (+)= A[smem] * B[smem]^T, kind::tf32, issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

struct TcSmemHeader {
    unsigned long long empty_bar[4];
    unsigned int tmem_base;
    int n_act;
    int act[kTcMaxK3];
    int any[kTcMaxK3];
};

// dynamic smem: [header 1 KiB][nbr tile: K3 x 128 ints, padded to 1 KiB][stages x (A tile 16 KiB | B tile N x 128 B)]
template <int kStages>
__global__ void __launch_bounds__(kTcThreads)
sc_conv_tc_kernel(const float *__restrict__ in, int ldi, int cin, const float *__restrict__ wt, int cout_total,
                  const int *__restrict__ nbr, int n_out, int k3, const float *__restrict__ bias, const float *__restrict__ residual,
                  int ldr, int relu, float *__restrict__ out, int ldo, int tmem_cols, int cout) {
    // grid = (row tiles, channel splits, offset splits): this CTA produces out[row tile, n0 .. n0+cout) from the kernel
    // offsets k with k % gridDim.z == blockIdx.z (partial sums of different offset splits are combined with float atomics)
    const int n0 = blockIdx.y * cout;
    extern __shared__ __align__(1024) unsigned char smem[];
    TcSmemHeader &H = *reinterpret_cast<TcSmemHeader *>(smem);
    int *s_nbr = reinterpret_cast<int *>(smem + 1024);
    const int nbr_bytes = ((k3 * kTcM * 4 + 1023) / 1024) * 1024;
    unsigned char *stage0 = smem + 1024 + nbr_bytes;
    const int a_bytes = kTcM * 128, b_bytes = cout * 128, stage_bytes = a_bytes + b_bytes;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * kTcM;

    // ---- prologue: neighbour tile, list of active offsets, barriers, tensor memory
    if (tid < kTcMaxK3) H.any[tid] = 0;
    __syncthreads();
    for (int e = tid; e < k3 * kTcM; e += kTcThreads) {
        const int r = e / k3, k = e - r * k3;   // consecutive threads read consecutive table entries
        const int v = row0 + r < n_out ? __ldg(nbr + (size_t)(row0 + r) * k3 + k) : -1;
        s_nbr[k * kTcM + r] = v;
        if (v >= 0 && k % (int)gridDim.z == (int)blockIdx.z) H.any[k] = 1;   // benign race
    }
    if (tid == 0) {
        for (int s = 0; s < kStages; s++) mbar_init(smem_u32(&H.empty_bar[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&H.tmem_base)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        for (int k = 0; k < k3; k++)
            if (H.any[k]) H.act[n++] = k;
        H.n_act = n;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = H.tmem_base;
    const int cblocks = cin / kTcKB;
    const int total = H.n_act * cblocks;
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N = cout, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(cout >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);

    // ---- per-thread copy plan: thread (rbase, c) copies the 16-byte chunk c of rows rbase + 32 j of the A tile (4 rows)
    // and of the B tile (cout / 32 rows) of every k-block; only the source row pointers change, once per kernel offset.
    const int c = tid & 7, rbase = tid >> 3;
    uint32_t t_off[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int r = rbase + 32 * j;
        t_off[j] = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
    }
    const float *a_src[4];
    uint32_t a_ok[4];
    const float *b_src = wt;
    int l_act = 0, l_cb = 0;   // load cursor: (active offset, channel block) of the next k-block to fetch
    auto setup_offset = [&](int act_i) {
        const int k = H.act[act_i];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int idx = s_nbr[k * kTcM + rbase + 32 * j];
            a_src[j] = in + (size_t)(idx >= 0 ? idx : 0) * ldi + c * 4;
            a_ok[j] = idx >= 0 ? 16u : 0u;
        }
        b_src = wt + ((size_t)k * cout_total + n0 + rbase) * cin + c * 4;
    };
    auto issue_loads = [&](int stage) {
        const uint32_t a_s = smem_u32(stage0 + (size_t)stage * stage_bytes), b_s = a_s + a_bytes;
        const int coff = l_cb * kTcKB;
#pragma unroll
        for (int j = 0; j < 4; j++) cp_async16(a_s + t_off[j], a_src[j] + coff, a_ok[j]);
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (rbase + 32 * j < cout) cp_async16(b_s + t_off[j], b_src + (size_t)(32 * j) * cin + coff, 16u);
        if (++l_cb == cblocks) {
            l_cb = 0;
            if (++l_act < H.n_act) setup_offset(l_act);
        }
    };

    // ---- main loop: kStages-deep ring of (gathered A tile, weight tile); one thread issues the MMAs
    if (total > 0) setup_offset(0);
    for (int b = 0; b < kStages - 1; b++) {
        if (b < total) issue_loads(b);
        cp_async_commit();
    }
    for (int it = 0; it < total; it++) {
        cp_async_wait<kStages - 2>();                                  // this thread's copies of k-block `it` have landed
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes -> visible to the tensor core
        // the stage refilled below was read by the MMAs of k-block it-1: one thread waits for their completion
        if (tid == 0 && it >= 1) mbar_wait(smem_u32(&H.empty_bar[(it - 1) % kStages]), (uint32_t)(((it - 1) / kStages) & 1));
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_s = smem_u32(stage0 + (size_t)(it % kStages) * stage_bytes), b_s = a_s + a_bytes;
            const uint64_t a_desc = umma_desc_k_sw128(a_s), b_desc = umma_desc_k_sw128(b_s);
#pragma unroll
            for (int kk = 0; kk < kTcKB / 8; kk++)   // UMMA K = 8 tf32 = 32 bytes: advance the start address inside the swizzle row
                umma_tf32(tmem, a_desc + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2), idesc, (it > 0 || kk > 0) ? 1u : 0u);
            umma_commit(smem_u32(&H.empty_bar[it % kStages]));   // arrives when these MMAs (and all before) have completed
        }
        if (it + kStages - 1 < total) issue_loads((it + kStages - 1) % kStages);
        cp_async_commit();
    }
    cp_async_wait<0>();

    // ---- epilogue: TMEM -> registers -> (+bias) -> global, each thread one output row, 16 columns at a time
    if (total > 0) {
        mbar_wait(smem_u32(&H.empty_bar[(total - 1) % kStages]), (uint32_t)(((total - 1) / kStages) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    {
        const int q = warp & 3, half = warp >> 2;          // TMEM lane quarter of this warp, column half
        const int r = row0 + q * 32 + lane;
        const int ncol16 = cout / 16;
        for (int cb = half; cb < ncol16; cb += 2) {
            uint32_t v[16];
            if (total > 0) {
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * 16);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int j = 0; j < 16; j++) v[j] = 0u;
            }
            if (r < n_out && (total > 0 || gridDim.z == 1 || (bias != nullptr && blockIdx.z == 0))) {
                float *dst = out + (size_t)r * ldo + n0 + cb * 16;
                const bool add_bias = bias != nullptr && blockIdx.z == 0;
                const float *res = residual ? residual + (size_t)r * ldr + n0 + cb * 16 : nullptr;   // only with gridDim.z == 1
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                           __uint_as_float(v[4 * j + 3]));
                    if (add_bias) {
                        const float *bp = bias + n0 + cb * 16 + 4 * j;
                        o.x += __ldg(bp); o.y += __ldg(bp + 1); o.z += __ldg(bp + 2); o.w += __ldg(bp + 3);
                    }
                    if (res) {
                        const float4 rv = __ldg(reinterpret_cast<const float4 *>(res) + j);
                        o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
                    }
                    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    if (gridDim.z == 1)
                        reinterpret_cast<float4 *>(dst)[j] = o;
                    else   // split over kernel offsets: out was zero-filled by the host wrapper
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * j), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
}

// finishing pass of a convolution whose kernel offsets were split over several CTAs (partial sums combined with
// atomics): out = [relu](out + bias + residual)
__global__ void sc_finish_kernel(float *__restrict__ out, int ldo, int n, int cout, const float *__restrict__ bias,
                                 const float *__restrict__ residual, int ldr, int relu) {
    const int c4 = cout / 4;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * c4) return;
    const int r = (int)(t / c4), c = (int)(t % c4) * 4;
    float4 o = *reinterpret_cast<float4 *>(out + (size_t)r * ldo + c);
    if (bias) { o.x += __ldg(bias + c); o.y += __ldg(bias + c + 1); o.z += __ldg(bias + c + 2); o.w += __ldg(bias + c + 3); }
    if (residual) {
        const float4 rv = __ldg(reinterpret_cast<const float4 *>(residual + (size_t)r * ldr + c));
        o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
    }
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    *reinterpret_cast<float4 *>(out + (size_t)r * ldo + c) = o;
}

// out[o, 0:cout) (row stride ldo) = [relu]( sum_k in[nbr[o,k], 0:cin) (row stride ldi) @ Wt[k]^T + bias + residual )
int launch_finish(float *d_out, int ldo, int64_t n_out, int cout, const float *d_bias, const float *d_res, int ldr, int relu,
                  cudaStream_t stream) {
    const long long total = (long long)n_out * (cout / 4);
    sc_finish_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(d_out, ldo, (int)n_out, cout, d_bias, d_res, ldr, relu);
    CVB_LAUNCH_CHECK("sc_finish_kernel");
    return 0;
}

int launch_conv_tma(int gather_a, const float *d_in, int64_t n_in, int ldi, int cin, const float *d_wt, int cout, const int32_t *d_nbr,
                    int64_t n_out, int k3, const float *d_bias, const float *d_res, int ldr, int relu, float *d_out, int ldo,
                    cudaStream_t stream);
int launch_conv_persist(const float *d_in, int64_t n_in, int ldi, int cin, const float *d_wt, int cout, const int32_t *d_nbr,
                        int64_t n_out, int k3, const float *d_bias, const float *d_res, int ldr, int relu, float *d_out, int ldo,
                        cudaStream_t stream, int g4 = 0);
// 3: persistent warp-specialised kernel, two TMEM accumulators, split tiles reduced in-kernel (sparse_conv_persist.cu);
// 2: warp-specialised kernel, one tile per CTA, A by cp.async producers + B by TMA (sparse_conv_tma.cu); 1: same kernel, A by
// TMA gather4; 0: cp.async kernel with a CTA-wide barrier per k-block (this file)
int g_conv_impl = 3;

int launch_conv_tc(const float *d_in, int64_t n_in, int ldi, int cin, const float *d_wt, int cout, const int32_t *d_nbr, int64_t n_out,
                   int k3, const float *d_bias, const float *d_res, int ldr, int relu, float *d_out, int ldo, cudaStream_t stream) {
    if (g_conv_impl == 3)
        return launch_conv_persist(d_in, n_in, ldi, cin, d_wt, cout, d_nbr, n_out, k3, d_bias, d_res, ldr, relu, d_out, ldo, stream);
    if (g_conv_impl != 0)
        return launch_conv_tma(g_conv_impl == 1, d_in, n_in, ldi, cin, d_wt, cout, d_nbr, n_out, k3, d_bias, d_res, ldr, relu, d_out, ldo, stream);
    CVB_REQUIRE(cin > 0 && cin % kTcKB == 0 && cout >= 16 && cout <= 256 && cout % 16 == 0 && k3 > 0 && k3 <= kTcMaxK3,
                CVB200_EINVAL, "sc_conv_forward_tc: needs cin %% 32 == 0, cout %% 16 == 0, 16 <= cout <= 256, K^3 <= %d (got %d, %d, %d)",
                kTcMaxK3, cin, cout, k3);
    CVB_REQUIRE(n_out >= 0 && n_out < (1LL << 31), CVB200_EINVAL, "sc_conv_forward_tc: bad n_out");
    if (n_out == 0) return 0;
    CVB_REQUIRE(d_in && d_wt && d_nbr && d_out, CVB200_EINVAL, "sc_conv_forward_tc: NULL argument");
    CVB_REQUIRE(((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_wt) | reinterpret_cast<uintptr_t>(d_out) |
                  reinterpret_cast<uintptr_t>(d_res)) & 15) == 0 && ldi % 4 == 0 && ldo % 4 == 0 && ldr % 4 == 0,
                CVB200_EINVAL, "sc_conv_forward_tc: 16-byte aligned pointers and row strides required");
    // Small levels of the U-Net have only a handful of 128-row tiles: split the output channels (64 per CTA) and the
    // kernel offsets over more CTAs until the grid covers the 148 SMs.
    const int m_tiles = (int)ceil_div(n_out, kTcM);
    int n_splits = 1, k_splits = 1;
    if (m_tiles < kNumSMs / 2 && cout >= 128 && cout % 64 == 0) n_splits = cout / 64;
    const int kmul = (k3 % 3 == 0) ? 3 : 2, cblocks = cin / kTcKB;
    while (m_tiles * n_splits * k_splits < kNumSMs && k_splits * kmul <= k3 && (k3 / (k_splits * kmul)) * cblocks >= 8)
        k_splits *= kmul;   // keep at least 8 k-blocks per CTA so that the prologue stays amortised
    const int nc = cout / n_splits;
    int tmem_cols = 32;
    while (tmem_cols < nc) tmem_cols <<= 1;
    const int nbr_bytes = ((k3 * kTcM * 4 + 1023) / 1024) * 1024;
    const int stage_bytes = kTcM * 128 + nc * 128;
    const dim3 grid((unsigned)m_tiles, (unsigned)n_splits, (unsigned)k_splits);
    const bool split = k_splits > 1;
    if (split) CVB_CUDA(cudaMemset2DAsync(d_out, sizeof(float) * (size_t)ldo, 0, sizeof(float) * (size_t)cout, (size_t)n_out, stream));
    const float *k_bias = split ? nullptr : d_bias, *k_res = split ? nullptr : d_res;
    const int k_relu = split ? 0 : relu;
    if (nc <= 64) {   // 4 stages of 24 KiB / 3 stages of <= 48 KiB: two CTAs per SM up to 128 channels per CTA
        const size_t smem = 1024 + nbr_bytes + 4 * (size_t)stage_bytes;
        static bool set4 = false;
        if (!set4) {
            CVB_CUDA(cudaFuncSetAttribute(sc_conv_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            set4 = true;
        }
        sc_conv_tc_kernel<4><<<grid, kTcThreads, smem, stream>>>(d_in, ldi, cin, d_wt, cout, d_nbr, (int)n_out, k3, k_bias, k_res, ldr,
                                                                 k_relu, d_out, ldo, tmem_cols, nc);
    } else {
        const size_t smem = 1024 + nbr_bytes + 3 * (size_t)stage_bytes;
        static bool set3 = false;
        if (!set3) {
            CVB_CUDA(cudaFuncSetAttribute(sc_conv_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            set3 = true;
        }
        sc_conv_tc_kernel<3><<<grid, kTcThreads, smem, stream>>>(d_in, ldi, cin, d_wt, cout, d_nbr, (int)n_out, k3, k_bias, k_res, ldr,
                                                                 k_relu, d_out, ldo, tmem_cols, nc);
    }
    CVB_LAUNCH_CHECK("sc_conv_tc_kernel");
    if (split && (d_bias || d_res || relu)) return launch_finish(d_out, ldo, n_out, cout, d_bias, d_res, ldr, relu, stream);
    return 0;
}
}  // namespace cvb200

using namespace cvb200;

extern "C" int cvb200_sc_conv_forward_tc(const float *d_in, int64_t n_in, int32_t cin, const float *d_wt, int32_t cout,
                                         const int32_t *d_nbr, int64_t n_out, int32_t k3, const float *d_bias, float *d_out,
                                         void *stream_) {
    return launch_conv_tc(d_in, n_in, cin, cin, d_wt, cout, d_nbr, n_out, k3, d_bias, nullptr, 0, 0, d_out, cout, (cudaStream_t)stream_);
}

extern "C" int cvb200_sc_set_conv_impl(int32_t impl) {
    CVB_REQUIRE(impl >= 0 && impl <= 3, CVB200_EINVAL, "sc_set_conv_impl: 0 (cp.async), 1 (TMA gather4), 2 (cp.async A + TMA B) or 3 (persistent)");
    g_conv_impl = impl;
    return 0;
}
