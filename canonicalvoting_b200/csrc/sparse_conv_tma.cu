// canonicalvoting_b200/csrc/sparse_conv_tma.cu -- warp-specialised tcgen05 sparse convolution fed by TMA.
//
// Same contraction and the same shared-memory / tensor-memory layout as sparse_conv_tc.cu, but the operand tiles are
// moved by the Tensor Memory Accelerator instead of 16-byte cp.async copies (which made the cp.async version
// LSU-issue-bound: 1792 copies per k-block):
//   * A (gathered neighbour rows): `cp.async.bulk.tensor.2d ... tile::gather4` -- ONE instruction fetches four
//     arbitrary rows x 128 bytes of the feature matrix, named by four row coordinates, into the 128B-swizzled
//     tile; a missing neighbour (-1) is an out-of-bounds coordinate, which TMA zero-fills.  32 lanes x gather4
//     = the 128-row tile.
//   * B (weights): one tiled 2-D TMA load of the [Cout x 32] block of Wt[k].
//   * roles: warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warps 2-5 = epilogue (one TMEM lane
//     quarter each).  full/empty mbarriers per stage; no __syncthreads in the main loop.
#include <cuda.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace cvb200 {

constexpr int kTmM = 128, kTmKB = 32, kTmThreads = 192, kTmMaxK3 = 32;


struct TmHeader {
    unsigned long long full_bar[4], empty_bar[4], accum_bar;
    unsigned int tmem_base;
    int n_act;
    int act[kTmMaxK3];
    int any[kTmMaxK3];
};

// dynamic smem: [header 1 KiB][nbr tile K3 x 128 ints, padded to 1 KiB][stages x (A 16 KiB | B cout x 128 B)]
// kGatherA: true  -> warp 0 fetches the A tile with TMA tile::gather4 (32 ops per k-block; measured slower: the TMA
//                     unit serialises the small gathers);
//           false -> the four epilogue warps are the A producers during the main loop: 16-byte cp.async copies whose
//                     completion arrives on the stage's full barrier by itself (cp.async.mbarrier.arrive.noinc; no
//                     CTA-wide barrier, no waiting producer), while the weight block arrives by one TMA load.
template <int kStages, bool kGatherA>
__global__ void __launch_bounds__(kTmThreads)
sc_conv_tma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const float *__restrict__ in,
                   int ldi, int cin, int cout_total,
                   const int *__restrict__ nbr, int n_out, int k3, const float *__restrict__ bias,
                   const float *__restrict__ residual, int ldr, int relu, float *__restrict__ out, int ldo, int tmem_cols, int cout) {
    extern __shared__ __align__(1024) unsigned char smem[];
    TmHeader &H = *reinterpret_cast<TmHeader *>(smem);
    int *s_nbr = reinterpret_cast<int *>(smem + 1024);
    const int nbr_bytes = ((k3 * kTmM * 4 + 1023) / 1024) * 1024;
    unsigned char *stage0 = smem + 1024 + nbr_bytes;
    const int a_bytes = kTmM * 128, b_bytes = cout * 128, stage_bytes = a_bytes + b_bytes;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * kTmM, n0 = blockIdx.y * cout;

    if (tid < kTmMaxK3) H.any[tid] = 0;
    __syncthreads();
    for (int e = tid; e < k3 * kTmM; e += kTmThreads) {
        const int r = e / k3, k = e - r * k3;
        const int v = row0 + r < n_out ? __ldg(nbr + (size_t)(row0 + r) * k3 + k) : -1;
        s_nbr[k * kTmM + r] = v;
        if (v >= 0 && k % (int)gridDim.z == (int)blockIdx.z) H.any[k] = 1;
    }
    if (tid == 0) {
        for (int s = 0; s < kStages; s++) {
            tm_mbar_init(tm_smem_u32(&H.full_bar[s]), kGatherA ? 1 : 1 + 128);
            tm_mbar_init(tm_smem_u32(&H.empty_bar[s]), 1);
        }
        tm_mbar_init(tm_smem_u32(&H.accum_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tm_smem_u32(&H.tmem_base)), "r"(tmem_cols) : "memory");
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
    const int cblocks = cin / kTmKB;
    const int total = H.n_act * cblocks;

    if (warp == 0) {
        // ===== TMA producer: lane j gathers rows 4j .. 4j+3 of the A tile; lane 0 also loads the weight block
        for (int it = 0; it < total; it++) {
            const int s = it % kStages;
            tm_mbar_wait(tm_smem_u32(&H.empty_bar[s]), (uint32_t)(((it / kStages) & 1) ^ 1));
            const int k = H.act[it / cblocks], c0 = (it % cblocks) * kTmKB;
            const uint32_t a_s = tm_smem_u32(stage0 + (size_t)s * stage_bytes), b_s = a_s + a_bytes;
            const uint32_t full = tm_smem_u32(&H.full_bar[s]);
            if (lane == 0) tm_expect_tx(full, (uint32_t)(kGatherA ? stage_bytes : b_bytes));
            __syncwarp();
            if (kGatherA) {
                const int4 idx = *reinterpret_cast<const int4 *>(s_nbr + k * kTmM + 4 * lane);   // -1 = out of bounds -> zero rows
                tma_gather4(a_s + lane * 512, &map_a, full, c0, idx.x, idx.y, idx.z, idx.w);
            }
            if (lane == 0) tma_load_2d(b_s, &map_b, full, c0, k * cout_total + n0);
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(cout >> 3) << 17) | ((uint32_t)(kTmM >> 4) << 24);
            for (int it = 0; it < total; it++) {
                const int s = it % kStages;
                tm_mbar_wait(tm_smem_u32(&H.full_bar[s]), (uint32_t)((it / kStages) & 1));
                if (!kGatherA) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async (generic proxy) -> tensor core
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_s = tm_smem_u32(stage0 + (size_t)s * stage_bytes), b_s = a_s + a_bytes;
                const uint64_t a_desc = tm_desc_k_sw128(a_s), b_desc = tm_desc_k_sw128(b_s);
#pragma unroll
                for (int kk = 0; kk < kTmKB / 8; kk++)
                    tm_umma_tf32(tmem, a_desc + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2), idesc, (it > 0 || kk > 0) ? 1u : 0u);
                tm_commit(tm_smem_u32(&H.empty_bar[s]));
            }
            if (total > 0) tm_commit(tm_smem_u32(&H.accum_bar));
        }
    } else {
        if (!kGatherA) {
            // ===== A producers (the epilogue warps, idle until the accumulator is complete): thread (rb, c) copies the
            // 16-byte chunk c of rows rb + 16 j of every k-block
            const int pt = tid - 64, c = pt & 7, rb = pt >> 3;
            uint32_t t_off[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int r = rb + 16 * j;
                t_off[j] = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
            }
            const float *a_src[8];
            uint32_t a_ok[8];
            for (int it = 0; it < total; it++) {
                const int s = it % kStages, cb = it % cblocks;
                if (cb == 0) {
                    const int k = H.act[it / cblocks];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int idx = s_nbr[k * kTmM + rb + 16 * j];
                        a_src[j] = in + (size_t)(idx >= 0 ? idx : 0) * ldi + c * 4;
                        a_ok[j] = idx >= 0 ? 16u : 0u;
                    }
                }
                tm_mbar_wait(tm_smem_u32(&H.empty_bar[s]), (uint32_t)(((it / kStages) & 1) ^ 1));
                const uint32_t a_s = tm_smem_u32(stage0 + (size_t)s * stage_bytes);
#pragma unroll
                for (int j = 0; j < 8; j++)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(a_s + t_off[j]), "l"(a_src[j] + cb * kTmKB), "r"(a_ok[j]) : "memory");
                // the hardware arrives on the stage's full barrier when this thread's copies have landed (pre-counted in
                // the barrier's arrival count): the producer never waits for its own data
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tm_smem_u32(&H.full_bar[s])) : "memory");
            }
        }
        // ===== epilogue warps 2..5: TMEM lane quarter (warp & 3) -> registers -> (+bias, +residual, relu) -> global
        if (total > 0) {
            tm_mbar_wait(tm_smem_u32(&H.accum_bar), 0u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const int q = warp & 3;
        const int r = row0 + q * 32 + lane;
        const bool add_bias = bias != nullptr && blockIdx.z == 0;
        for (int cb = 0; cb < cout / 16; cb++) {
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
            if (r < n_out && (total > 0 || gridDim.z == 1 || add_bias)) {
                float *dst = out + (size_t)r * ldo + n0 + cb * 16;
                const float *res = residual ? residual + (size_t)r * ldr + n0 + cb * 16 : nullptr;
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
                    else
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * j), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
}

int launch_finish(float *d_out, int ldo, int64_t n_out, int cout, const float *d_bias, const float *d_res, int ldr, int relu,
                  cudaStream_t stream);   // sparse_conv_tc.cu

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int make_map_2d(CUtensorMap *m, const float *base, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes, uint32_t box_cols,
                uint32_t box_rows, int swizzle_atom_32b) {
    EncodeTiledFn fn = encode_tiled();
    CVB_REQUIRE(fn != nullptr, CVB200_EINVAL, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {cols, rows}, strides[1] = {row_stride_bytes};
    const cuuint32_t box[2] = {box_cols, box_rows}, estr[2] = {1, 1};
    const CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swizzle_atom_32b ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CVB_REQUIRE(rc == CUDA_SUCCESS, CVB200_EINVAL, "cuTensorMapEncodeTiled failed (CUresult %d; %llu x %llu, stride %llu, box %u x %u)",
                (int)rc, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)row_stride_bytes, box_cols, box_rows);
    return 0;
}

int launch_conv_tma(int gather_a, const float *d_in, int64_t n_in, int ldi, int cin, const float *d_wt, int cout, const int32_t *d_nbr,
                    int64_t n_out, int k3, const float *d_bias, const float *d_res, int ldr, int relu, float *d_out, int ldo,
                    cudaStream_t stream) {
    CVB_REQUIRE(cin > 0 && cin % kTmKB == 0 && cout >= 16 && cout <= 256 && cout % 16 == 0 && k3 > 0 && k3 <= kTmMaxK3, CVB200_EINVAL,
                "sc_conv_forward_tc: needs cin %% 32 == 0, cout %% 16 == 0, 16 <= cout <= 256, K^3 <= %d (got %d, %d, %d)", kTmMaxK3, cin,
                cout, k3);
    CVB_REQUIRE(n_out >= 0 && n_out < (1LL << 31) && n_in > 0 && n_in < (1LL << 31), CVB200_EINVAL, "sc_conv_forward_tc: bad n_out / n_in");
    if (n_out == 0) return 0;
    CVB_REQUIRE(d_in && d_wt && d_nbr && d_out, CVB200_EINVAL, "sc_conv_forward_tc: NULL argument");
    CVB_REQUIRE(((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_wt) | reinterpret_cast<uintptr_t>(d_out) |
                  reinterpret_cast<uintptr_t>(d_res)) & 15) == 0 && ldi % 4 == 0 && ldo % 4 == 0 && ldr % 4 == 0,
                CVB200_EINVAL, "sc_conv_forward_tc: 16-byte aligned pointers and row strides required");
    const int m_tiles = (int)ceil_div(n_out, kTmM);
    int n_splits = 1, k_splits = 1;
    if (m_tiles < kNumSMs / 2 && cout >= 128 && cout % 64 == 0) n_splits = cout / 64;
    const int kmul = (k3 % 3 == 0) ? 3 : 2, cblocks = cin / kTmKB;
    while (m_tiles * n_splits * k_splits < kNumSMs / 2 && k_splits * kmul <= k3 && (k3 / (k_splits * kmul)) * cblocks >= 8) k_splits *= kmul;
    const int nc = cout / n_splits;
    int tmem_cols = 32;
    while (tmem_cols < nc) tmem_cols <<= 1;
    alignas(64) CUtensorMap map_a, map_b;
    if (int rc = make_map_2d(&map_a, d_in, (uint64_t)cin, (uint64_t)n_in, (uint64_t)ldi * 4, kTmKB, 1)) return rc;   // tile::gather4: box = one row x 128 B, four row coordinates per instruction
    if (int rc = make_map_2d(&map_b, d_wt, (uint64_t)cin, (uint64_t)k3 * cout, (uint64_t)cin * 4, kTmKB, (uint32_t)nc)) return rc;
    const int nbr_bytes = ((k3 * kTmM * 4 + 1023) / 1024) * 1024;
    const int stage_bytes = kTmM * 128 + nc * 128;
    const dim3 grid((unsigned)m_tiles, (unsigned)n_splits, (unsigned)k_splits);
    const bool split = k_splits > 1;
    if (split) CVB_CUDA(cudaMemset2DAsync(d_out, sizeof(float) * (size_t)ldo, 0, sizeof(float) * (size_t)cout, (size_t)n_out, stream));
    const float *k_bias = split ? nullptr : d_bias, *k_res = split ? nullptr : d_res;
    const int k_relu = split ? 0 : relu;
    const size_t smem = 1024 + nbr_bytes + (nc <= 64 ? 4 : 3) * (size_t)stage_bytes;
#define CVB_TMA_LAUNCH(S, G)                                                                                                      \
    do {                                                                                                                          \
        static bool set = false;                                                                                                  \
        if (!set) {                                                                                                               \
            CVB_CUDA(cudaFuncSetAttribute(sc_conv_tma_kernel<S, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));     \
            set = true;                                                                                                           \
        }                                                                                                                         \
        sc_conv_tma_kernel<S, G><<<grid, kTmThreads, smem, stream>>>(map_a, map_b, d_in, ldi, cin, cout, d_nbr, (int)n_out, k3,    \
                                                                     k_bias, k_res, ldr, k_relu, d_out, ldo, tmem_cols, nc);      \
    } while (0)
    if (nc <= 64) {   // 4 stages of 24 KiB / 3 stages of <= 48 KiB: two CTAs per SM up to 128 channels per CTA
        if (gather_a) CVB_TMA_LAUNCH(4, true); else CVB_TMA_LAUNCH(4, false);
    } else {
        if (gather_a) CVB_TMA_LAUNCH(3, true); else CVB_TMA_LAUNCH(3, false);
    }
#undef CVB_TMA_LAUNCH
    CVB_LAUNCH_CHECK("sc_conv_tma_kernel");
    if (split && (d_bias || d_res || relu)) return launch_finish(d_out, ldo, n_out, cout, d_bias, d_res, ldr, relu, stream);
    return 0;
}

}  // namespace cvb200
