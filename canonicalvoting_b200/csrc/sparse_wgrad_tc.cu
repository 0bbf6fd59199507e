// canonicalvoting_b200/csrc/sparse_wgrad_tc.cu -- weight gradient of the sparse convolution on the tensor cores.
//
//     dW[k][ci][co] = sum over output rows o of  x[table[o,k], ci] * dout[o, co]          (train_joint.py:277 loss.backward())
//
// The fp32 CUDA-core kernel (sparse_conv.cu: 64x64 tiles, 16-row steps, zero rows multiplied like any other) took 128 of
// the 167 ms of a training step (8 x 50k-voxel scenes, tools/profile_train.py).  Here the contraction runs over the ROW
// dimension on tcgen05: per kernel offset k a GEMM  dW[k] (M = cin, N = cout) += X_k^T (M x K) * dOut (K x N),  K = rows.
//   * Both operands are "MN-major" for this product (a feature row holds consecutive channels), which kind::tf32 accepts
//     on sm_100 (instruction-descriptor bits 15/16): a stage holds 32 rows; X rows gathered through the neighbour table
//     land as [channel block of 32][32 rows][128 B] in the 128-byte swizzle with 32-byte atoms (the one layout MN-major
//     tf32 operands may have), written by the same kind of gather producers as the forward kernel's, and dOut rows arrive
//     by TMA boxes of 32 channels x 32 rows (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).  Four MMAs (M=128, N=cout, K=8)
//     consume a stage.
//   * The accumulators of several offsets live side by side in tensor memory (512 columns / cout offsets per unit), so a
//     unit = (row chunk, 128-channel tile of cin, offset group) streams its rows once per offset of the group.
//   * Units are spread over one persistent CTA per SM; partial dW tiles of different row chunks are combined with
//     red.global.add.v4.f32 into the (small, L2-resident) gradient.
// Roles as in sparse_conv_persist.cu: warp 0 = TMA (dOut tiles), warp 1 = MMA issuer + TMEM owner, warps 2-5 = gather
// producers (one whole stage per warp, completion by cp.async.mbarrier.arrive.noinc), warps 6-9 = epilogue.
#include <cuda.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace cvb200 {

constexpr int kWtRows = 32;          // rows per stage (K of a stage)
constexpr int kWtM = 128;            // channels of cin per unit (UMMA M)
constexpr int kWtThreads = 320;
constexpr int kWtStages = 6;
constexpr int kWtABytes = 4 * kWtRows * 128;   // 4 channel blocks x 32 rows x 128 B = 16 KiB

struct WtHeader {
    unsigned long long full_bar[kWtStages], empty_bar[kWtStages], acc_full, acc_empty;
    unsigned int tmem_base;
};

struct WtPlan {
    int n_chunks, rows_per_chunk;   // row chunks (multiple of 32 rows)
    int m_tiles;                    // ceil(cin / 128)
    int group, n_groups;            // offsets per unit, ceil(k3 / group)
    int n_units;
    int nb;                         // cout / 32: TMA boxes per stage
    int tmem_cols;
    int stages;                     // ring depth (<= kWtStages, as many as fit)
};

// shared-memory matrix descriptor, MN-major tf32: the only layout the hardware takes is SWIZZLE_128B_BASE32B (layout type 1;
// cute: Layout_MN_SW128_32B_Atom): 32 elements (128 B) contiguous along M/N, atoms of 4 K-rows (512 B), the 32-byte chunk
// index of a row XOR-ed with (row & 3).  Leading byte offset = distance between blocks of 32 elements along M/N, stride byte
// offset = distance between the 4-row groups along K (an MMA of K = 8 spans two).
__device__ __forceinline__ uint64_t wt_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffff) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}

// dynamic smem: [header 1 KiB][stages x (A 16 KiB | B nb x 4 KiB)]
__global__ void __launch_bounds__(kWtThreads, 1)
sc_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_b, const float *__restrict__ x, int ldx, int cin, int cout,
                   const int *__restrict__ table, int n_rows, int k3, float *__restrict__ dw, const WtPlan P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    WtHeader &H = *reinterpret_cast<WtHeader *>(smem);
    unsigned char *stage0 = smem + 1024;
    const int b_bytes = P.nb * kWtRows * 128, stage_bytes = kWtABytes + b_bytes;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < P.stages; s++) {
            tm_mbar_init(tm_smem_u32(&H.full_bar[s]), 1 + 32);
            tm_mbar_init(tm_smem_u32(&H.empty_bar[s]), 1);
        }
        tm_mbar_init(tm_smem_u32(&H.acc_full), 1);
        tm_mbar_init(tm_smem_u32(&H.acc_empty), 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tm_smem_u32(&H.tmem_base)), "r"(P.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = H.tmem_base;

    // unit u -> (row chunk, cin tile, offset group); the stage sequence of a unit: for each 32-row sub-tile, for each offset
    auto unit = [&](int u, int &r_begin, int &r_end, int &m0, int &k0, int &k1) {
        const int g = u % P.n_groups, t = (u / P.n_groups) % P.m_tiles, c = u / (P.n_groups * P.m_tiles);
        r_begin = c * P.rows_per_chunk;
        r_end = min(n_rows, r_begin + P.rows_per_chunk);
        m0 = t * kWtM;
        k0 = g * P.group;
        k1 = min(k3, k0 + P.group);
    };

    if (warp == 0) {
        // ===== dOut tiles by TMA: nb boxes of [32 channels x 32 rows] per stage (rows beyond n_rows are zero-filled)
        int s = 0;
        uint32_t ph = 0;
        for (int u = blockIdx.x; u < P.n_units; u += gridDim.x) {
            int r_begin, r_end, m0, k0, k1;
            unit(u, r_begin, r_end, m0, k0, k1);
            for (int r0 = r_begin; r0 < r_end; r0 += kWtRows)
                for (int k = k0; k < k1; k++) {
                    tm_mbar_wait(tm_smem_u32(&H.empty_bar[s]), ph ^ 1u);
                    const uint32_t b_s = tm_smem_u32(stage0 + (size_t)s * stage_bytes) + kWtABytes;
                    const uint32_t full = tm_smem_u32(&H.full_bar[s]);
                    if (tm_elect_one()) {
                        tm_expect_tx(full, (uint32_t)b_bytes);
                        for (int j = 0; j < P.nb; j++) tma_load_2d(b_s + j * (kWtRows * 128), &map_b, full, j * 32, r0);
                    }
                    __syncwarp();
                    if (++s == P.stages) { s = 0; ph ^= 1u; }
                }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: accumulator (k - k0) at TMEM column (k - k0) * cout
        // instruction descriptor: fp32 accumulate, tf32 x tf32, A and B MN-major (bits 15, 16), N = cout, M = 128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(cout >> 3) << 17) |
                               ((uint32_t)(kWtM >> 4) << 24);
        int s = 0, li = 0;
        uint32_t ph = 0;
        for (int u = blockIdx.x; u < P.n_units; u += gridDim.x, li++) {
            int r_begin, r_end, m0, k0, k1;
            unit(u, r_begin, r_end, m0, k0, k1);
            tm_mbar_wait(tm_smem_u32(&H.acc_empty), (uint32_t)((li & 1) ^ 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int r0 = r_begin; r0 < r_end; r0 += kWtRows)
                for (int k = k0; k < k1; k++) {
                    tm_mbar_wait(tm_smem_u32(&H.full_bar[s]), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_s = tm_smem_u32(stage0 + (size_t)s * stage_bytes), b_s = a_s + kWtABytes;
                    const uint32_t d_tmem = tmem + (uint32_t)((k - k0) * cout);
                    if (tm_elect_one()) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async (generic proxy) -> tensor core
#pragma unroll
                        for (int kk = 0; kk < kWtRows / 8; kk++) {      // 8 rows per MMA: one 1024-byte swizzle atom per channel block
                            const uint64_t a_desc = wt_desc_mn_sw128(a_s + kk * 1024, kWtRows * 128);
                            const uint64_t b_desc = wt_desc_mn_sw128(b_s + kk * 1024, kWtRows * 128);
                            tm_umma_tf32(d_tmem, a_desc, b_desc, idesc, (r0 > r_begin || kk > 0) ? 1u : 0u);
                        }
                        tm_commit(tm_smem_u32(&H.empty_bar[s]));
                    }
                    __syncwarp();
                    if (++s == P.stages) { s = 0; ph ^= 1u; }
                }
            if (tm_elect_one()) tm_commit(tm_smem_u32(&H.acc_full));
            __syncwarp();
        }
    } else if (warp < 6) {
        // ===== gather producers: warp w fills the stages n % 4 == w alone.  Lane (rb, c8): 16-byte chunk c8 of channel
        // block j of rows rb + 4 i  ->  smem j * 4096 + row * 128 + 32-byte chunk ((c8 >> 1) ^ (row & 3)) + 16 (c8 & 1)
        const int w = warp - 2, c8 = lane & 7, rb = lane >> 3;
        int n = 0;
        for (int u = blockIdx.x; u < P.n_units; u += gridDim.x) {
            int r_begin, r_end, m0, k0, k1;
            unit(u, r_begin, r_end, m0, k0, k1);
            const int nblk = min(4, (cin - m0) / 32);          // channel blocks of this tile that exist
            for (int r0 = r_begin; r0 < r_end; r0 += kWtRows)
                for (int k = k0; k < k1; k++, n++) {
                    if ((n & 3) != w) continue;
                    // neighbour id of row r0 + lane (one coalesced-ish load per stage), distributed by shuffles
                    const int mine = r0 + lane < r_end ? __ldg(table + (size_t)(r0 + lane) * k3 + k) : -1;
                    const int round = n / P.stages, s = n - round * P.stages;
                    if (lane == 0) tm_mbar_wait(tm_smem_u32(&H.empty_bar[s]), (uint32_t)((round & 1) ^ 1));
                    __syncwarp();
                    const uint32_t a_s = tm_smem_u32(stage0 + (size_t)s * stage_bytes);
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int row = rb + 4 * i;
                        const int id = __shfl_sync(0xffffffffu, mine, row);
                        const float *src = x + (size_t)(id >= 0 ? id : 0) * ldx + m0 + c8 * 4;
                        const uint32_t dst = a_s + (uint32_t)(row * 128 + (((c8 >> 1) ^ (row & 3)) << 5) + ((c8 & 1) << 4));
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            if (j < nblk)
                                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + j * 4096), "l"(src + j * 32), "r"(id >= 0 ? 16u : 0u) : "memory");
                    }
                    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tm_smem_u32(&H.full_bar[s])) : "memory");
                }
        }
    } else {
        // ===== epilogue warps 6..9: lane quarter (warp & 3) = channels m0 + 32 q + lane of cin; every accumulator of the group
        const int q = warp & 3;
        int li = 0;
        for (int u = blockIdx.x; u < P.n_units; u += gridDim.x, li++) {
            int r_begin, r_end, m0, k0, k1;
            unit(u, r_begin, r_end, m0, k0, k1);
            if (lane == 0) tm_mbar_wait(tm_smem_u32(&H.acc_full), (uint32_t)(li & 1));
            __syncwarp();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int ci = m0 + q * 32 + lane;
            for (int k = k0; k < k1; k++) {
                float *dst = dw + ((size_t)k * cin + ci) * cout;
                for (int cb = 0; cb < cout / 16; cb++) {
                    uint32_t v[16];
                    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((k - k0) * cout + cb * 16);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (ci < cin) {
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + cb * 16 + 4 * j), "f"(__uint_as_float(v[4 * j])),
                                         "f"(__uint_as_float(v[4 * j + 1])), "f"(__uint_as_float(v[4 * j + 2])), "f"(__uint_as_float(v[4 * j + 3])) : "memory");
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tm_smem_u32(&H.acc_empty)) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(P.tmem_cols) : "memory");
}

// dw must be zero-filled by the caller (partial tiles are added)
int launch_wgrad_tc(const float *d_x, int cin, const float *d_dout, int cout, const int32_t *d_table, int64_t n_rows, int k3, float *d_dw,
                    cudaStream_t stream) {
    WtPlan P;
    P.nb = cout / 32;
    P.m_tiles = (int)ceil_div(cin, kWtM);
    P.group = 512 / cout;
    if (P.group > k3) P.group = k3;
    P.n_groups = (int)ceil_div(k3, P.group);
    int cols = 32;
    while (cols < P.group * cout) cols <<= 1;
    P.tmem_cols = cols;
    // row chunks: about two units per SM, at least 8 stages of rows per unit
    const int64_t per = (int64_t)P.m_tiles * P.n_groups;
    int64_t chunks = ceil_div(2 * kNumSMs, per);
    int64_t rows = ceil_div(ceil_div(n_rows, chunks), kWtRows) * kWtRows;
    if (rows < 8 * kWtRows) rows = 8 * kWtRows;
    P.rows_per_chunk = (int)rows;
    P.n_chunks = (int)ceil_div(n_rows, rows);
    P.n_units = (int)(P.n_chunks * per);
    alignas(64) CUtensorMap map_b;
    if (int rc = make_map_2d(&map_b, d_dout, (uint64_t)cout, (uint64_t)n_rows, (uint64_t)cout * 4, 32, kWtRows, 1)) return rc;
    const int stage_bytes = kWtABytes + P.nb * kWtRows * 128;
    P.stages = (223 * 1024) / stage_bytes;
    if (P.stages > kWtStages) P.stages = kWtStages;
    const size_t smem = 1024 + (size_t)P.stages * stage_bytes;
    static DeviceOnce once;
    if (!once.done()) {
        CVB_CUDA(cudaFuncSetAttribute(sc_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        once.mark();
    }
    const int grid = P.n_units < kNumSMs ? P.n_units : kNumSMs;
    sc_wgrad_tc_kernel<<<grid, kWtThreads, smem, stream>>>(map_b, d_x, cin, cin, cout, (const int *)d_table, (int)n_rows, k3, d_dw, P);
    CVB_LAUNCH_CHECK("sc_wgrad_tc_kernel");
    return 0;
}

}  // namespace cvb200

using namespace cvb200;

extern "C" int cvb200_sc_conv_wgrad_tc(const float *d_x, int32_t cin, const float *d_dout, int32_t cout, const int32_t *d_table,
                                       int64_t n_rows, int32_t k3, float *d_dw, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVB_REQUIRE(cin >= 32 && cin % 32 == 0 && cout >= 32 && cout % 32 == 0 && cout <= 256 && k3 > 0 && n_rows >= 0 && n_rows < (1LL << 31),
                CVB200_EINVAL, "sc_conv_wgrad_tc: needs cin %% 32 == 0, cout %% 32 == 0, cout <= 256 (got %d, %d)", cin, cout);
    CVB_REQUIRE(d_dw, CVB200_EINVAL, "sc_conv_wgrad_tc: NULL dw");
    CVB_CUDA(cudaMemsetAsync(d_dw, 0, sizeof(float) * (size_t)k3 * cin * cout, stream));
    if (n_rows == 0) return 0;
    CVB_REQUIRE(d_x && d_dout && d_table, CVB200_EINVAL, "sc_conv_wgrad_tc: NULL argument");
    CVB_REQUIRE(((reinterpret_cast<uintptr_t>(d_x) | reinterpret_cast<uintptr_t>(d_dout) | reinterpret_cast<uintptr_t>(d_dw)) & 15) == 0,
                CVB200_EINVAL, "sc_conv_wgrad_tc: 16-byte aligned pointers required");
    return launch_wgrad_tc(d_x, cin, d_dout, cout, d_table, n_rows, k3, d_dw, stream);
}
