// canonicalvoting_b200/csrc/sparse_conv_bf16.cu -- EXPERIMENTAL bf16 variant of the persistent tcgen05 sparse convolution
// (sparse_conv_persist.cu).  Written at the end of round 1 after the GPU budget was spent: it compiles for sm_100a, its index
// arithmetic is checked on the CPU against the convolution oracle (tests/test_conv_bf16_layout.py emulates the tile
// assembly), but it HAS NOT RUN ON A GPU yet.  Nothing on the product path calls it; tools/try_bf16_conv.py is the first
// thing to run in round 2.
//
//     out[o, n0:n0+nc) = [relu]( sum_k in[nbr[o,k], :] @ W[k][:, n0:n0+nc) + bias + residual )      in / W / residual / out: bf16
//
// Why: the TF32 kernel is bound by the gather warps' instruction stream (~550 cycles per k-block of 128 rows x 32 channels
// against ~400 of MMA time).  With bf16 operands a 128-byte row holds 64 channels: half the k-blocks, half
// the cp.async copies and hand-shakes per FLOP, twice the MMA rate (kind::f16), same shared-memory layout and descriptors.
//
// Differences from sparse_conv_persist.cu (everything else -- roles, ring, two TMEM accumulators, split tiles, PDL -- is the
// same protocol):
//   * the contraction runs over the FLATTENED (kernel offset, channel) axis of length K^3 * cin, cut into k-blocks of 64
//     elements.  Chunk c (16 bytes = 8 channels) of k-block j is flat index 64 j + 8 c -> offset k = flat / cin, channel
//     flat % cin; cin % 32 == 0, so a chunk never straddles an offset and a k-block touches at most two offsets (exactly one
//     when cin % 64 == 0): 32- and 96-channel layers need no padding.  Chunks at or beyond K^3 * cin are zero-filled.
//   * weights are packed [cout][K^3 * cin] (bf16), so the B tile of k-block j is the TMA box at column 64 j, rows n0..n0+nc;
//     the tail beyond K^3 * cin is zero-filled by TMA's out-of-bounds rule.
//   * four tcgen05.mma.kind::f16 (bf16 x bf16 -> fp32, M=128, N=nc, K=16) per k-block; +32-byte descriptor step as before.
//   * the gather warp keeps the neighbour ids of TWO offsets (k_lo, k_lo + 1) per k-block and prefetches the pair of its next
//     k-block; a k-block inside one offset (warp-uniform test) needs one id shuffle per row, a straddling one two.
//   * epilogue: fp32 accumulators -> (+bias, +residual (bf16), relu) -> bf16 (8-byte stores, 64 contiguous bytes per row and
//     warp instruction) or fp32 (`out_f32`, for the last layer whose output feeds the head decode).
#include <cuda.h>
#include <cuda_bf16.h>

#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"
#include "tcgen05.cuh"

namespace cvb200 {

constexpr int kBfM = 128, kBfKB = 64 /* bf16 elements = 128 bytes */, kBfThreads = 416, kBfMaxStages = 8;
constexpr int kBfProducers = 4;
constexpr int kBfSmemBytes = 224 * 1024;
constexpr int kBfMaxSplitTiles = 2 * kNumSMs;
constexpr size_t kBfScratchFloats = (size_t)kBfMaxSplitTiles * kBfM * 128;

struct BfHeader {
    unsigned long long full_bar[kBfMaxStages], empty_bar[kBfMaxStages], acc_full[2], acc_empty[2], turn[2];
    unsigned int tmem_base;
    int last_flag;
};

struct BfPlan {
    int n_tiles, n_splits, n_whole, ks, n_units;
    int total_kb;     // ceil(K^3 * cin / 64)
    int stages, nc, acc_stride, tmem_cols;
};

struct BfUnit {
    int row0, n0, kb0, kb1, pieces, split_tile;
};

__host__ __device__ __forceinline__ BfUnit bf_unit(const BfPlan &P, int u) {
    int tile, piece, pieces;
    if (u < P.n_whole) {
        tile = u; piece = 0; pieces = 1;
    } else {
        const int v = u - P.n_whole;
        tile = P.n_whole + v / P.ks; piece = v % P.ks; pieces = P.ks;
    }
    BfUnit U;
    U.row0 = (tile / P.n_splits) * kBfM;
    U.n0 = (tile % P.n_splits) * P.nc;
    U.kb0 = (int)((long long)piece * P.total_kb / pieces);
    U.kb1 = (int)((long long)(piece + 1) * P.total_kb / pieces);
    U.pieces = pieces;
    U.split_tile = tile - P.n_whole;
    return U;
}

// Where chunk c (16 bytes = 8 channels) of k-block `it` comes from on the flattened (offset, channel) axis.  k_lo / single are
// warp-uniform (they depend on `it` only): the k-block touches offsets k_lo and, unless `single`, k_lo + 1.
struct BfChunk {
    int k_lo, single, valid, use_hi, ch;
};
__host__ __device__ __forceinline__ BfChunk bf_chunk(int it, int c, int cin, int ktot) {
    BfChunk r;
    r.k_lo = (it * kBfKB) / cin;
    r.single = (it * kBfKB + kBfKB - 1) / cin == r.k_lo;
    const int flat = it * kBfKB + 8 * c;
    const int k_mine = flat / cin;
    r.valid = flat < ktot;
    r.use_hi = k_mine != r.k_lo;
    r.ch = flat - k_mine * cin;
    return r;
}

__device__ __forceinline__ void bf_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void bf_umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ uint2 bf_pack4(float4 o) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
    uint2 r;
    r.x = *reinterpret_cast<const uint32_t *>(&lo);
    r.y = *reinterpret_cast<const uint32_t *>(&hi);
    return r;
}
__device__ __forceinline__ float4 bf_unpack4(uint2 v) {
    const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162 *>(&v.x), hi = *reinterpret_cast<const __nv_bfloat162 *>(&v.y);
    const float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
    return make_float4(a.x, a.y, b.x, b.y);
}

// dynamic smem: [header 1 KiB][epilogue staging: 4 warps x 32 rows x 128 B][stages x (A 16 KiB | B nc x 128 B)]
__global__ void __launch_bounds__(kBfThreads, 1)
sc_conv_bf16_kernel(const __grid_constant__ CUtensorMap map_b, const __nv_bfloat16 *__restrict__ in, int ldi, int cin,
                    const int *__restrict__ nbr, int n_out, int k3, const float *__restrict__ bias,
                    const __nv_bfloat16 *__restrict__ residual, int ldr, int relu, void *__restrict__ out_, int ldo, int out_f32,
                    const BfPlan P, float *__restrict__ scratch, int *__restrict__ counters) {
    extern __shared__ __align__(1024) unsigned char smem[];
    BfHeader &H = *reinterpret_cast<BfHeader *>(smem);
    unsigned char *stage0 = smem + 1024 + 16384;
    const int a_bytes = kBfM * 128, b_bytes = P.nc * 128, stage_bytes = a_bytes + b_bytes;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ktot = k3 * cin;                          // length of the flattened contraction axis

    if (tid == 0) {
        for (int s = 0; s < P.stages; s++) {
            tm_mbar_init(tm_smem_u32(&H.full_bar[s]), 1 + 32);
            tm_mbar_init(tm_smem_u32(&H.empty_bar[s]), 1);
        }
        for (int b = 0; b < 2; b++) {
            tm_mbar_init(tm_smem_u32(&H.acc_full[b]), 2);      // both MMA warps
            tm_mbar_init(tm_smem_u32(&H.acc_empty[b]), 4);
            tm_mbar_init(tm_smem_u32(&H.turn[b]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tm_smem_u32(&H.tmem_base)), "r"(P.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = H.tmem_base;
    // programmatic dependent launch: see sparse_conv_persist.cu (only the gather and epilogue roles wait for the previous kernel)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        // ===== weight producer: the [nc x 64] block of the packed weights at column 64 * it, one TMA load per k-block
        int s = 0;
        uint32_t ph = 0;
        for (int u = blockIdx.x; u < P.n_units; u += gridDim.x) {
            const BfUnit U = bf_unit(P, u);
            for (int it = U.kb0; it < U.kb1; it++) {
                tm_mbar_wait(tm_smem_u32(&H.empty_bar[s]), ph ^ 1u);
                const uint32_t b_s = tm_smem_u32(stage0 + (size_t)s * stage_bytes) + a_bytes;
                const uint32_t full = tm_smem_u32(&H.full_bar[s]);
                if (tm_elect_one()) {
                    tm_expect_tx(full, (uint32_t)b_bytes);
                    tma_load_2d(b_s, &map_b, full, it * kBfKB, U.n0);
                }
                __syncwarp();
                if (++s == P.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 4 || warp == 12) {
        // ===== MMA issuers: two warps alternating k-blocks under a strict issue order (turn barriers), as in the TF32 kernel
        const int me = warp == 4 ? 0 : 1;
        int li = 0, n_base = 0;
        // instruction descriptor: D = f32 (bit 4), A and B = bf16 (format 1 at bits 7 and 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P.nc >> 3) << 17) | ((uint32_t)(kBfM >> 4) << 24);
        for (int u = blockIdx.x; u < P.n_units; u += gridDim.x, li++) {
            const BfUnit U = bf_unit(P, u);
            const int buf = li & 1;
            const int first = U.kb0 + ((me - n_base) & 1);        // my first k-block of this unit
            const uint32_t d_tmem = tmem + (uint32_t)(buf * P.acc_stride);
            if (first == U.kb0 && first < U.kb1) {
                tm_mbar_wait(tm_smem_u32(&H.acc_empty[buf]), (uint32_t)(((li >> 1) & 1) ^ 1));   // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            for (int it = first; it < U.kb1; it += 2) {
                const int n = n_base + it - U.kb0;
                const int round = n / P.stages, s = n - round * P.stages;
                tm_mbar_wait(tm_smem_u32(&H.full_bar[s]), (uint32_t)(round & 1));
                const uint32_t a_s = tm_smem_u32(stage0 + (size_t)s * stage_bytes), b_s = a_s + a_bytes;
                const uint64_t a_desc = tm_desc_k_sw128(a_s), b_desc = tm_desc_k_sw128(b_s);
                if (lane == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async (generic proxy) -> tensor core
                if (n > 0) tm_mbar_wait(tm_smem_u32(&H.turn[me]), (uint32_t)(((n >> 1) + me + 1) & 1));   // the other warp has issued k-block n - 1
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tm_elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < kBfKB / 16; kk++)      // K = 16 bf16 = 32 bytes per instruction
                        bf_umma(d_tmem, a_desc + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2), idesc, (it > U.kb0 || kk > 0) ? 1u : 0u);
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tm_smem_u32(&H.turn[me ^ 1])) : "memory");
                    tm_commit(tm_smem_u32(&H.empty_bar[s]));
                }
                __syncwarp();
            }
            if (tm_elect_one()) {
                if (first < U.kb1) tm_commit(tm_smem_u32(&H.acc_full[buf]));       // arrives when my MMAs of the unit are complete
                else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tm_smem_u32(&H.acc_full[buf])) : "memory");
            }
            __syncwarp();
            n_base += U.kb1 - U.kb0;
        }
    } else if (warp < 8) {
        // ===== gather producers (warps 1,2,3,5; 6,7 idle): warp w fills the k-blocks n % W == w of the CTA's running sequence
        // alone.  Lane (rb, c) copies the 16-byte chunk c of rows rb + 4 j, j < 32.
        const int w = warp < 4 ? warp - 1 : warp - 2, c = lane & 7, rb = lane >> 3;
        const int W = P.stages < kBfProducers ? P.stages : kBfProducers;
        const uint32_t off_even = (uint32_t)(rb * 128 + ((c ^ rb) << 4)), off_odd = (uint32_t)((rb + 4) * 128 + ((c ^ (rb + 4)) << 4));
        const size_t ld = (size_t)ldi;
        int n_base = 0;
        bool dep_done = false;
        for (int u = blockIdx.x; u < P.n_units && w < W; u += gridDim.x) {
            const BfUnit U = bf_unit(P, u);
            // neighbour ids: lane l holds the ids of rows l, l + 32, l + 64, l + 96 for the two offsets a k-block can touch
            const int *nlane = nbr + (size_t)(U.row0 + lane) * k3;
            const int lane_rows = n_out - U.row0 - lane;          // row lane + 32 m exists iff 32 m < lane_rows
            int lo[4], hi[4], nlo[4], nhi[4];
            int it = U.kb0 + (((w - n_base) % W) + W) % W;
            if (it < U.kb1) {
                const int k0 = (it * kBfKB) / cin;
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    nlo[m] = (32 * m < lane_rows && k0 < k3) ? __ldg(nlane + (size_t)(32 * m) * k3 + k0) : -1;
                    nhi[m] = (32 * m < lane_rows && k0 + 1 < k3) ? __ldg(nlane + (size_t)(32 * m) * k3 + k0 + 1) : -1;
                }
            }
            for (; it < U.kb1; it += W) {
#pragma unroll
                for (int m = 0; m < 4; m++) { lo[m] = nlo[m]; hi[m] = nhi[m]; }
                if (it + W < U.kb1) {       // ids of my next k-block: in flight while this one is copied
                    const int kn = ((it + W) * kBfKB) / cin;
#pragma unroll
                    for (int m = 0; m < 4; m++) {
                        nlo[m] = (32 * m < lane_rows && kn < k3) ? __ldg(nlane + (size_t)(32 * m) * k3 + kn) : -1;
                        nhi[m] = (32 * m < lane_rows && kn + 1 < k3) ? __ldg(nlane + (size_t)(32 * m) * k3 + kn + 1) : -1;
                    }
                }
                const BfChunk mine = bf_chunk(it, c, cin, ktot);            // my chunk on the flattened (offset, channel) axis
                const bool valid = mine.valid != 0, use_hi = mine.use_hi != 0;
                const bool single = mine.single != 0;                       // warp-uniform: the k-block lies inside one offset
                const int ch = mine.ch;
                const int n = n_base + it - U.kb0;
                const int round = n / P.stages, s = n - round * P.stages;
                if (!dep_done) {
                    asm volatile("griddepcontrol.wait;" ::: "memory");
                    dep_done = true;
                }
                if (lane == 0) tm_mbar_spin(tm_smem_u32(&H.empty_bar[s]), (uint32_t)((round & 1) ^ 1));
                __syncwarp();
                const uint32_t a_s = tm_smem_u32(stage0 + (size_t)s * stage_bytes);
                const __nv_bfloat16 *src0 = in + ch;
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    int id = __shfl_sync(0xffffffffu, lo[j >> 3], rb + 4 * (j & 7));   // row rb + 4 j = 32 (j >> 3) + rb + 4 (j & 7)
                    if (!single) {
                        const int idh = __shfl_sync(0xffffffffu, hi[j >> 3], rb + 4 * (j & 7));
                        id = use_hi ? idh : id;
                    }
                    const bool ok = valid && id >= 0;
                    const __nv_bfloat16 *src = src0 + (size_t)(ok ? id : 0) * ld;
                    const uint32_t dst = a_s + ((j & 1) ? off_odd : off_even) + (uint32_t)((j >> 1) * 1024);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16u : 0u) : "memory");
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tm_smem_u32(&H.full_bar[s])) : "memory");
            }
            n_base += U.kb1 - U.kb0;
        }
    } else if (warp < 12) {
        // ===== epilogue warps 8..11: TMEM lane quarter (warp & 3) -> registers -> (+bias, +residual, relu) -> global
        const int q = warp & 3, et = tid - 256;
        float *out32 = reinterpret_cast<float *>(out_);
        __nv_bfloat16 *out16 = reinterpret_cast<__nv_bfloat16 *>(out_);
        int li = 0;
        asm volatile("griddepcontrol.wait;" ::: "memory");
        for (int u = blockIdx.x; u < P.n_units; u += gridDim.x, li++) {
            const BfUnit U = bf_unit(P, u);
            const int buf = li & 1;
            if (lane == 0) tm_mbar_wait(tm_smem_u32(&H.acc_full[buf]), (uint32_t)((li >> 1) & 1));
            __syncwarp();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int r = U.row0 + q * 32 + lane;
            const uint32_t taddr0 = tmem + (uint32_t)(buf * P.acc_stride) + ((uint32_t)(q * 32) << 16);
            const bool split = U.pieces > 1;
            float *part = scratch + (size_t)U.split_tile * (kBfM * 128);   // [4-column group][128 rows] float4
            bool finish = !split;
            if (split) {
                for (int cb = 0; cb < P.nc / 16; cb++) {
                    uint32_t v[16];
                    bf_tmem_ld16(taddr0 + (uint32_t)(cb * 16), v);
                    if (r < n_out)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        float *dst = part + ((size_t)(cb * 4 + j) * kBfM + q * 32 + lane) * 4;
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__uint_as_float(v[4 * j])),
                                     "f"(__uint_as_float(v[4 * j + 1])), "f"(__uint_as_float(v[4 * j + 2])), "f"(__uint_as_float(v[4 * j + 3])) : "memory");
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tm_smem_u32(&H.acc_empty[buf])) : "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (et == 0) {
                    __threadfence();
                    const int old = atomicAdd(counters + U.split_tile, 1);
                    const int last = old == U.pieces - 1;
                    if (last) counters[U.split_tile] = 0;
                    __threadfence();
                    H.last_flag = last;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                finish = H.last_flag != 0;
            }
            if (finish && (P.nc & 31) == 0) {
                // coalesced write-out through the swizzled per-warp staging tile, 32 columns per pass (sparse_conv_persist.cu)
                const bool row_ok = r < n_out;
                const uint32_t st = tm_smem_u32(smem + 1024 + q * 4096);
                const int g = lane & 7, sub = lane >> 3;
                for (int ch = 0; ch < P.nc / 32; ch++) {
                    uint32_t v[32];
                    if (!split) {
                        uint32_t lo[16], hi[16];
                        bf_tmem_ld16(taddr0 + (uint32_t)(ch * 32), lo);
                        bf_tmem_ld16(taddr0 + (uint32_t)(ch * 32 + 16), hi);
#pragma unroll
                        for (int j = 0; j < 16; j++) { v[j] = lo[j]; v[16 + j] = hi[j]; }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            float4 *src = reinterpret_cast<float4 *>(part + ((size_t)(ch * 8 + j) * kBfM + q * 32 + lane) * 4);
                            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (row_ok) {
                                t = __ldcg(src);
                                __stcg(src, make_float4(0.f, 0.f, 0.f, 0.f));   // the scratch tile is zero again for its next user
                            }
                            v[4 * j] = __float_as_uint(t.x); v[4 * j + 1] = __float_as_uint(t.y);
                            v[4 * j + 2] = __float_as_uint(t.z); v[4 * j + 3] = __float_as_uint(t.w);
                        }
                    }
                    const int col = U.n0 + ch * 32 + g * 4;
                    float4 rv[8];
                    if (residual) {
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int rr = U.row0 + q * 32 + 4 * i + sub;
                            rv[i] = rr < n_out ? bf_unpack4(__ldg(reinterpret_cast<const uint2 *>(residual + (size_t)rr * ldr + col)))
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                    __syncwarp();                                    // the previous pass has been read out of the staging tile
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st + (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4))),
                                     "r"(v[4 * j]), "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3]) : "memory");
                    __syncwarp();
                    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (bias) bv = __ldg(reinterpret_cast<const float4 *>(bias + col));
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int row = 4 * i + sub, rr = U.row0 + q * 32 + row;
                        float4 o;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                                     : "r"(st + (uint32_t)(row * 128 + ((g ^ (row & 7)) << 4))) : "memory");
                        o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
                        if (residual) { o.x += rv[i].x; o.y += rv[i].y; o.z += rv[i].z; o.w += rv[i].w; }
                        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        if (rr < n_out) {
                            if (out_f32) *reinterpret_cast<float4 *>(out32 + (size_t)rr * ldo + col) = o;
                            else *reinterpret_cast<uint2 *>(out16 + (size_t)rr * ldo + col) = bf_pack4(o);
                        }
                    }
                }
                if (!split) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tm_smem_u32(&H.acc_empty[buf])) : "memory");
                }
            } else if (finish) {
                // channel counts that are not a multiple of 32: direct stores per row
                const bool row_ok = r < n_out;
                for (int cb = 0; cb < P.nc / 16; cb++) {
                    float4 o[4];
                    if (!split) {
                        uint32_t v[16];
                        bf_tmem_ld16(taddr0 + (uint32_t)(cb * 16), v);
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                               __uint_as_float(v[4 * j + 3]));
                    } else if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            float4 *src = reinterpret_cast<float4 *>(part + ((size_t)(cb * 4 + j) * kBfM + q * 32 + lane) * 4);
                            o[j] = __ldcg(src);
                            __stcg(src, make_float4(0.f, 0.f, 0.f, 0.f));
                        }
                    }
                    if (row_ok) {
                        const int col0 = U.n0 + cb * 16;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            if (bias) {
                                const float *bp = bias + col0 + 4 * j;
                                o[j].x += __ldg(bp); o[j].y += __ldg(bp + 1); o[j].z += __ldg(bp + 2); o[j].w += __ldg(bp + 3);
                            }
                            if (residual) {
                                const float4 rv = bf_unpack4(__ldg(reinterpret_cast<const uint2 *>(residual + (size_t)r * ldr + col0 + 4 * j)));
                                o[j].x += rv.x; o[j].y += rv.y; o[j].z += rv.z; o[j].w += rv.w;
                            }
                            if (relu) {
                                o[j].x = fmaxf(o[j].x, 0.f); o[j].y = fmaxf(o[j].y, 0.f); o[j].z = fmaxf(o[j].z, 0.f); o[j].w = fmaxf(o[j].w, 0.f);
                            }
                            if (out_f32) *reinterpret_cast<float4 *>(out32 + (size_t)r * ldo + col0 + 4 * j) = o[j];
                            else *reinterpret_cast<uint2 *>(out16 + (size_t)r * ldo + col0 + 4 * j) = bf_pack4(o[j]);
                        }
                    }
                }
                if (!split) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tm_smem_u32(&H.acc_empty[buf])) : "memory");
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(P.tmem_cols) : "memory");
}

// ---------------------------------------------------------------------------------------------------------- host side
struct BfWorkspace {
    float *scratch = nullptr;
    int *counters = nullptr;
};

static int bf_workspace(cudaStream_t stream, BfWorkspace *ws) {
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, BfWorkspace> table;
    int dev = 0;
    CVB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    auto it = table.find({dev, stream});
    if (it == table.end()) {
        BfWorkspace w;
        const size_t bytes = kBfScratchFloats * sizeof(float) + 4096;
        void *p = nullptr;
        CVB_CUDA(cudaMalloc(&p, bytes));
        CVB_CUDA(cudaMemset(p, 0, bytes));
        w.counters = reinterpret_cast<int *>(p);
        w.scratch = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(p) + 4096);
        it = table.emplace(std::make_pair(dev, stream), w).first;
    }
    *ws = it->second;
    return 0;
}

// same cost model as the TF32 planner (sparse_conv_persist.cu ps_plan), on k-blocks of 64 elements of the flattened axis
static void bf_plan(int64_t n_out, int cin, int cout, int k3, BfPlan *P) {
    const int m_tiles = (int)ceil_div(n_out, kBfM);
    int n_splits = 1;
    while (cout / n_splits > 128 || cout % n_splits != 0 || (cout / n_splits) % 16 != 0) n_splits++;
    P->n_splits = n_splits;
    P->nc = cout / n_splits;
    P->n_tiles = m_tiles * n_splits;
    P->total_kb = (int)ceil_div((int64_t)k3 * cin, kBfKB);
    const int stage_bytes = kBfM * 128 + P->nc * 128;
    const int S = kNumSMs;
    P->n_whole = (P->n_tiles / S) * S;
    const int R = P->n_tiles - P->n_whole;
    int best_ks = 1;
    if (R > 0) {
        double best = 1e30;
        for (int ks = 1; ks <= 32 && ks <= P->total_kb; ks++) {
            const int rounds = (int)ceil_div((int64_t)R * ks, S);
            const int per = (int)ceil_div(P->total_kb, ks);
            const double cost = rounds * (per + 6.0) + (ks > 1 ? 8.0 : 0.0);
            if (cost < best - 1e-9) { best = cost; best_ks = ks; }
        }
    }
    P->ks = best_ks;
    if (best_ks == 1) P->n_whole = P->n_tiles;
    P->n_units = P->n_whole + (P->n_tiles - P->n_whole) * P->ks;
    const int stages = (kBfSmemBytes - 1024 - 16384) / stage_bytes;
    P->stages = stages > kBfMaxStages ? kBfMaxStages : stages;
    int cols = 32;
    while (cols < 2 * P->nc) cols <<= 1;
    P->tmem_cols = cols;
    P->acc_stride = cols / 2;
}

typedef CUresult (*BfEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int bf_weight_map(CUtensorMap *m, const void *d_w, uint64_t ktot, uint64_t cout, uint32_t nc) {
    static BfEncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (BfEncodeTiledFn)p;
    }
    CVB_REQUIRE(fn != nullptr, CVB200_EINVAL, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {ktot, cout}, strides[1] = {ktot * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kBfKB, nc}, estr[2] = {1, 1};
    const CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(d_w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CVB_REQUIRE(rc == CUDA_SUCCESS, CVB200_EINVAL, "cuTensorMapEncodeTiled (bf16 weights %llu x %llu, box 64 x %u) failed: CUresult %d",
                (unsigned long long)ktot, (unsigned long long)cout, nc, (int)rc);
    return 0;
}

}  // namespace cvb200

using namespace cvb200;

extern "C" int cvb200_sc_conv_forward_bf16(const void *d_in, int64_t n_in, int32_t ldi, int32_t cin, const void *d_w, int32_t cout,
                                           const int32_t *d_nbr, int64_t n_out, int32_t k3, const float *d_bias, const void *d_res,
                                           int32_t ldr, int32_t relu, void *d_out, int32_t ldo, int32_t out_f32, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVB_REQUIRE(cin >= 32 && cin % 32 == 0 && cout >= 16 && cout <= 1024 && cout % 16 == 0 && k3 > 0, CVB200_EINVAL,
                "sc_conv_forward_bf16: needs cin %% 32 == 0, cout %% 16 == 0, 16 <= cout <= 1024 (got %d, %d, %d)", cin, cout, k3);
    CVB_REQUIRE(n_out >= 0 && n_out < (1LL << 31) && n_in > 0 && n_in < (1LL << 31) && (int64_t)k3 * cin < (1LL << 30), CVB200_EINVAL,
                "sc_conv_forward_bf16: bad n_out / n_in / K^3 * cin");
    if (n_out == 0) return 0;
    CVB_REQUIRE(d_in && d_w && d_nbr && d_out, CVB200_EINVAL, "sc_conv_forward_bf16: NULL argument");
    CVB_REQUIRE(((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_w) | reinterpret_cast<uintptr_t>(d_out) |
                  reinterpret_cast<uintptr_t>(d_res) | reinterpret_cast<uintptr_t>(d_bias)) & 15) == 0 &&
                    ldi % 8 == 0 && ldo % (out_f32 ? 4 : 8) == 0 && ldr % 8 == 0,
                CVB200_EINVAL, "sc_conv_forward_bf16: 16-byte aligned pointers and row strides required");
    BfPlan P;
    bf_plan(n_out, cin, cout, k3, &P);
    BfWorkspace ws;
    if (int rc = bf_workspace(stream, &ws)) return rc;
    alignas(64) CUtensorMap map_b;
    if (int rc = bf_weight_map(&map_b, d_w, (uint64_t)k3 * cin, (uint64_t)cout, (uint32_t)P.nc)) return rc;
    const size_t smem = 1024 + 16384 + (size_t)P.stages * (kBfM * 128 + P.nc * 128);
    static bool set = false;
    if (!set) {
        CVB_CUDA(cudaFuncSetAttribute(sc_conv_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBfSmemBytes));
        set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(P.n_units < kNumSMs ? P.n_units : kNumSMs));
    cfg.blockDim = dim3(kBfThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CVB_CUDA(cudaLaunchKernelEx(&cfg, sc_conv_bf16_kernel, map_b, (const __nv_bfloat16 *)d_in, (int)ldi, (int)cin, (const int *)d_nbr,
                                (int)n_out, (int)k3, d_bias, (const __nv_bfloat16 *)d_res, (int)ldr, (int)relu, d_out, (int)ldo, (int)out_f32, P,
                                ws.scratch, ws.counters));
    return 0;
}

/* Host-only test hook: the gather warp's index arithmetic (bf_chunk) for k-block `it`, chunk c: out[5] = {k_lo, single, valid, use_hi, ch}. */
extern "C" int cvb200_sc_conv_bf16_chunk(int32_t it, int32_t c, int32_t cin, int32_t k3, int32_t *out) {
    CVB_REQUIRE(out && it >= 0 && c >= 0 && c < 8 && cin >= 32 && cin % 32 == 0 && k3 > 0, CVB200_EINVAL, "sc_conv_bf16_chunk: bad argument");
    const BfChunk r = bf_chunk(it, c, cin, k3 * cin);
    out[0] = r.k_lo; out[1] = r.single; out[2] = r.valid; out[3] = r.use_hi; out[4] = r.ch;
    return 0;
}

/* Host-only: the bf16 kernel's work plan, same layout as cvb200_sc_conv_plan (cblocks slot = 0). */
extern "C" int cvb200_sc_conv_plan_bf16(int64_t n_out, int32_t cin, int32_t cout, int32_t k3, int32_t *h_plan, int32_t *h_units,
                                        int32_t max_units) {
    CVB_REQUIRE(h_plan && n_out > 0 && n_out < (1LL << 31) && cin >= 32 && cin % 32 == 0 && cout >= 16 && cout <= 1024 && cout % 16 == 0 && k3 > 0,
                CVB200_EINVAL, "sc_conv_plan_bf16: needs n_out > 0, cin %% 32 == 0, cout %% 16 == 0, 16 <= cout <= 1024");
    BfPlan P;
    bf_plan(n_out, cin, cout, k3, &P);
    const int v[12] = {P.n_tiles, P.n_splits, P.n_whole, P.ks, P.n_units, P.total_kb, 0, P.stages, P.nc, P.acc_stride, P.tmem_cols,
                       1024 + 16384 + P.stages * (kBfM * 128 + P.nc * 128)};
    for (int i = 0; i < 12; i++) h_plan[i] = v[i];
    for (int u = 0; h_units && u < P.n_units && u < max_units; u++) {
        const BfUnit U = bf_unit(P, u);
        const int w[6] = {U.row0, U.n0, U.kb0, U.kb1, U.pieces, U.split_tile};
        for (int i = 0; i < 6; i++) h_units[6 * u + i] = w[i];
    }
    return 0;
}
