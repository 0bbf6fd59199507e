// canonicalvoting_b200/csrc/sparse_conv_persist.cu -- persistent, warp-specialised tcgen05 sparse convolution.
//
//     out[o, n0:n0+nc) = [relu]( sum_k in[nbr[o,k], :] @ Wt[k][n0:n0+nc, :]^T + bias + residual )
//
// Third generation of the tensor-core forward (sparse_conv_tc.cu: CTA-wide barrier per k-block; sparse_conv_tma.cu:
// warp-specialised, one tile per CTA).  The ncu captures of round 1 (profiles/r1n_*) showed that a third of a CTA's
// life was prologue (neighbour tile staging) and epilogue (TMEM read-back), serial with the MMA loop, and that the
// 1.3-wave grids of the large levels left 30 % of the SM-cycles idle; the small levels paid a memset, float atomics
// and a finishing launch per convolution.  This kernel removes those:
//   * ONE CTA per SM walks a list of work units.  A unit = (row tile of 128 outputs, channel split, piece of the
//     k-block sequence); the host plans whole tiles for the full waves and cuts the tiles of the last, partial wave
//     (or all tiles of a small level) into `ks` pieces so that every SM gets the same amount of k-blocks.
//   * roles (13 warps): warp 0 = weight-tile TMA producer; warps 4 and 12 = MMA issuers, alternating k-blocks under a
//     strict issue order (warp 4 owns the TMEM allocation); warps 1,2,3,5 = gather producers, each filling whole k-blocks
//     alone (16-byte cp.async straight from the feature matrix into the 128B-swizzled A tile, zero-length copy = zero
//     fill for a missing neighbour; neighbour ids are read from the table one k-block ahead and passed by shuffles, no
//     staging pass); warps 8-11 = epilogue; warps 6,7 idle (spare producers).  Issuers and TMA share one warp scheduler
//     (warp % 4 == 0) that hosts no producer.  full/empty mbarriers per ring stage (7-8 stages, ~213 KB of shared memory).
//   * TWO accumulators in tensor memory: the epilogue of unit i (tcgen05.ld -> bias/residual/ReLU -> global) runs
//     (through a swizzled staging tile: coalesced 128-byte row stores) while the MMA warps already accumulate unit i+1.
//   * pieces of a split tile add their partial sums into a zero-initialised, self-cleaning scratch tile
//     (red.global.add.v4.f32, coalesced [4-column group][row] layout); the piece that arrives last (per-tile counter)
//     reads the sum back, re-zeroes the scratch and applies the epilogue -- no memset, no finishing launch.
//   * programmatic dependent launch: barrier init, TMEM allocation, the first weight tiles and neighbour ids of
//     convolution i+1 overlap the tail of convolution i (griddepcontrol.wait only in the gather and epilogue roles).
//   * a 4-channel input (the padded 3-channel 5^3 stem) is gathered 8 neighbours per k-block (g4 mode): no im2col.
#include <cuda.h>

#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"
#include "tcgen05.cuh"

namespace cvb200 {

// The measurement aids (cvb200_sc_set_conv_debug masks, clock64 traces; tools/conv_probe.py, tools/conv_trace.py) exist only in a
// probe build (CVB200_PROBE=1 python -m canonicalvoting_b200.build --force): the shipped kernel carries none of their branches.
#ifdef CVB200_PROBE
#define PS_DBG(P, mask) (((P).dbg & (mask)) != 0)
#define PS_TRACE(trace) ((trace) != nullptr && blockIdx.x == 0)
#else
#define PS_DBG(P, mask) false
#define PS_TRACE(trace) false
#endif

constexpr int kPsM = 128, kPsKB = 32, kPsThreads = 416, kPsMaxStages = 8;
constexpr int kPsProducers = 4;   // gather warps 1,2,3,5 (6,7 too when set to 6: measured slower); TMA warp 0; MMA issuers 4 and 12 (same
                                  // scheduler as warp 0); epilogue warps 8..11
constexpr int kPsSmemBytes = 224 * 1024;                   // of the 227 KiB a CTA may use
constexpr int kPsMaxSplitTiles = 2 * kNumSMs;                   // tiles that can be split in one launch (one partial wave)
constexpr size_t kPsScratchFloats = (size_t)kPsMaxSplitTiles * kPsM * 128;


struct PsPlan {
    int n_tiles;      // row tiles x channel splits
    int n_splits;     // channel splits (tile t -> row tile t / n_splits, channel block t % n_splits)
    int n_whole;      // tiles [0, n_whole) are one unit each
    int ks;           // tiles [n_whole, n_tiles) are cut into ks pieces of the k-block sequence
    int n_units;
    int total_kb;     // k3 * cin / 32
    int cblocks;      // cin / 32
    int stages;
    int nc;           // output channels per tile
    int acc_stride;   // TMEM column offset of the second accumulator
    int tmem_cols;
    int dbg;          // measurement aid (cvb200_sc_set_conv_debug): 1 no gather copies, 2 no zero-fill copies, 4 no MMA, 8 no weight TMA
    int allow_split;  // 0: never cut tiles into pieces
    int force_ks;     // probe override of ks (0 = planner's choice)
};

struct PsHeader {
    unsigned long long full_bar[kPsMaxStages], empty_bar[kPsMaxStages], acc_full[2], acc_empty[2], turn[2];
    unsigned int tmem_base;
    int last_flag;
    int n_out;        // rows of this launch (read from device memory when the launch is size-agnostic)
    PsPlan plan;      // the plan for n_out rows
};

// The row-dependent half of the plan: how the n_out rows are cut into units for `slots` persistent CTAs.  Host and device:
// a size-agnostic launch (row count in device memory, e.g. inside a CUDA graph) plans in the kernel.
__host__ __device__ inline void ps_plan_rows(long long n_out, int slots, PsPlan *P) {
    const int m_tiles = (int)((n_out + kPsM - 1) / kPsM);
    P->n_tiles = m_tiles * P->n_splits;
    const int S = slots;
    P->n_whole = (P->n_tiles / S) * S;
    const int R = P->n_tiles - P->n_whole;
    int best_ks = 1;
    if (R > 0 && P->allow_split) {
        // rounds of the partial wave x (k-blocks per piece + fixed cost of a unit) + cost of the split epilogue, in k-block units
        double best = 1e30;
        for (int ks = 1; ks <= 32 && ks <= P->total_kb; ks++) {
            const int rounds = (int)(((long long)R * ks + S - 1) / S);
            const int per = (P->total_kb + ks - 1) / ks;
            const double cost = rounds * (per + 6.0) + (ks > 1 ? 8.0 : 0.0);
            if (cost < best - 1e-9) { best = cost; best_ks = ks; }
        }
    }
    if (P->force_ks > 0 && R > 0) best_ks = P->force_ks < P->total_kb ? P->force_ks : P->total_kb;   // probe override
    P->ks = best_ks;
    if (best_ks == 1) P->n_whole = P->n_tiles;
    if (P->n_tiles - P->n_whole > 2 * kNumSMs) {   // cannot happen with S <= 2 * kNumSMs and the scratch sized for it; be safe
        P->ks = 1;
        P->n_whole = P->n_tiles;
    }
    P->n_units = P->n_whole + (P->n_tiles - P->n_whole) * P->ks;
}

// fused head decode (cvb200_decode_args) as the kernel sees it; xyz == nullptr: off
struct PsDecode {
    float *xyz, *scale, *prob, *points;
    long long *cls;
    const int4 *coords;
    float res;
    int log_scale;
};

struct PsUnit {
    int row0, n0, kb0, kb1, pieces, split_tile;
};

__host__ __device__ __forceinline__ PsUnit ps_unit(const PsPlan &P, int u) {
    int tile, piece, pieces;
    if (u < P.n_whole) {
        tile = u; piece = 0; pieces = 1;
    } else {
        const int v = u - P.n_whole;
        tile = P.n_whole + v / P.ks; piece = v % P.ks; pieces = P.ks;
    }
    PsUnit U;
    U.row0 = (tile / P.n_splits) * kPsM;
    U.n0 = (tile % P.n_splits) * P.nc;
    U.kb0 = (int)((long long)piece * P.total_kb / pieces);
    U.kb1 = (int)((long long)(piece + 1) * P.total_kb / pieces);
    U.pieces = pieces;
    U.split_tile = tile - P.n_whole;
    return U;
}

__device__ __forceinline__ void ps_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// wait until at most n of this thread's cp.async groups are pending (the PTX operand is an immediate)
__device__ __forceinline__ void ps_cp_async_wait(int n) {
    switch (n) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
        case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
    }
}

// dynamic smem: [header 1 KiB][epilogue staging: 4 warps x 32 rows x 128 B][stages x (A 16 KiB | B nc x 128 B)]
__global__ void __launch_bounds__(kPsThreads, 1)
sc_conv_persist_kernel(const __grid_constant__ CUtensorMap map_b, const float *__restrict__ in, int ldi, int cout_total,
                       const int *__restrict__ nbr, int n_out, int k3, const float *__restrict__ bias,
                       const float *__restrict__ residual, int ldr, int relu, float *__restrict__ out, int ldo, const PsPlan P0,
                       float *__restrict__ scratch, int *__restrict__ counters, long long *__restrict__ trace, int g4,
                       const int *__restrict__ n_out_dev, const PsDecode dec) {
    extern __shared__ __align__(1024) unsigned char smem[];
    PsHeader &H = *reinterpret_cast<PsHeader *>(smem);
    unsigned char *stage0 = smem + 1024 + 16384;       // [header 1 KiB][epilogue staging 4 x 4 KiB][ring]
    const int a_bytes = kPsM * 128, b_bytes = P0.nc * 128, stage_bytes = a_bytes + b_bytes;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool tr = PS_TRACE(trace);   // measurement aid: clock64 stamps of the first 256 k-blocks of CTA 0
    int tn = 0;
    if (tr && tid == 0) trace[3 * 768 + 0] = clock64();

    if (tid == 0) {
        for (int s = 0; s < P0.stages; s++) {
            tm_mbar_init(tm_smem_u32(&H.full_bar[s]), 1 + 32);
            tm_mbar_init(tm_smem_u32(&H.empty_bar[s]), 1);
        }
        for (int b = 0; b < 2; b++) {
            tm_mbar_init(tm_smem_u32(&H.acc_full[b]), PS_DBG(P0, 0x40000) ? 1 : 2);      // both MMA warps
            tm_mbar_init(tm_smem_u32(&H.acc_empty[b]), 4);
            tm_mbar_init(tm_smem_u32(&H.turn[b]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid == 32) {
        // size-agnostic launch: the row count was written by the map builder (complete before the first convolution of the
        // program starts, like the neighbour tables), n_out is its upper bound; the plan is made here
        PsPlan Q = P0;
        int rows = n_out;
        if (n_out_dev) {
            const int dev_rows = __ldg(n_out_dev);
            rows = dev_rows < n_out ? (dev_rows > 0 ? dev_rows : 0) : n_out;
            ps_plan_rows(rows, (int)gridDim.x, &Q);
        }
        H.plan = Q;
        H.n_out = rows;
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tm_smem_u32(&H.tmem_base)), "r"(P0.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = H.tmem_base;
    const PsPlan P = H.plan;
    n_out = H.n_out;
    // everything above overlapped the previous kernel of the stream; its results are visible after this wait
    if (tr && tid == 0) trace[3 * 768 + 1] = clock64();
    // Programmatic dependent launch: everything above overlapped the previous kernel of the stream.  Only what depends
    // on that kernel waits for it (griddepcontrol.wait, per role): the feature gather and the epilogue (residual, output,
    // scratch).  The weight tiles and the neighbour ids are static, so their first loads run ahead of the wait.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        // ===== weight producer: one TMA load of the [nc x 32] block of Wt[k] per k-block.  The whole warp walks the loop
        // (warp-uniform control flow), one elected lane issues.
        int s = 0;
        uint32_t ph = 0;
        for (int u = blockIdx.x; u < P.n_units; u += gridDim.x) {
            const PsUnit U = ps_unit(P, u);
            int k = U.kb0 / P.cblocks, cb = U.kb0 - k * P.cblocks;
            for (int it = U.kb0; it < U.kb1; it++) {
                const long long t0 = tr ? clock64() : 0;
                tm_mbar_wait(tm_smem_u32(&H.empty_bar[s]), ph ^ 1u);   // all lanes poll: measured faster than one polling lane + __syncwarp
                if (tr && lane == 0 && tn < 256) { trace[2 * 768 + 3 * tn] = t0; trace[2 * 768 + 3 * tn + 1] = clock64(); }
                const uint32_t b_s = tm_smem_u32(stage0 + (size_t)s * stage_bytes) + a_bytes;
                const uint32_t full = tm_smem_u32(&H.full_bar[s]);
                if (tm_elect_one()) {
                    if (PS_DBG(P, 8)) {
                        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full) : "memory");
                    } else {
                        tm_expect_tx(full, (uint32_t)b_bytes);
                        tma_load_2d(b_s, &map_b, full, cb * kPsKB, k * cout_total + U.n0);
                    }
                }
                __syncwarp();
                if (tr && lane == 0 && tn < 256) { trace[2 * 768 + 3 * tn + 2] = clock64(); }
                tn++;
                if (++s == P.stages) { s = 0; ph ^= 1u; }
                if (++cb == P.cblocks) { cb = 0; k++; }
            }
        }
    } else if (warp == 4 || warp == 12) {
        // ===== MMA issuers: TWO warps, k-block n of the CTA's running sequence belongs to warp n & 1.  One warp's
        // per-k-block protocol (poll the full barrier, proxy fence, four tcgen05.mma, commit, loop) takes longer than the
        // tensor core needs for the k-block, so a single issuer bounded the kernel ("barriers only": 66 of 90 us); with two,
        // one warp's hand-shake overlaps the other's MMAs.  The MMAs themselves are issued in strict sequence (a `turn`
        // barrier per warp, handed over right after the four MMAs of a k-block): the tensor pipe executes in issue
        // order, stages are released in order, and the first k-block of a unit overwrites the accumulator before anything
        // is added to it.  Both warps commit to acc_full (count 2).
        const int me = warp == 4 ? 0 : 1;
        const int issuers = PS_DBG(P, 0x40000) ? 1 : 2;            // measurement aid: one issuer only (warp 12 idles)
        int li = 0, n_base = 0;
        for (int u = blockIdx.x; u < P.n_units && me < issuers; u += gridDim.x, li++) {
            const PsUnit U = ps_unit(P, u);
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(P.nc >> 3) << 17) | ((uint32_t)(kPsM >> 4) << 24);
            const int buf = li & 1;
            const int first = issuers == 1 ? U.kb0 : U.kb0 + ((me - n_base) & 1);        // my first k-block of this unit
            const uint32_t d_tmem = tmem + (uint32_t)(buf * P.acc_stride);
            if (first == U.kb0 && first < U.kb1) {
                tm_mbar_wait(tm_smem_u32(&H.acc_empty[buf]), (uint32_t)(((li >> 1) & 1) ^ 1));   // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            for (int it = first; it < U.kb1; it += issuers) {
                const int n = n_base + it - U.kb0;
                const int round = n / P.stages, s = n - round * P.stages;
                const long long t0 = tr ? clock64() : 0;
                tm_mbar_wait(tm_smem_u32(&H.full_bar[s]), (uint32_t)(round & 1));
                if (tr && me == 0 && lane == 0 && tn < 256) { trace[3 * tn] = t0; trace[3 * tn + 1] = clock64(); }
                const uint32_t a_s = tm_smem_u32(stage0 + (size_t)s * stage_bytes), b_s = a_s + a_bytes;
                const uint64_t a_desc = tm_desc_k_sw128(a_s), b_desc = tm_desc_k_sw128(b_s);
                // cp.async wrote through the generic proxy, the tensor core reads through the async proxy.  (Fencing on the producer side
                // instead -- wait for the copy group, fence, one arrival per warp -- was measured in round 2, with one and with two
                // issuers: 1.31-1.39 ms per scene against 1.23 ms; the asynchronous arrivals below never stall a producer.)
                if (!PS_DBG(P, 16) && lane == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                // my turn: the other warp has issued k-block n - 1
                if (n > 0 && issuers == 2) tm_mbar_wait(tm_smem_u32(&H.turn[me]), (uint32_t)(((n >> 1) + me + 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tm_elect_one()) {
                    if (!PS_DBG(P, 4)) {
#pragma unroll
                        for (int kk = 0; kk < kPsKB / 8; kk++)
                            tm_umma_tf32(d_tmem, a_desc + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2), idesc, (it > U.kb0 || kk > 0) ? 1u : 0u);
                    }
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tm_smem_u32(&H.turn[me ^ 1])) : "memory");
                    tm_commit(tm_smem_u32(&H.empty_bar[s]));
                }
                __syncwarp();
                if (tr && me == 0 && lane == 0 && tn < 256) { trace[3 * tn + 2] = clock64(); }
                tn++;
            }
            if (tm_elect_one()) {
                if (first < U.kb1) tm_commit(tm_smem_u32(&H.acc_full[buf]));       // arrives when my MMAs of the unit are complete
                else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tm_smem_u32(&H.acc_full[buf])) : "memory");
            }
            __syncwarp();
            n_base += U.kb1 - U.kb0;
        }
    } else if (warp < 8) {
        // ===== gather producers.  A single warp's instruction stream (barrier poll, address arithmetic, copies, arrival)
        // costs ~1000 cycles per k-block whatever the copy count (tools/conv_trace.py), so the gather warps do NOT share
        // a k-block: warp w owns the k-blocks n % W == w (W = min(4, stages)) of the CTA's running k-block sequence and fills their stages
        // alone -- W k-blocks are being filled at any time.  (n, n + W are less than a ring apart, so a warp is never
        // more than one phase ahead of an empty barrier, which is all a parity wait can tell apart.)  Lane (rb, c) copies
        // the 16-byte chunk c of rows rb + 4 j, j < 32; completion arrives on the stage's full barrier by itself
        // (cp.async.mbarrier.arrive.noinc, 32 arrivals per stage), nobody waits for data.
        // warp placement: a scheduler (warp % 4) that hosts an MMA issuer or the TMA warp hosts no gather warp, whose long
        // instruction streams would delay the issuers' latency-critical hand-shakes
        const int w = warp < 4 ? warp - 1 : warp - 2, c = lane & 7, rb = lane >> 3;   // warps 1,2,3,5,6,7 -> 0..5
        const int W = P.stages < kPsProducers ? P.stages : kPsProducers;   // producing warps: own k-blocks must be less than a ring apart
        const uint32_t off_even = (uint32_t)(rb * 128 + ((c ^ rb) << 4)), off_odd = (uint32_t)((rb + 4) * 128 + ((c ^ (rb + 4)) << 4));
        const size_t ld4 = (size_t)ldi;
        int n_base = 0;
        bool dep_done = false;
        for (int u = blockIdx.x; u < P.n_units && w < W; u += gridDim.x) {
            const PsUnit U = ps_unit(P, u);
            // neighbour ids: lane l holds the ids of rows l, l + 32, l + 64, l + 96 for one kernel offset (4 coalesced-ish
            // loads, issued one k-block ahead); the id of row rb + 4 j reaches lane (rb, c) by a shuffle
            const int *nlane = nbr + (size_t)(U.row0 + lane) * k3;
            const int lane_rows = n_out - U.row0 - lane;          // row lane + 32 m exists iff 32 m < lane_rows
            int cur[4], nxt[4];
            int it = U.kb0 + (((w - n_base) % W) + W) % W;
            int k_cur = -1, k_nxt = -1;
            if (it < U.kb1 && !g4) {
                k_nxt = it / P.cblocks;
#pragma unroll
                for (int m = 0; m < 4; m++) nxt[m] = 32 * m < lane_rows ? __ldg(nlane + (size_t)(32 * m) * k3 + k_nxt) : -1;
            }
            for (; it < U.kb1; it += W) {
                const int k = it / P.cblocks, cb = it - k * P.cblocks;
                if (k != k_cur) {       // k == k_nxt by construction
#pragma unroll
                    for (int m = 0; m < 4; m++) cur[m] = nxt[m];
                    k_cur = k;
                }
                if (it + W < U.kb1 && !g4) {
                    const int kn = (it + W) / P.cblocks;
                    if (kn != k_cur) {
#pragma unroll
                        for (int m = 0; m < 4; m++) nxt[m] = 32 * m < lane_rows ? __ldg(nlane + (size_t)(32 * m) * k3 + kn) : -1;
                        k_nxt = kn;
                    }
                }
                const int n = n_base + it - U.kb0;
                const int round = n / P.stages, s = n - round * P.stages;
                if (!dep_done) {
                    asm volatile("griddepcontrol.wait;" ::: "memory");
                    dep_done = true;
                    if (tr && tid == 32) trace[3 * 768 + 2] = clock64();
                }
                const long long t0 = tr ? clock64() : 0;
                if (lane == 0) tm_mbar_spin(tm_smem_u32(&H.empty_bar[s]), (uint32_t)((round & 1) ^ 1));
                __syncwarp();
                if (tr && tid == 32 && tn < 256) { trace[768 + 3 * tn] = t0; trace[768 + 3 * tn + 1] = clock64(); }
                const uint32_t a_s = tm_smem_u32(stage0 + (size_t)s * stage_bytes);
                if (g4) {
                    // 4-channel input (the padded 3-channel stem): the k-block's 32 "channels" are 8 neighbours x 4 channels, chunk
                    // c of row r is the 16-byte feature vector of neighbour 8 cb + c -- the im2col matrix is never materialised
                    const int col = 8 * cb + c;
                    int ids[32];
#pragma unroll
                    for (int j = 0; j < 32; j++)
                        ids[j] = (col < g4 && U.row0 + rb + 4 * j < n_out) ? __ldg(nbr + (size_t)(U.row0 + rb + 4 * j) * g4 + col) : -1;
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const float *src = in + (size_t)(ids[j] >= 0 ? ids[j] : 0) * ld4;
                        const uint32_t dst = a_s + ((j & 1) ? off_odd : off_even) + (uint32_t)((j >> 1) * 1024);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ids[j] >= 0 ? 16u : 0u) : "memory");
                    }
                } else {
                const float *src0 = in + c * 4 + cb * kPsKB;
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const int id = __shfl_sync(0xffffffffu, cur[j >> 3], rb + 4 * (j & 7));   // row rb + 4 j = 32 (j >> 3) + rb + 4 (j & 7)
                    const float *src = src0 + (size_t)(id >= 0 ? id : 0) * ld4;
                    const uint32_t dst = a_s + ((j & 1) ? off_odd : off_even) + (uint32_t)((j >> 1) * 1024);
                    if (!PS_DBG(P, 1))
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(id >= 0 ? 16u : 0u) : "memory");
                }
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tm_smem_u32(&H.full_bar[s])) : "memory");
                if (tr && tid == 32 && tn < 256) { trace[768 + 3 * tn + 2] = clock64(); }
                tn++;
            }
            (void)k_nxt;
            n_base += U.kb1 - U.kb0;
        }
    } else {
        // ===== epilogue warps 8..11: TMEM lane quarter (warp & 3) -> registers -> (+bias, +residual, relu) -> global
        const int q = warp & 3, et = tid - 256;
        int li = 0;
        asm volatile("griddepcontrol.wait;" ::: "memory");
        for (int u = blockIdx.x; u < P.n_units; u += gridDim.x, li++) {
            const PsUnit U = ps_unit(P, u);
            const int buf = li & 1;
            // one polling lane per warp: 128 threads parked in try_wait on the header slowed every other mbarrier operation
            // of the CTA down (tools/conv_trace.py)
            if (lane == 0 || PS_DBG(P, 128)) tm_mbar_wait(tm_smem_u32(&H.acc_full[buf]), (uint32_t)((li >> 1) & 1));
            __syncwarp();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tr && et == 0 && li == 0) trace[3 * 768 + 3] = clock64();
            const int r = U.row0 + q * 32 + lane;
            const uint32_t taddr0 = tmem + (uint32_t)(buf * P.acc_stride) + ((uint32_t)(q * 32) << 16);
            const bool split = U.pieces > 1;
            float *part = scratch + (size_t)U.split_tile * (kPsM * 128);   // [4-column group][128 rows] float4
            bool finish = !split;
            if (split) {
                for (int cb = 0; cb < P.nc / 16; cb++) {
                    uint32_t v[16];
                    ps_tmem_ld16(taddr0 + (uint32_t)(cb * 16), v);
                    if (r < n_out)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        float *dst = part + ((size_t)(cb * 4 + j) * kPsM + q * 32 + lane) * 4;
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(__uint_as_float(v[4 * j])),
                                     "f"(__uint_as_float(v[4 * j + 1])), "f"(__uint_as_float(v[4 * j + 2])), "f"(__uint_as_float(v[4 * j + 3])) : "memory");
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tm_smem_u32(&H.acc_empty[buf])) : "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (tr && et == 0 && li == 0) trace[3 * 768 + 4] = clock64();
                if (et == 0) {
                    __threadfence();
                    const int old = atomicAdd(counters + U.split_tile, 1);
                    const int last = old == U.pieces - 1;
                    if (last) counters[U.split_tile] = 0;
                    __threadfence();
                    H.last_flag = last;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (tr && et == 0 && li == 0) trace[3 * 768 + 5] = clock64();
                finish = H.last_flag != 0;
            }
            if (dec.xyz != nullptr) {
                // ===== fused head decode of the joint model (eval_joint.py:173-193, 9 classes: 64 columns = xyz[9][3] | scale[9][3] |
                // logits[10]): the row never leaves the SM.  Same operations in the same order as head_decode_kernel
                // (sparse_engine.cu), on the accumulator columns + bias.  The launcher guarantees whole tiles (no pieces), nc == 64.
                uint32_t t3[16];
                ps_tmem_ld16(taddr0 + 48u, t3);                           // columns 48..63: scale of classes 7, 8 and the 10 logits
                float lg[10];
#pragma unroll
                for (int c = 0; c < 10; c++) lg[c] = __uint_as_float(t3[6 + c]) + (bias ? __ldg(bias + 54 + c) : 0.f);
                float best = -INFINITY, best_obj = -INFINITY;
                int k = 0, k_obj = 0;
#pragma unroll
                for (int c = 0; c < 10; c++) {                            // first maximum, like torch.argmax
                    if (lg[c] > best) { best = lg[c]; k = c; }
                    if (c < 9 && lg[c] > best_obj) { best_obj = lg[c]; k_obj = c; }
                }
                float sum = 0.f;
#pragma unroll
                for (int c = 0; c < 10; c++) sum += expf(lg[c] - best);
                if (k == 9) k = 0;                                        // class_label_idx[class_label_idx == nclasses] = 0  (:178)
                float px[3] = {0.f, 0.f, 0.f}, ps[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int cb = 0; cb < 4; cb++) {
                    uint32_t t[16];
                    if (cb < 3) ps_tmem_ld16(taddr0 + (uint32_t)(cb * 16), t);
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const int col = cb * 16 + j;
                        if (col >= 54) continue;
                        const float v = __uint_as_float(cb < 3 ? t[j] : t3[j]) + (bias ? __ldg(bias + col) : 0.f);
#pragma unroll
                        for (int e = 0; e < 3; e++) {
                            if (col == 3 * k + e) px[e] = v;
                            if (col == 27 + 3 * k + e) ps[e] = v;
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tm_smem_u32(&H.acc_empty[buf])) : "memory");
                if (r < n_out) {
#pragma unroll
                    for (int e = 0; e < 3; e++) {
                        dec.xyz[3 * (size_t)r + e] = px[e];
                        dec.scale[3 * (size_t)r + e] = dec.log_scale ? expf(ps[e]) : ps[e];
                    }
                    dec.cls[r] = k_obj;                                   // argmax over the object classes          (:188)
                    dec.prob[r] = expf(best_obj - best) / sum;            // max softmax over the object classes     (:189)
                    if (dec.coords) {
                        const int4 c = __ldg(dec.coords + r);
                        dec.points[3 * (size_t)r] = __fmul_rn((float)c.y, dec.res);
                        dec.points[3 * (size_t)r + 1] = __fmul_rn((float)c.z, dec.res);
                        dec.points[3 * (size_t)r + 2] = __fmul_rn((float)c.w, dec.res);
                    }
                }
            } else if (finish && (P.nc & 31) == 0) {
                // Coalesced write-out: 32-column chunks go through a swizzled per-warp staging tile, so that a warp
                // instruction stores (and reads the residual of) 4 rows x 128 contiguous bytes instead of 32 rows x 16 bytes.
                // (The uncoalesced version kept so many sectors in flight that the MMA warp's proxy fence -- a MEMBAR --
                // stalled ~7500 cycles per chunk and with it the whole ring: 159 -> 90 us on the 96-channel 3^3 layers.)
                const bool row_ok = r < n_out;
                const uint32_t st = tm_smem_u32(smem + 1024 + q * 4096);
                const int g = lane & 7, sub = lane >> 3;
                for (int ch = 0; ch < P.nc / 32; ch++) {
                    uint32_t v[32];
                    if (!split) {
                        if (PS_DBG(P, 0x20000)) {
#pragma unroll
                            for (int j = 0; j < 32; j++) v[j] = 0u;
                        } else {
                            uint32_t lo[16], hi[16];
                            ps_tmem_ld16(taddr0 + (uint32_t)(ch * 32), lo);
                            ps_tmem_ld16(taddr0 + (uint32_t)(ch * 32 + 16), hi);
#pragma unroll
                            for (int j = 0; j < 16; j++) { v[j] = lo[j]; v[16 + j] = hi[j]; }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            float4 *src = reinterpret_cast<float4 *>(part + ((size_t)(ch * 8 + j) * kPsM + q * 32 + lane) * 4);
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
                    if (residual && !PS_DBG(P, 0x10000)) {
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int rr = U.row0 + q * 32 + 4 * i + sub;
                            rv[i] = rr < n_out ? __ldg(reinterpret_cast<const float4 *>(residual + (size_t)rr * ldr + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                    __syncwarp();                                    // the previous chunk has been read out of the staging tile
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
                        if (residual && !PS_DBG(P, 0x10000)) { o.x += rv[i].x; o.y += rv[i].y; o.z += rv[i].z; o.w += rv[i].w; }
                        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        if (rr < n_out && !PS_DBG(P, 0x10000)) *reinterpret_cast<float4 *>(out + (size_t)rr * ldo + col) = o;
                    }
                }
                if (!split) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tm_smem_u32(&H.acc_empty[buf])) : "memory");
                }
            } else if (finish) {
                // channel counts that are not a multiple of 32: direct 16-byte stores per row
                const bool row_ok = r < n_out;
                for (int cb = 0; cb < P.nc / 16; cb++) {
                    float4 o[4];
                    if (!split) {
                        uint32_t v[16];
                        ps_tmem_ld16(taddr0 + (uint32_t)(cb * 16), v);
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                               __uint_as_float(v[4 * j + 3]));
                    } else if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            float4 *src = reinterpret_cast<float4 *>(part + ((size_t)(cb * 4 + j) * kPsM + q * 32 + lane) * 4);
                            o[j] = __ldcg(src);
                            __stcg(src, make_float4(0.f, 0.f, 0.f, 0.f));
                        }
                    }
                    if (row_ok) {
                        float *dst = out + (size_t)r * ldo + U.n0 + cb * 16;
                        const float *res = residual ? residual + (size_t)r * ldr + U.n0 + cb * 16 : nullptr;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            if (bias) {
                                const float *bp = bias + U.n0 + cb * 16 + 4 * j;
                                o[j].x += __ldg(bp); o[j].y += __ldg(bp + 1); o[j].z += __ldg(bp + 2); o[j].w += __ldg(bp + 3);
                            }
                            if (res) {
                                const float4 rv = __ldg(reinterpret_cast<const float4 *>(res) + j);
                                o[j].x += rv.x; o[j].y += rv.y; o[j].z += rv.z; o[j].w += rv.w;
                            }
                            if (relu) {
                                o[j].x = fmaxf(o[j].x, 0.f); o[j].y = fmaxf(o[j].y, 0.f); o[j].z = fmaxf(o[j].z, 0.f); o[j].w = fmaxf(o[j].w, 0.f);
                            }
                            reinterpret_cast<float4 *>(dst)[j] = o[j];
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
    if (tr && tid == 256) trace[3 * 768 + 6] = clock64();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tr && tid == 0) trace[3 * 768 + 7] = clock64();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(P.tmem_cols) : "memory");
}

// ---------------------------------------------------------------------------------------------------------- host side
int g_ps_allow_split = 1;   // 0: never cut tiles into pieces (bit-reproducible summation order; used by the tests)
int g_ps_use_pdl = 1;
int g_ps_debug = 0;
long long *g_ps_trace = nullptr;

// per-(device, stream) scratch: zero-initialised partial-sum tiles + arrival counters, both self-cleaning
struct PsWorkspace {
    float *scratch = nullptr;
    int *counters = nullptr;
};

static int ps_workspace(cudaStream_t stream, PsWorkspace *ws) {
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, PsWorkspace> table;
    int dev = 0;
    CVB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    auto it = table.find({dev, stream});
    if (it == table.end()) {
        PsWorkspace w;
        const size_t bytes = kPsScratchFloats * sizeof(float) + 4096;
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


static void ps_plan(int64_t n_out, int cin, int cout, int k3, PsPlan *P, int *ctas_per_sm) {
    int n_splits = 1;
    while (cout / n_splits > 128 || cout % n_splits != 0 || (cout / n_splits) % 16 != 0) n_splits++;
    P->n_splits = n_splits;
    P->nc = cout / n_splits;
    P->cblocks = cin / kPsKB;
    P->total_kb = k3 * P->cblocks;
    const int stage_bytes = kPsM * 128 + P->nc * 128;
    // two CTAs per SM with 3-stage rings were measured slower than one deep ring on every level (113 vs 91 us on the large
    // layers) and are no longer offered; the register budget is now spent on one CTA
    *ctas_per_sm = 1;
    P->allow_split = g_ps_allow_split;
    P->force_ks = (g_ps_debug >> 8) & 255;
    ps_plan_rows(n_out, kNumSMs * *ctas_per_sm, P);
    int stages = (kPsSmemBytes - 1024 - 16384) / stage_bytes;
    P->stages = stages > kPsMaxStages ? kPsMaxStages : stages;
    int cols = 32;
    while (cols < 2 * P->nc) cols <<= 1;
    P->tmem_cols = cols;
    P->acc_stride = cols / 2;
    P->dbg = g_ps_debug & 0xff00ff;
}

// g4 > 0: `d_in` has 4 channels per row (ldi = 4), `d_nbr` is a table [n_out][g4] and the contraction runs over
// K = 32 * ceil(g4 / 8) = (neighbour, channel) pairs; cin must be that K, k3 must be 1, d_wt = [cout][K]
int launch_conv_persist(const float *d_in, int64_t n_in, int ldi, int cin, const float *d_wt, int cout, const int32_t *d_nbr,
                        int64_t n_out, int k3, const float *d_bias, const float *d_res, int ldr, int relu, float *d_out, int ldo,
                        cudaStream_t stream, int g4, const int32_t *d_n_out, const cvb200_decode_args *decode) {
    CVB_REQUIRE(g4 == 0 || (k3 == 1 && ldi == 4 && cin == 32 * ((g4 + 7) / 8)), CVB200_EINVAL,
                "sc_conv (4-channel gather): needs k3 == 1, ldi == 4, cin == 32 * ceil(width / 8) (got %d, %d, %d for width %d)", k3, ldi, cin, g4);
    CVB_REQUIRE(cin > 0 && cin % kPsKB == 0 && cout >= 16 && cout <= 1024 && cout % 16 == 0 && k3 > 0, CVB200_EINVAL,
                "sc_conv_forward_tc: needs cin %% 32 == 0, cout %% 16 == 0, 16 <= cout (got %d, %d, %d)", cin, cout, k3);
    CVB_REQUIRE(n_out >= 0 && n_out < (1LL << 31) && n_in > 0 && n_in < (1LL << 31), CVB200_EINVAL, "sc_conv_forward_tc: bad n_out / n_in");
    if (n_out == 0) return 0;
    CVB_REQUIRE(d_in && d_wt && d_nbr && (d_out || decode), CVB200_EINVAL, "sc_conv_forward_tc: NULL argument");
    CVB_REQUIRE(((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_wt) | reinterpret_cast<uintptr_t>(d_out) |
                  reinterpret_cast<uintptr_t>(d_res)) & 15) == 0 && ldi % 4 == 0 && ldo % 4 == 0 && ldr % 4 == 0,
                CVB200_EINVAL, "sc_conv_forward_tc: 16-byte aligned pointers and row strides required");
    PsDecode dec = {};
    if (decode) {
        CVB_REQUIRE(decode->nclasses == 9 && cout == 64 && !d_res && !relu, CVB200_EINVAL,
                    "sc_conv (fused decode): needs the joint head (9 classes, 64 output channels), no residual, no ReLU (got %d classes, %d channels)",
                    decode->nclasses, cout);
        CVB_REQUIRE(decode->xyz && decode->scale && decode->class_pred && decode->prob && (decode->coords == nullptr) == (decode->points == nullptr),
                    CVB200_EINVAL, "sc_conv (fused decode): NULL output (coords and points go together)");
        dec.xyz = decode->xyz; dec.scale = decode->scale; dec.prob = decode->prob; dec.points = decode->points;
        dec.cls = (long long *)decode->class_pred; dec.coords = (const int4 *)decode->coords; dec.res = decode->res;
        dec.log_scale = decode->log_scale;
    }
    PsPlan P;
    int ctas_per_sm = 1;
    const int keep_split = g_ps_allow_split;
    if (decode) g_ps_allow_split = 0;        // the decode epilogue works on whole tiles (a 1x1x1 convolution has 3 k-blocks: nothing to cut)
    ps_plan(n_out, cin, cout, k3, &P, &ctas_per_sm);
    g_ps_allow_split = keep_split;
    PsWorkspace ws;
    if (int rc = ps_workspace(stream, &ws)) return rc;
    alignas(64) CUtensorMap map_b;
    if (int rc = make_map_2d(&map_b, d_wt, (uint64_t)cin, (uint64_t)k3 * cout, (uint64_t)cin * 4, kPsKB, (uint32_t)P.nc)) return rc;
    const size_t smem = 1024 + 16384 + (size_t)P.stages * (kPsM * 128 + P.nc * 128);
    static DeviceOnce once;
    if (!once.done()) {
        CVB_CUDA(cudaFuncSetAttribute(sc_conv_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPsSmemBytes));
        once.mark();
    }
    cudaLaunchConfig_t cfg = {};
    const int slots = kNumSMs * ctas_per_sm;
    // size-agnostic launch (d_n_out: row count in device memory, n_out its upper bound): always the full grid, the kernel plans
    cfg.gridDim = dim3((unsigned)(d_n_out ? slots : (P.n_units < slots ? P.n_units : slots)));
    cfg.blockDim = dim3(kPsThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_ps_use_pdl ? 1 : 0;
    CVB_CUDA(cudaLaunchKernelEx(&cfg, sc_conv_persist_kernel, map_b, d_in, ldi, cout, (const int *)d_nbr, (int)n_out, k3, d_bias, d_res,
                                ldr, relu, d_out, ldo, P, ws.scratch, ws.counters, g_ps_trace, g4, (const int *)d_n_out, dec));
    return 0;
}

}  // namespace cvb200

extern "C" int cvb200_sc_conv_forward_tc(const float *d_in, int64_t n_in, int32_t cin, const float *d_wt, int32_t cout,
                                         const int32_t *d_nbr, int64_t n_out, int32_t k3, const float *d_bias, float *d_out,
                                         void *stream_) {
    return cvb200::launch_conv_persist(d_in, n_in, cin, cin, d_wt, cout, d_nbr, n_out, k3, d_bias, nullptr, 0, 0, d_out, cout,
                                       (cudaStream_t)stream_, 0, nullptr, nullptr);
}

extern "C" int cvb200_sc_set_conv_options(int32_t allow_split, int32_t use_pdl) {
    cvb200::g_ps_allow_split = allow_split != 0;
    cvb200::g_ps_use_pdl = use_pdl != 0;
    return 0;
}


/* Host-only inspection of the work plan of one convolution (no device work): what the persistent kernel's CTAs will walk.
 * h_plan[12] = {n_tiles, n_splits, n_whole, ks, n_units, total_kb, cblocks, stages, nc, acc_stride, tmem_cols, smem_bytes};
 * h_units (may be NULL) receives n_units rows {row0, n0, kb0, kb1, pieces, split_tile}, at most max_units of them. */
extern "C" int cvb200_sc_conv_plan(int64_t n_out, int32_t cin, int32_t cout, int32_t k3, int32_t *h_plan, int32_t *h_units,
                                   int32_t max_units) {
    using namespace cvb200;
    CVB_REQUIRE(h_plan && n_out > 0 && n_out < (1LL << 31) && cin > 0 && cin % kPsKB == 0 && cout >= 16 && cout <= 1024 && cout % 16 == 0 && k3 > 0,
                CVB200_EINVAL, "sc_conv_plan: needs n_out > 0, cin %% 32 == 0, cout %% 16 == 0, 16 <= cout <= 1024 (got %lld, %d, %d, %d)",
                (long long)n_out, cin, cout, k3);
    PsPlan P;
    int ctas_per_sm = 1;
    ps_plan(n_out, cin, cout, k3, &P, &ctas_per_sm);
    const int v[12] = {P.n_tiles, P.n_splits, P.n_whole, P.ks, P.n_units, P.total_kb, P.cblocks, P.stages, P.nc, P.acc_stride, P.tmem_cols,
                       1024 + 16384 + P.stages * (kPsM * 128 + P.nc * 128)};
    for (int i = 0; i < 12; i++) h_plan[i] = v[i];
    for (int u = 0; h_units && u < P.n_units && u < max_units; u++) {
        const PsUnit U = ps_unit(P, u);
        const int w[6] = {U.row0, U.n0, U.kb0, U.kb1, U.pieces, U.split_tile};
        for (int i = 0; i < 6; i++) h_units[6 * u + i] = w[i];
    }
    return 0;
}

/* Measurement aid: switch off parts of the persistent kernel (results become garbage): 1 no gather copies, 2 no zero-fill
 * copies, 4 no MMA, 8 no weight TMA.  0 = normal operation. */
extern "C" int cvb200_sc_set_conv_debug(int32_t mask) {
    cvb200::g_ps_debug = mask;
    return 0;
}

/* Measurement aid: device buffer of 3 x 768 int64 receiving clock64 stamps (before wait, after wait, after issue) of the
 * first 256 k-blocks of CTA 0 for the MMA thread, gather warp 2 and the weight-TMA thread.  NULL switches it off. */
extern "C" int cvb200_sc_set_conv_trace(void *d_trace) {
    cvb200::g_ps_trace = (long long *)d_trace;
    return 0;
}
