// canonicalvoting_b200/csrc/hv_vote.cu -- the Hough-voting op on sm_100a.
//
// Replaces the reference's three kernels + host wrapper
//   hv_cuda_forward_kernel / hv_cuda_average_kernel / hv_cuda_forward
//   (houghvoting/src/hv_cuda_kernel.cu:12-97, :100-119, :121-165)
//   hv_cuda_backward_kernel / hv_cuda_backward (:168-261, :265-302)
// with a different decomposition (see DESIGN.md "vote op" and the "forward" section below):
// one work item per (point, theta), the six channels of a voxel accumulated in one 32-byte workspace
// sector by two vector reductions per corner (16 per vote instead of 48 scalar atomics), and a
// write-out pass that normalises, writes every output once and leaves the workspace zero again.
//
// Float contract: the integer voxel index of a vote must be bit-identical to the
// reference's sm_100 build.  vote_center() spells out that build's exact operation
// order and FMA fusion with non-contractable intrinsics (see oracle/hv_oracle.c header
// for the derivation from the reference SASS).  Never compile this file with
// --use_fast_math: cosf/sinf/div must be the accurate versions the reference uses.
#include "common.cuh"

namespace cvb200 {

constexpr float kTwoPiF = 6.2831854820251464844f;  // 2 * 3.141592654f  (hv_cuda_kernel.cu:35)

struct HvGeom {
    float cx, cy, cz;  // grid corner = min(points)            (hv_cuda_kernel.cu:130,151)
    float res;
    int X, Y, Z;
};

// theta_i = i * rot_interval; accurate cosf/sinf as in the reference (:35-38).
__device__ __forceinline__ void theta_cs(int i, int num_rots, float &cs, float &sn) {
    const float rot_interval = __fdiv_rn(kTwoPiF, (float)num_rots);
    const float theta = __fmul_rn((float)i, rot_interval);
    cs = cosf(theta);
    sn = sinf(theta);
}

__device__ __forceinline__ void fill_theta_table(float *s_cos, float *s_sin, int num_rots) {
    for (int i = threadIdx.x; i < num_rots; i += blockDim.x) theta_cs(i, num_rots, s_cos[i], s_sin[i]);
}

// theta-independent part of a point's vote: corr.x, corr.z and the grid y coordinate.
__device__ __forceinline__ void point_prep(float py, float xx, float xy, float xz, float sx, float sy, float sz,
                                           const HvGeom &g, float &corr_x, float &corr_z, float &gy) {
    corr_x = __fmul_rn(xx, sx);
    corr_z = __fmul_rn(xz, sz);
    gy = __fdiv_rn(__fadd_rn(__fmaf_rn(xy, -sy, py), -g.cy), g.res);
}

// theta-dependent part: grid x/z coordinates (:38-40 as compiled for sm_100).
__device__ __forceinline__ void vote_center(float px, float pz, float corr_x, float corr_z, float cs, float sn,
                                            const HvGeom &g, float &gx, float &gz) {
    const float off_x = __fmaf_rn(corr_z, sn, -__fmul_rn(corr_x, cs));
    const float off_z = __fmaf_rn(corr_x, -sn, -__fmul_rn(corr_z, cs));
    gx = __fdiv_rn(__fadd_rn(__fadd_rn(px, off_x), -g.cx), g.res);
    gz = __fdiv_rn(__fadd_rn(__fadd_rn(pz, off_z), -g.cz), g.res);
}

// bounds test (:41-44).  NaN coordinates are dropped (the reference's behaviour is
// undefined there: it would index with int(NaN)).
__device__ __forceinline__ bool vote_in_bounds(float gx, float gy, float gz, const HvGeom &g) {
    return gx >= 0.f && gy >= 0.f && gz >= 0.f && gx < (float)(g.X - 1) && gy < (float)(g.Y - 1) &&
           gz < (float)(g.Z - 1);
}

// ------------------------------------------------------------------ forward ------
// Workspace sector of voxel v: work[8v + {0:obj, 1:rot_cos, 2:rot_sin, 3:scale0, 4:scale1, 5:scale2, 6,7: unused}]
// (A counting-sort + per-voxel gather formulation without float atomics, and a variant of this one with
// per-block "touched" flags that lets the write-out skip untouched workspace, were built and measured:
// profiles/exp_*; both are slower -- see DESIGN.md "what was tried".)
constexpr int kScatterThreads = 256;
constexpr int kPtsPerBlock = 64;   // power of two >= 32: a warp = 32 consecutive points, one theta

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d));
}
__device__ __forceinline__ void red_add_v2(float *addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b));
}

__global__ void __launch_bounds__(kScatterThreads)
hv_scatter_kernel(const float *__restrict__ points, const float *__restrict__ xyz, const float *__restrict__ scale,
                  const float *__restrict__ obj, int64_t n, int num_rots, HvGeom g, float *__restrict__ work) {
    extern __shared__ float s_theta[];  // [2 * num_rots]
    __shared__ float s_p[3 * kPtsPerBlock];   // points            -> (px, py, pz)
    __shared__ float s_x[3 * kPtsPerBlock];   // xyz               -> (corr_x, gy, corr_z)
    __shared__ float s_s[3 * kPtsPerBlock];   // scale
    __shared__ float s_o[kPtsPerBlock];       // objectness

    float *s_cos = s_theta, *s_sin = s_theta + num_rots;
    fill_theta_table(s_cos, s_sin, num_rots);

    const int64_t p0 = (int64_t)blockIdx.x * kPtsPerBlock;
    const int npts = (int)min((int64_t)kPtsPerBlock, n - p0);
    // coalesced flat staging of this CTA's AoS rows
    for (int k = threadIdx.x; k < 3 * npts; k += kScatterThreads) {
        s_p[k] = __ldg(points + 3 * p0 + k);
        s_x[k] = __ldg(xyz + 3 * p0 + k);
        s_s[k] = __ldg(scale + 3 * p0 + k);
    }
    for (int k = threadIdx.x; k < npts; k += kScatterThreads) s_o[k] = __ldg(obj + p0 + k);
    __syncthreads();
    if (threadIdx.x < npts) {
        const int t = threadIdx.x;
        float corr_x, corr_z, gy;
        point_prep(s_p[3 * t + 1], s_x[3 * t], s_x[3 * t + 1], s_x[3 * t + 2], s_s[3 * t], s_s[3 * t + 1],
                   s_s[3 * t + 2], g, corr_x, corr_z, gy);
        s_x[3 * t] = corr_x;
        s_x[3 * t + 1] = gy;
        s_x[3 * t + 2] = corr_z;
    }
    __syncthreads();

    const int items = kPtsPerBlock * num_rots;
    const int64_t YZ = (int64_t)g.Y * g.Z;
    for (int j = threadIdx.x; j < items; j += kScatterThreads) {
        const int t = j & (kPtsPerBlock - 1);
        const int i = j / kPtsPerBlock;  // warp-uniform
        if (t >= npts) continue;
        const float cs = s_cos[i], sn = s_sin[i];
        const float gy = s_x[3 * t + 1];
        float gx, gz;
        vote_center(s_p[3 * t], s_p[3 * t + 2], s_x[3 * t], s_x[3 * t + 2], cs, sn, g, gx, gz);
        if (!vote_in_bounds(gx, gy, gz, g)) continue;

        const int fx = (int)gx, fy = (int)gy, fz = (int)gz;                          // make_int3 (:45)
        const float rx = gx - floorf(gx), ry = gy - floorf(gy), rz = gz - floorf(gz);  // fracf (:47)
        const float wx0 = 1.f - rx, wy0 = 1.f - ry, wz0 = 1.f - rz;
        const float objness = s_o[t];
        const float s0 = s_s[3 * t], s1 = s_s[3 * t + 1], s2 = s_s[3 * t + 2];
        const int64_t v0 = (int64_t)fx * YZ + (int64_t)fy * g.Z + fz;
#pragma unroll
        for (int a = 0; a < 2; a++) {
#pragma unroll
            for (int b = 0; b < 2; b++) {
                const float wxy = __fmul_rn(a ? rx : wx0, b ? ry : wy0);
                const int64_t vab = v0 + (int64_t)a * YZ + (int64_t)b * g.Z;
#pragma unroll
                for (int d = 0; d < 2; d++) {
                    // ((wx*wy)*wz)*objness, the reference's association (:52-59)
                    const float w = __fmul_rn(__fmul_rn(wxy, d ? rz : wz0), objness);
                    float *sec = work + 8 * (vab + d);
                    red_add_v4(sec, w, __fmul_rn(w, cs), __fmul_rn(w, sn), __fmul_rn(w, s0));
                    red_add_v2(sec + 4, __fmul_rn(w, s1), __fmul_rn(w, s2));
                }
            }
        }
    }
}

// ------------------------------------------------------------------ write-out ----
// x / (w + 1e-7) as the reference average kernel evaluates it (:100-119): the literal is a double, so the
// quotient is float(double(x) / (double(w) + 1e-7)).  One reciprocal per voxel and a Markstein correction
// step per channel give the correctly rounded double quotient (which is then rounded to float).
__device__ __forceinline__ void avg_voxel(const float4 a, const float4 b, float &rc, float &rs, float &s0, float &s1,
                                          float &s2) {
    rc = rs = s0 = s1 = s2 = 0.f;
    if (a.x == 0.f && a.y == 0.f && a.z == 0.f && a.w == 0.f && b.x == 0.f && b.y == 0.f) return;
    const double den = (double)a.x + 1e-7;
    if (!(fabs(den) > 1e-300 && fabs(den) < 1e300)) {   // degenerate denominators: plain IEEE division
        rc = (float)((double)a.y / den); rs = (float)((double)a.z / den);
        s0 = (float)((double)a.w / den); s1 = (float)((double)b.x / den); s2 = (float)((double)b.y / den);
        return;
    }
    const double r = __drcp_rn(den);
    auto div = [&](float xf) {
        const double x = (double)xf;
        const double q0 = x * r;
        const double e = __fma_rn(-den, q0, x);
        return (float)__fma_rn(e, r, q0);
    };
    rc = div(a.y); rs = div(a.z); s0 = div(a.w); s1 = div(b.x); s2 = div(b.y);
}

constexpr int kFinalizeThreads = 256;
constexpr int kVoxPerThread = 4;

// One thread = 4 consecutive voxels: 8 x 16-byte workspace loads, zero stores to the touched sectors only (the
// workspace is left all-zero for the next call) and 6 x 16-byte streaming output stores.
__global__ void __launch_bounds__(kFinalizeThreads)
hv_finalize_kernel(float4 *__restrict__ work, int64_t G, float *__restrict__ grid_obj, float *__restrict__ grid_rot,
                   float *__restrict__ grid_scale) {
    const int64_t v = ((int64_t)blockIdx.x * kFinalizeThreads + threadIdx.x) * kVoxPerThread;
    if (v >= G) return;
    float o4[4] = {0.f, 0.f, 0.f, 0.f}, r8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
          s12[12] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int nv = (int)min((int64_t)kVoxPerThread, G - v);
    float4 a[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (j < nv) {
            a[j] = __ldcs(work + 2 * (v + j));
            b[j] = __ldcs(work + 2 * (v + j) + 1);
        }
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (j < nv) {
            // re-zero only the sectors a vote touched (a voxel = one 32-byte sector): untouched ones are already zero,
            // which saves ~3/4 of the zeroing traffic on typical scenes
            const bool touched = a[j].x != 0.f || a[j].y != 0.f || a[j].z != 0.f || a[j].w != 0.f || b[j].x != 0.f || b[j].y != 0.f;
            if (touched) {
                work[2 * (v + j)] = zero;
                work[2 * (v + j) + 1] = zero;
            }
            o4[j] = a[j].x;
            avg_voxel(a[j], b[j], r8[2 * j], r8[2 * j + 1], s12[3 * j], s12[3 * j + 1], s12[3 * j + 2]);
        }
    if (nv == kVoxPerThread) {   // v is a multiple of 4: all three addresses are 16-byte aligned
        __stcs(reinterpret_cast<float4 *>(grid_obj + v), make_float4(o4[0], o4[1], o4[2], o4[3]));
        __stcs(reinterpret_cast<float4 *>(grid_rot + 2 * v), make_float4(r8[0], r8[1], r8[2], r8[3]));
        __stcs(reinterpret_cast<float4 *>(grid_rot + 2 * v) + 1, make_float4(r8[4], r8[5], r8[6], r8[7]));
#pragma unroll
        for (int k = 0; k < 3; k++)
            __stcs(reinterpret_cast<float4 *>(grid_scale + 3 * v) + k,
                   make_float4(s12[4 * k], s12[4 * k + 1], s12[4 * k + 2], s12[4 * k + 3]));
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (j < nv) {
                grid_obj[v + j] = o4[j];
                grid_rot[2 * (v + j)] = r8[2 * j];
                grid_rot[2 * (v + j) + 1] = r8[2 * j + 1];
                grid_scale[3 * (v + j)] = s12[3 * j];
                grid_scale[3 * (v + j) + 1] = s12[3 * j + 1];
                grid_scale[3 * (v + j) + 2] = s12[3 * j + 2];
            }
    }
}

// ------------------------------------------------------------------ forward, sorted-tile path ------
// No float atomics and no workspace grid: the in-bounds votes are binned by the 8x8x8-voxel tile(s) their 2x2x2 footprint
// touches (count -> scan -> fill: a counting sort on 4-byte item ids, item = point * num_rots + theta), then ONE CTA per
// tile accumulates its votes in shared memory and writes the tile's 512 voxels of the three output grids exactly once,
// already normalised.  Inside the tile kernel conflicts are excluded by construction instead of by atomics:
//   * the tile's records are sorted by base cell (shared-memory counting sort: one integer atomic per record), so all
//     votes of a cell form one run, owned by one thread (runs of <= kLmax votes) or one warp (longer runs, tree-reduced
//     with shuffles -- the peaks of the vote map are exactly such cells);
//   * in phase d (one per corner of the 2x2x2 footprint) every owner adds its cell's corner-d sum to voxel cell + d with
//     a plain read-modify-write: two different cells never touch the same voxel in the same phase.
// Bytes: 40 N read + 24 G written (the contract figure) + 4 bytes per record written and read twice (~2 MB at C2).
constexpr int kTE = 8;                      // tile edge in voxels
constexpr int kTV = kTE * kTE * kTE;        // 512 voxels per tile
constexpr int kCE = kTE + 1;                // base cells per edge whose footprint touches the tile: -1 .. 7
constexpr int kNC = kCE * kCE * kCE;        // 729
constexpr int kNCpad = 768;
constexpr int kTileThreads = 256;
constexpr int kChunk = 512;                 // records sorted / accumulated at a time
constexpr int kLmax = 16;                   // longest run a single thread sums
constexpr int kLongMax = kChunk / (kLmax + 1) + 2;
constexpr int kThetaSmem = 256;             // cos/sin table in shared memory up to this many rotations

struct HvTiles {
    int ntx, nty, ntz, n;
};

__host__ __device__ inline HvTiles make_tiles(int X, int Y, int Z) {
    HvTiles t;
    t.ntx = (X + kTE - 1) / kTE; t.nty = (Y + kTE - 1) / kTE; t.ntz = (Z + kTE - 1) / kTE;
    t.n = t.ntx * t.nty * t.ntz;
    return t;
}

// workspace (uint32 units): count[T] | start[T+1] | cursor[T] | ticket | pad | records[8 n num_rots]
struct HvBinWork {
    unsigned int *count, *start, *cursor, *ticket, *rec;
};

static inline size_t bin_header_words(int T) { return ((size_t)3 * T + 2 + 3) / 4 * 4; }

static inline HvBinWork bin_work(void *d_work, int T) {
    HvBinWork w;
    unsigned int *p = (unsigned int *)d_work;
    w.count = p; w.start = p + T; w.cursor = p + 2 * (size_t)T + 1; w.ticket = p + 3 * (size_t)T + 1;
    w.rec = p + bin_header_words(T);
    return w;
}

__device__ __forceinline__ int block_exclusive_scan_256(int v, int *warp_sums, int &total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int s = warp_sums[k];
        if (k < w) base += s;
        tot += s;
    }
    __syncthreads();
    total = tot;
    return base + incl - v;
}

// Binning pass over the (point, theta) items: kFill = false counts records per tile (and the last CTA to finish turns the
// counts into offsets), kFill = true writes the item ids into the tiles' segments.
template <bool kFill>
__global__ void __launch_bounds__(kScatterThreads)
hv_bin_kernel(const float *__restrict__ points, const float *__restrict__ xyz, const float *__restrict__ scale, int64_t n, int num_rots,
              HvGeom g, HvTiles tl, HvBinWork W) {
    extern __shared__ float s_theta[];  // [2 * num_rots]
    __shared__ float s_p[3 * kPtsPerBlock];
    __shared__ float s_x[3 * kPtsPerBlock];
    __shared__ float s_s[3 * kPtsPerBlock];
    __shared__ int s_scan[8];
    __shared__ bool s_last;
    float *s_cos = s_theta, *s_sin = s_theta + num_rots;
    fill_theta_table(s_cos, s_sin, num_rots);
    const int64_t p0 = (int64_t)blockIdx.x * kPtsPerBlock;
    const int npts = (int)min((int64_t)kPtsPerBlock, n - p0);
    for (int k = threadIdx.x; k < 3 * npts; k += kScatterThreads) {
        s_p[k] = __ldg(points + 3 * p0 + k);
        s_x[k] = __ldg(xyz + 3 * p0 + k);
        s_s[k] = __ldg(scale + 3 * p0 + k);
    }
    __syncthreads();
    if (threadIdx.x < npts) {
        const int t = threadIdx.x;
        float corr_x, corr_z, gy;
        point_prep(s_p[3 * t + 1], s_x[3 * t], s_x[3 * t + 1], s_x[3 * t + 2], s_s[3 * t], s_s[3 * t + 1], s_s[3 * t + 2], g, corr_x, corr_z, gy);
        s_x[3 * t] = corr_x;
        s_x[3 * t + 1] = gy;
        s_x[3 * t + 2] = corr_z;
    }
    __syncthreads();
    const int items = kPtsPerBlock * num_rots;
    for (int j = threadIdx.x; j < items; j += kScatterThreads) {
        const int t = j & (kPtsPerBlock - 1);
        const int i = j / kPtsPerBlock;  // warp-uniform
        if (t >= npts) continue;
        const float gy = s_x[3 * t + 1];
        float gx, gz;
        vote_center(s_p[3 * t], s_p[3 * t + 2], s_x[3 * t], s_x[3 * t + 2], s_cos[i], s_sin[i], g, gx, gz);
        if (!vote_in_bounds(gx, gy, gz, g)) continue;
        const int fx = (int)gx, fy = (int)gy, fz = (int)gz;
        const int tx0 = fx / kTE, tx1 = (fx + 1) / kTE, ty0 = fy / kTE, ty1 = (fy + 1) / kTE, tz0 = fz / kTE, tz1 = (fz + 1) / kTE;
        const unsigned int item = (unsigned int)((p0 + t) * num_rots + i);
        for (int tx = tx0; tx <= tx1; tx++)
            for (int ty = ty0; ty <= ty1; ty++)
                for (int tz = tz0; tz <= tz1; tz++) {
                    const int tile = (tx * tl.nty + ty) * tl.ntz + tz;
                    if (kFill) W.rec[atomicAdd(W.cursor + tile, 1u)] = item;
                    else atomicAdd(W.count + tile, 1u);
                }
    }
    if (kFill) return;
    // the last CTA to get here scans the counts: start[t] = cursor[t] = exclusive prefix, count[t] = 0 again
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(W.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    int carry = 0;
    for (int b0 = 0; b0 < tl.n; b0 += 4 * kScatterThreads) {
        const int e0 = b0 + 4 * threadIdx.x;
        int c[4];
#pragma unroll
        for (int k = 0; k < 4; k++) c[k] = e0 + k < tl.n ? (int)__ldcg(W.count + e0 + k) : 0;
        int total;
        int ex = carry + block_exclusive_scan_256(c[0] + c[1] + c[2] + c[3], s_scan, total);
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (e0 + k < tl.n) {
                W.start[e0 + k] = (unsigned int)ex;
                W.cursor[e0 + k] = (unsigned int)ex;
                W.count[e0 + k] = 0u;
                ex += c[k];
            }
        carry += total;
    }
    if (threadIdx.x == 0) {
        W.start[tl.n] = (unsigned int)carry;
        *W.ticket = 0u;
    }
}

struct HvTileSmem {
    float acc[kTV * 8];                // voxel v: acc[8 v + {obj, rot_cos, rot_sin, scale0, scale1, scale2, -, -}]
    float pay[9][kChunk];              // sorted records: rx, ry, rz, objness, cos, sin, scale0..2
    unsigned short cell[kChunk];       // base cell of the sorted record (tile-local, 0 .. 728)
    int hist[kNCpad];                  // per-cell counts, then exclusive offsets (entries >= 729 hold the chunk size)
    int warp_sums[8];
    int long_cells[kLongMax];
    int n_long;
    float cs[kThetaSmem], sn[kThetaSmem];
};

// corner-d contribution of sorted record j, added to v[0..5]
__device__ __forceinline__ void tile_add_record(const HvTileSmem &S, int j, int a, int b, int d, float (&v)[6]) {
    const float rx = S.pay[0][j], ry = S.pay[1][j], rz = S.pay[2][j];
    const float wxy = __fmul_rn(a ? rx : 1.f - rx, b ? ry : 1.f - ry);
    const float w = __fmul_rn(__fmul_rn(wxy, d ? rz : 1.f - rz), S.pay[3][j]);     // ((wx*wy)*wz)*objness  (:52-59)
    v[0] += w;
    v[1] += __fmul_rn(w, S.pay[4][j]);
    v[2] += __fmul_rn(w, S.pay[5][j]);
    v[3] += __fmul_rn(w, S.pay[6][j]);
    v[4] += __fmul_rn(w, S.pay[7][j]);
    v[5] += __fmul_rn(w, S.pay[8][j]);
}

__device__ __forceinline__ void tile_rmw(HvTileSmem &S, int c, int a, int b, int d, const float (&v)[6]) {
    const int lx = c / (kCE * kCE) - 1 + a, ly = (c / kCE) % kCE - 1 + b, lz = c % kCE - 1 + d;
    if ((unsigned)lx >= (unsigned)kTE || (unsigned)ly >= (unsigned)kTE || (unsigned)lz >= (unsigned)kTE) return;   // the neighbour tile's voxel
    float4 *p = reinterpret_cast<float4 *>(S.acc + 8 * ((lx * kTE + ly) * kTE + lz));
    float4 u = p[0];
    float2 t = *reinterpret_cast<float2 *>(p + 1);
    u.x += v[0]; u.y += v[1]; u.z += v[2]; u.w += v[3];
    t.x += v[4]; t.y += v[5];
    p[0] = u;
    *reinterpret_cast<float2 *>(p + 1) = t;
}

__global__ void __launch_bounds__(kTileThreads)
hv_tile_kernel(const float *__restrict__ points, const float *__restrict__ xyz, const float *__restrict__ scale,
               const float *__restrict__ obj, int num_rots, HvGeom g, HvTiles tl, const unsigned int *__restrict__ start,
               const unsigned int *__restrict__ rec, float *__restrict__ grid_obj, float *__restrict__ grid_rot,
               float *__restrict__ grid_scale) {
    __shared__ HvTileSmem S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int tz = tile % tl.ntz, ty = (tile / tl.ntz) % tl.nty, tx = tile / (tl.ntz * tl.nty);
    const int x0 = tx * kTE, y0 = ty * kTE, z0 = tz * kTE;
    const unsigned int r0 = __ldg(start + tile), r1 = __ldg(start + tile + 1);
    const int nrec = (int)(r1 - r0);

    if (nrec > 0) {
        for (int e = tid; e < kTV * 2; e += kTileThreads) reinterpret_cast<float4 *>(S.acc)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool table = num_rots <= kThetaSmem;
        if (table) fill_theta_table(S.cs, S.sn, num_rots);
        for (int c0 = 0; c0 < nrec; c0 += kChunk) {
            const int m = min(kChunk, nrec - c0);
            for (int e = tid; e < kNCpad; e += kTileThreads) S.hist[e] = 0;
            if (tid == 0) S.n_long = 0;
            __syncthreads();
            // ---- step 1: recompute the vote of every record, count per base cell
            int lc[kChunk / kTileThreads], rank[kChunk / kTileThreads];
            float pl[kChunk / kTileThreads][9];
#pragma unroll
            for (int q = 0; q < kChunk / kTileThreads; q++) {
                const int i = tid + q * kTileThreads;
                lc[q] = -1;
                if (i < m) {
                    const unsigned int item = __ldg(rec + r0 + c0 + i);
                    const unsigned int p = item / (unsigned int)num_rots;
                    const int th = (int)(item - p * (unsigned int)num_rots);
                    const float *pp = points + 3 * (size_t)p, *px = xyz + 3 * (size_t)p, *ps = scale + 3 * (size_t)p;
                    const float ptx = __ldg(pp), pty = __ldg(pp + 1), ptz = __ldg(pp + 2);
                    const float xx = __ldg(px), xy = __ldg(px + 1), xz = __ldg(px + 2);
                    const float sx = __ldg(ps), sy = __ldg(ps + 1), sz = __ldg(ps + 2);
                    float corr_x, corr_z, gy, gx, gz, cs, sn;
                    point_prep(pty, xx, xy, xz, sx, sy, sz, g, corr_x, corr_z, gy);
                    if (table) { cs = S.cs[th]; sn = S.sn[th]; }
                    else theta_cs(th, num_rots, cs, sn);
                    vote_center(ptx, ptz, corr_x, corr_z, cs, sn, g, gx, gz);
                    const int fx = (int)gx, fy = (int)gy, fz = (int)gz;                       // make_int3 (:45); in bounds by construction
                    pl[q][0] = gx - floorf(gx); pl[q][1] = gy - floorf(gy); pl[q][2] = gz - floorf(gz);   // fracf (:47)
                    pl[q][3] = __ldg(obj + p);
                    pl[q][4] = cs; pl[q][5] = sn; pl[q][6] = sx; pl[q][7] = sy; pl[q][8] = sz;
                    lc[q] = ((fx - x0 + 1) * kCE + (fy - y0 + 1)) * kCE + (fz - z0 + 1);
                    rank[q] = atomicAdd(&S.hist[lc[q]], 1);
                }
            }
            __syncthreads();
            // ---- step 2: exclusive scan of the 768 counters (3 per thread)
            {
                const int a0 = S.hist[3 * tid], a1 = S.hist[3 * tid + 1], a2 = S.hist[3 * tid + 2];
                int total;
                const int ex = block_exclusive_scan_256(a0 + a1 + a2, S.warp_sums, total);
                S.hist[3 * tid] = ex; S.hist[3 * tid + 1] = ex + a0; S.hist[3 * tid + 2] = ex + a0 + a1;
            }
            __syncthreads();
            // ---- step 3: records into their sorted slots
#pragma unroll
            for (int q = 0; q < kChunk / kTileThreads; q++)
                if (lc[q] >= 0) {
                    const int pos = S.hist[lc[q]] + rank[q];
#pragma unroll
                    for (int k = 0; k < 9; k++) S.pay[k][pos] = pl[q][k];
                    S.cell[pos] = (unsigned short)lc[q];
                }
            __syncthreads();
            // ---- step 4: short runs, one owner thread per run
            for (int p0 = 0; p0 < m; p0 += kTileThreads) {
                const int pos = p0 + tid;
                int c = 0, st = 0, en = 0;
                bool own = false;
                if (pos < m) {
                    c = S.cell[pos];
                    st = S.hist[c]; en = S.hist[c + 1];
                    if (pos == st) {
                        own = en - st <= kLmax;
                        if (!own) S.long_cells[atomicAdd(&S.n_long, 1)] = c;
                    }
                }
#pragma unroll
                for (int d8 = 0; d8 < 8; d8++) {
                    const int a = d8 >> 2, b = (d8 >> 1) & 1, d = d8 & 1;
                    if (own) {
                        float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        for (int j = st; j < en; j++) tile_add_record(S, j, a, b, d, v);
                        tile_rmw(S, c, a, b, d, v);
                    }
                    __syncthreads();
                }
            }
            // ---- long runs (the peaks of the vote map): one warp per run, tree-reduced
            const int n_long = S.n_long;          // complete: every owner passed the barriers above
            for (int j0 = 0; j0 < n_long; j0 += kTileThreads / 32) {
                const bool have = j0 + warp < n_long;
                const int c = have ? S.long_cells[j0 + warp] : 0;
                const int st = have ? S.hist[c] : 0, en = have ? S.hist[c + 1] : 0;
#pragma unroll
                for (int d8 = 0; d8 < 8; d8++) {
                    const int a = d8 >> 2, b = (d8 >> 1) & 1, d = d8 & 1;
                    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    for (int j = st + lane; j < en; j += 32) tile_add_record(S, j, a, b, d, v);
#pragma unroll
                    for (int k = 0; k < 6; k++)
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
                    if (have && lane == 0) tile_rmw(S, c, a, b, d, v);
                    __syncthreads();
                }
            }
        }
    }
    // ---- write-out: every voxel of the tile exactly once, normalised like hv_cuda_average_kernel (:100-119)
    const int64_t YZ = (int64_t)g.Y * g.Z;
    for (int v = tid; v < kTV; v += kTileThreads) {
        const int X = x0 + (v >> 6), Y = y0 + ((v >> 3) & 7), Z = z0 + (v & 7);
        if (X >= g.X || Y >= g.Y || Z >= g.Z) continue;
        const int64_t V = (int64_t)X * YZ + (int64_t)Y * g.Z + Z;
        float o = 0.f, rc = 0.f, rs = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f;
        if (nrec > 0) {
            const float4 a = reinterpret_cast<const float4 *>(S.acc)[2 * v], b = reinterpret_cast<const float4 *>(S.acc)[2 * v + 1];
            o = a.x;
            avg_voxel(a, b, rc, rs, s0, s1, s2);
        }
        __stcs(grid_obj + V, o);
        __stcs(reinterpret_cast<float2 *>(grid_rot + 2 * V), make_float2(rc, rs));
        __stcs(grid_scale + 3 * V, s0);
        __stcs(grid_scale + 3 * V + 1, s1);
        __stcs(grid_scale + 3 * V + 2, s2);
    }
}

// ------------------------------------------------------------------ backward -----
// hv_cuda_backward_kernel (:183-259): per point, gather the 8 corners of grad_grid for
// every theta.  One thread per point like the reference, but accumulating in registers
// (the reference does 14 global read-modify-writes per theta) and with the theta table
// in shared memory.
constexpr int kBackwardThreads = 128;

__global__ void __launch_bounds__(kBackwardThreads)
hv_backward_kernel(const float *__restrict__ grad_grid, const float *__restrict__ points,
                   const float *__restrict__ xyz, const float *__restrict__ scale, const float *__restrict__ obj,
                   int64_t n, int num_rots, HvGeom g, float *__restrict__ d_xyz, float *__restrict__ d_scale,
                   float *__restrict__ d_obj) {
    extern __shared__ float s_theta[];
    float *s_cos = s_theta, *s_sin = s_theta + num_rots;
    fill_theta_table(s_cos, s_sin, num_rots);
    __syncthreads();
    const int64_t c = (int64_t)blockIdx.x * kBackwardThreads + threadIdx.x;
    if (c >= n) return;
    const float px = points[3 * c], py = points[3 * c + 1], pz = points[3 * c + 2];
    const float xx = xyz[3 * c], xy = xyz[3 * c + 1], xz = xyz[3 * c + 2];
    const float sx = scale[3 * c], sy = scale[3 * c + 1], sz = scale[3 * c + 2];
    const float objness = obj[c];
    float corr_x, corr_z, gy;
    point_prep(py, xx, xy, xz, sx, sy, sz, g, corr_x, corr_z, gy);
    const int64_t YZ = (int64_t)g.Y * g.Z;
    float dobj = 0.f, dcx_acc = 0.f, dcy_acc = 0.f, dcz_acc = 0.f;
    for (int i = 0; i < num_rots; i++) {
        const float cs = s_cos[i], sn = s_sin[i];
        float gx, gz;
        vote_center(px, pz, corr_x, corr_z, cs, sn, g, gx, gz);
        if (!vote_in_bounds(gx, gy, gz, g)) continue;
        const int fx = (int)gx, fy = (int)gy, fz = (int)gz;
        const float rx = gx - floorf(gx), ry = gy - floorf(gy), rz = gz - floorf(gz);
        const float wx[2] = {1.f - rx, rx}, wy[2] = {1.f - ry, ry}, wz[2] = {1.f - rz, rz};
        const float *base = grad_grid + ((int64_t)fx * YZ + (int64_t)fy * g.Z + fz);
        float ddx = 0.f, ddy = 0.f, ddz = 0.f;
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 2; b++)
#pragma unroll
                for (int d = 0; d < 2; d++) {
                    const float gg = __ldg(base + (int64_t)a * YZ + (int64_t)b * g.Z + d);
                    dobj += gg * wx[a] * wy[b] * wz[d];                    // :210-217
                    ddx += (a ? gg : -gg) * wy[b] * wz[d];                 // :219-227
                    ddy += (b ? gg : -gg) * wx[a] * wz[d];                 // :228-235
                    ddz += (d ? gg : -gg) * wx[a] * wy[b];                 // :236-243
                }
        ddx *= objness;
        ddy *= objness;
        ddz *= objness;
        // d_corr (:249-250); d_xyz += d_corr * scale, d_scale += d_corr * xyz (:252-258)
        dcx_acc += -cs * ddx - sn * ddz;
        dcy_acc += -ddy;
        dcz_acc += sn * ddx - cs * ddz;
    }
    d_obj[c] = dobj;
    d_xyz[3 * c] = dcx_acc * sx;
    d_xyz[3 * c + 1] = dcy_acc * sy;
    d_xyz[3 * c + 2] = dcz_acc * sz;
    d_scale[3 * c] = dcx_acc * xx;
    d_scale[3 * c + 1] = dcy_acc * xy;
    d_scale[3 * c + 2] = dcz_acc * xz;
}

// ------------------------------------------------------------------ indices ------
__global__ void __launch_bounds__(256)
hv_vote_indices_kernel(const float *__restrict__ points, const float *__restrict__ xyz,
                       const float *__restrict__ scale, int64_t n, int num_rots, HvGeom g,
                       int32_t *__restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n * num_rots) return;
    const int64_t c = j / num_rots;
    const int i = (int)(j - c * num_rots);
    float corr_x, corr_z, gy, gx, gz, cs, sn;
    point_prep(points[3 * c + 1], xyz[3 * c], xyz[3 * c + 1], xyz[3 * c + 2], scale[3 * c], scale[3 * c + 1],
               scale[3 * c + 2], g, corr_x, corr_z, gy);
    theta_cs(i, num_rots, cs, sn);
    vote_center(points[3 * c], points[3 * c + 2], corr_x, corr_z, cs, sn, g, gx, gz);
    const bool ok = vote_in_bounds(gx, gy, gz, g);
    out[3 * j] = ok ? (int)gx : -1;
    out[3 * j + 1] = ok ? (int)gy : -1;
    out[3 * j + 2] = ok ? (int)gz : -1;
}

__global__ void hv_theta_table_kernel(int num_rots, float *__restrict__ d_cos, float *__restrict__ d_sin) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < num_rots) theta_cs(i, num_rots, d_cos[i], d_sin[i]);
}

// ------------------------------------------------------------------ grid dims ----
struct HvDimsHeader {
    float corner[3];
    float maxpt[3];
    int32_t dims[3];
};
constexpr int kMinmaxThreads = 256;
constexpr int kMinmaxMaxBlocks = 2 * kNumSMs;
// work layout: [0,64) header | [64,128) ticket counter | [128, ...) partials [blocks][6]
constexpr size_t kDimsWorkBytes = 128 + sizeof(float) * 6 * kMinmaxMaxBlocks;

__global__ void __launch_bounds__(kMinmaxThreads)
hv_minmax_kernel(const float *__restrict__ points, int64_t n, float res, unsigned char *__restrict__ workb) {
    HvDimsHeader *hdr = reinterpret_cast<HvDimsHeader *>(workb);
    unsigned int *ticket = reinterpret_cast<unsigned int *>(workb + 64);
    float *partial = reinterpret_cast<float *>(workb + 128);
    __shared__ float s_red[kMinmaxThreads / 32][6];
    __shared__ bool s_last;

    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t c = (int64_t)blockIdx.x * kMinmaxThreads + threadIdx.x; c < n; c += (int64_t)gridDim.x * kMinmaxThreads) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float v = __ldg(points + 3 * c + k);
            mn[k] = fminf(mn[k], v);
            mx[k] = fmaxf(mx[k], v);
        }
    }
    auto block_reduce = [&]() {
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
                mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
            }
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        if (l == 0)
            for (int k = 0; k < 3; k++) { s_red[w][k] = mn[k]; s_red[w][3 + k] = mx[k]; }
        __syncthreads();
        if (threadIdx.x < 32) {
            for (int k = 0; k < 3; k++) {
                mn[k] = l < kMinmaxThreads / 32 ? s_red[l][k] : INFINITY;
                mx[k] = l < kMinmaxThreads / 32 ? s_red[l][3 + k] : -INFINITY;
            }
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
                    mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
                }
        }
        __syncthreads();
    };
    block_reduce();
    if (threadIdx.x == 0) {
        for (int k = 0; k < 3; k++) { partial[6 * blockIdx.x + k] = mn[k]; partial[6 * blockIdx.x + 3 + k] = mx[k]; }
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int k = 0; k < 3; k++) { mn[k] = INFINITY; mx[k] = -INFINITY; }
    for (int b = threadIdx.x; b < (int)gridDim.x; b += kMinmaxThreads)
        for (int k = 0; k < 3; k++) {
            mn[k] = fminf(mn[k], __ldcg(partial + 6 * b + k));
            mx[k] = fmaxf(mx[k], __ldcg(partial + 6 * b + 3 + k));
        }
    block_reduce();
    if (threadIdx.x == 0) {
        for (int k = 0; k < 3; k++) {
            hdr->corner[k] = mn[k];
            hdr->maxpt[k] = mx[k];
            // diff = (max - min) / res in float32, int() truncation, + 1   (:131-134)
            hdr->dims[k] = (int32_t)__fdiv_rn(__fsub_rn(mx[k], mn[k]), res) + 1;
        }
        *ticket = 0;
    }
}

static int make_geom(const float corner[3], const int32_t dims[3], float res, HvGeom &g) {
    CVB_REQUIRE(corner && dims, CVB200_EINVAL, "corner/dims must not be NULL");
    CVB_REQUIRE(dims[0] > 0 && dims[1] > 0 && dims[2] > 0, CVB200_EINVAL, "grid dims must be positive (%d,%d,%d)",
                dims[0], dims[1], dims[2]);
    CVB_REQUIRE((int64_t)dims[0] * dims[1] * dims[2] < ((int64_t)1 << 31), CVB200_EINVAL,
                "grid of %lld voxels exceeds the 2^31 voxel limit", (long long)dims[0] * dims[1] * dims[2]);
    g.cx = corner[0]; g.cy = corner[1]; g.cz = corner[2];
    g.res = res;
    g.X = dims[0]; g.Y = dims[1]; g.Z = dims[2];
    return 0;
}

}  // namespace cvb200

using namespace cvb200;

extern "C" size_t cvb200_hv_grid_dims_work_bytes(void) { return kDimsWorkBytes; }

extern "C" int cvb200_hv_grid_dims(const float *d_points, int64_t n, float res, void *d_work, float *h_corner,
                                   float *h_maxpt, int32_t *h_dims, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVB_REQUIRE(d_points && d_work && h_corner && h_dims, CVB200_EINVAL, "hv_grid_dims: NULL argument");
    CVB_REQUIRE(n > 0, CVB200_EEMPTY, "hv_grid_dims: empty point set (the reference's torch::min fails too)");
    unsigned char *workb = (unsigned char *)d_work;
    CVB_CUDA(cudaMemsetAsync(workb + 64, 0, 64, stream));
    const int blocks = (int)std::min<int64_t>(kMinmaxMaxBlocks, ceil_div(n, kMinmaxThreads));
    hv_minmax_kernel<<<blocks, kMinmaxThreads, 0, stream>>>(d_points, n, res, workb);
    CVB_LAUNCH_CHECK("hv_minmax_kernel");
    HvDimsHeader h;
    CVB_CUDA(cudaMemcpyAsync(&h, workb, sizeof(h), cudaMemcpyDeviceToHost, stream));
    CVB_CUDA(cudaStreamSynchronize(stream));
    for (int k = 0; k < 3; k++) {
        h_corner[k] = h.corner[k];
        if (h_maxpt) h_maxpt[k] = h.maxpt[k];
        h_dims[k] = h.dims[k];
    }
    return 0;
}

static int g_hv_impl = 1;   // 1: sorted tiles (default); 0: vector reductions into a [G][8] workspace + write-out pass (A/B measurements)

extern "C" int cvb200_hv_set_impl(int32_t impl) {
    g_hv_impl = impl ? 1 : 0;
    return 0;
}

extern "C" size_t cvb200_hv_forward_work_bytes(const int32_t dims[3]) {
    if (!dims || dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) return 0;
    return (size_t)dims[0] * dims[1] * dims[2] * 8 * sizeof(float);
}

extern "C" size_t cvb200_hv_forward_work_bytes_n(const int32_t dims[3], int64_t n, int32_t num_rots) {
    if (!dims || dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0 || n < 0 || num_rots <= 0) return 0;
    const HvTiles tl = make_tiles(dims[0], dims[1], dims[2]);
    // a vote's 2x2x2 footprint touches at most 8 tiles
    return sizeof(unsigned int) * (bin_header_words(tl.n) + (size_t)8 * (size_t)n * (size_t)num_rots);
}

extern "C" int cvb200_hv_forward(const float *d_points, const float *d_xyz, const float *d_scale,
                                 const float *d_obj, int64_t n, float res, int32_t num_rots,
                                 const float corner[3], const int32_t dims[3], float *d_grid_obj,
                                 float *d_grid_rot, float *d_grid_scale, void *d_work, size_t work_bytes,
                                 void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    HvGeom g;
    if (int rc = make_geom(corner, dims, res, g)) return rc;
    CVB_REQUIRE(n >= 0 && num_rots > 0 && num_rots <= 4096, CVB200_EINVAL, "hv_forward: bad n=%lld / num_rots=%d",
                (long long)n, num_rots);
    CVB_REQUIRE(d_grid_obj && d_grid_rot && d_grid_scale && d_work, CVB200_EINVAL, "hv_forward: NULL output/work");
    CVB_REQUIRE(n == 0 || (d_points && d_xyz && d_scale && d_obj), CVB200_EINVAL, "hv_forward: NULL input");
    CVB_REQUIRE(((reinterpret_cast<uintptr_t>(d_work) | reinterpret_cast<uintptr_t>(d_grid_obj) |
                  reinterpret_cast<uintptr_t>(d_grid_rot) | reinterpret_cast<uintptr_t>(d_grid_scale)) & 15) == 0,
                CVB200_EINVAL, "hv_forward: workspace and outputs must be 16-byte aligned");
    const int64_t G = (int64_t)g.X * g.Y * g.Z;
    if (g_hv_impl == 0) {
        const size_t need = cvb200_hv_forward_work_bytes(dims);
        CVB_REQUIRE(work_bytes >= need, CVB200_ESCRATCH, "hv_forward: workspace %zu < %zu bytes", work_bytes, need);
        if (n > 0) {
            const int64_t blocks = ceil_div(n, kPtsPerBlock);
            hv_scatter_kernel<<<(unsigned)blocks, kScatterThreads, 2 * num_rots * sizeof(float), stream>>>(
                d_points, d_xyz, d_scale, d_obj, n, num_rots, g, (float *)d_work);
            CVB_LAUNCH_CHECK("hv_scatter_kernel");
        }
        hv_finalize_kernel<<<(unsigned)ceil_div(G, kFinalizeThreads * kVoxPerThread), kFinalizeThreads, 0, stream>>>(
            (float4 *)d_work, G, d_grid_obj, d_grid_rot, d_grid_scale);
        CVB_LAUNCH_CHECK("hv_finalize_kernel");
        return 0;
    }
    const size_t need = cvb200_hv_forward_work_bytes_n(dims, n, num_rots);
    CVB_REQUIRE(work_bytes >= need, CVB200_ESCRATCH, "hv_forward: workspace %zu < %zu bytes", work_bytes, need);
    CVB_REQUIRE(n * (int64_t)num_rots < ((int64_t)1 << 32), CVB200_EINVAL, "hv_forward: n * num_rots = %lld exceeds the 2^32 item limit",
                (long long)(n * num_rots));
    if (n == 0) {
        CVB_CUDA(cudaMemsetAsync(d_grid_obj, 0, sizeof(float) * G, stream));
        CVB_CUDA(cudaMemsetAsync(d_grid_rot, 0, sizeof(float) * 2 * G, stream));
        CVB_CUDA(cudaMemsetAsync(d_grid_scale, 0, sizeof(float) * 3 * G, stream));
        return 0;
    }
    const HvTiles tl = make_tiles(g.X, g.Y, g.Z);
    const HvBinWork W = bin_work(d_work, tl.n);
    const unsigned blocks = (unsigned)ceil_div(n, kPtsPerBlock);
    const size_t theta_bytes = 2 * num_rots * sizeof(float);
    hv_bin_kernel<false><<<blocks, kScatterThreads, theta_bytes, stream>>>(d_points, d_xyz, d_scale, n, num_rots, g, tl, W);
    CVB_LAUNCH_CHECK("hv_bin_kernel<count>");
    hv_bin_kernel<true><<<blocks, kScatterThreads, theta_bytes, stream>>>(d_points, d_xyz, d_scale, n, num_rots, g, tl, W);
    CVB_LAUNCH_CHECK("hv_bin_kernel<fill>");
    hv_tile_kernel<<<(unsigned)tl.n, kTileThreads, 0, stream>>>(d_points, d_xyz, d_scale, d_obj, num_rots, g, tl, W.start, W.rec, d_grid_obj,
                                                               d_grid_rot, d_grid_scale);
    CVB_LAUNCH_CHECK("hv_tile_kernel");
    return 0;
}

extern "C" int cvb200_hv_backward(const float *d_grad_grid, const float *d_points, const float *d_xyz,
                                  const float *d_scale, const float *d_obj, int64_t n, float res,
                                  int32_t num_rots, const float corner[3], const int32_t dims[3], float *d_dxyz,
                                  float *d_dscale, float *d_dobj, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    HvGeom g;
    if (int rc = make_geom(corner, dims, res, g)) return rc;
    CVB_REQUIRE(n >= 0 && num_rots > 0 && num_rots <= 4096, CVB200_EINVAL, "hv_backward: bad n / num_rots");
    if (n == 0) return 0;
    CVB_REQUIRE(d_grad_grid && d_points && d_xyz && d_scale && d_obj && d_dxyz && d_dscale && d_dobj, CVB200_EINVAL,
                "hv_backward: NULL argument");
    hv_backward_kernel<<<(unsigned)ceil_div(n, kBackwardThreads), kBackwardThreads, 2 * num_rots * sizeof(float),
                         stream>>>(d_grad_grid, d_points, d_xyz, d_scale, d_obj, n, num_rots, g, d_dxyz, d_dscale,
                                   d_dobj);
    CVB_LAUNCH_CHECK("hv_backward_kernel");
    return 0;
}

extern "C" int cvb200_hv_vote_indices(const float *d_points, const float *d_xyz, const float *d_scale, int64_t n,
                                      float res, int32_t num_rots, const float corner[3], const int32_t dims[3],
                                      int32_t *d_vote_idx, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    HvGeom g;
    if (int rc = make_geom(corner, dims, res, g)) return rc;
    CVB_REQUIRE(n >= 0 && num_rots > 0, CVB200_EINVAL, "hv_vote_indices: bad n / num_rots");
    if (n == 0) return 0;
    CVB_REQUIRE(d_points && d_xyz && d_scale && d_vote_idx, CVB200_EINVAL, "hv_vote_indices: NULL argument");
    hv_vote_indices_kernel<<<(unsigned)ceil_div(n * num_rots, 256), 256, 0, stream>>>(d_points, d_xyz, d_scale, n,
                                                                                      num_rots, g, d_vote_idx);
    CVB_LAUNCH_CHECK("hv_vote_indices_kernel");
    return 0;
}

extern "C" int cvb200_hv_theta_table(int32_t num_rots, float *d_cos, float *d_sin, void *stream_) {
    CVB_REQUIRE(num_rots > 0 && d_cos && d_sin, CVB200_EINVAL, "hv_theta_table: bad argument");
    hv_theta_table_kernel<<<(num_rots + 127) / 128, 128, 0, (cudaStream_t)stream_>>>(num_rots, d_cos, d_sin);
    CVB_LAUNCH_CHECK("hv_theta_table_kernel");
    return 0;
}
