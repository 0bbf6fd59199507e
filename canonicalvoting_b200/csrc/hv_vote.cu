// canonicalvoting_b200/csrc/hv_vote.cu -- the Hough-voting op on sm_100a.
//
// Replaces the reference's three kernels + host wrapper
//   hv_cuda_forward_kernel / hv_cuda_average_kernel / hv_cuda_forward
//   (houghvoting/src/hv_cuda_kernel.cu:12-97, :100-119, :121-165)
//   hv_cuda_backward_kernel / hv_cuda_backward (:168-261, :265-302)
// with a different decomposition (see DESIGN.md "vote op" and the "forward" section below):
// one work item per (point, theta), the six channels of a voxel accumulated in one 32-byte workspace
// sector, four lanes per vote each issuing one 16-byte vector reduction per (x, y) corner pair (4 warp
// instructions per 8 votes instead of 48 scalar atomics per vote), and a write-out pass that
// normalises, writes every output once and leaves the workspace zero again.
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
// (Formulations without float atomics -- a counting sort + per-voxel gather, round 1; a counting sort by 8x8x8 tile + one
// CTA per tile accumulating in shared memory and writing every output once, round 2 -- and a variant of this one with
// per-block "touched" flags were built, pass the parity tests and are slower: profiles/exp_*, DESIGN.md "what was tried".)
constexpr int kScatterThreads = 256;
constexpr int kPtsPerBlock = 64;   // power of two >= 32: a warp = 32 consecutive points, one theta

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d));
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

    // A warp evaluates 32 votes (32 consecutive points, one theta: one vote per lane), then adds them to the workspace eight
    // votes at a time, FOUR LANES PER VOTE: lanes (4 v + q), q = 2 d + h, cover the two z-adjacent corners d = 0, 1 of one
    // (x, y) corner pair with one 16-byte reduction per lane -- 64 contiguous bytes per vote and instruction, 4 instructions per
    // 8 votes instead of 16 per vote.  (Measured, tools/probes/red_probe.cu: the L2 reduction rate is per instruction and per
    // 64-byte segment, not per byte: 61 -> 30 us for the C2 vote count.)
    const int lane = threadIdx.x & 31;
    const int q = lane & 3, d = q >> 1, h = q & 1;
    const int items = kPtsPerBlock * num_rots;          // a multiple of 32: every lane of a warp runs the same trip count
    const int64_t YZ = (int64_t)g.Y * g.Z;
    for (int j = threadIdx.x; j < items; j += kScatterThreads) {
        const int t = j & (kPtsPerBlock - 1);
        const int i = j / kPtsPerBlock;  // warp-uniform
        const float cs = s_cos[i], sn = s_sin[i];
        bool ok = t < npts;
        float gx = 0.f, gy = 0.f, gz = 0.f;
        if (ok) {
            gy = s_x[3 * t + 1];
            vote_center(s_p[3 * t], s_p[3 * t + 2], s_x[3 * t], s_x[3 * t + 2], cs, sn, g, gx, gz);
            ok = vote_in_bounds(gx, gy, gz, g);
        }
        const int fx = (int)gx, fy = (int)gy, fz = (int)gz;                          // make_int3 (:45)
        const float rx = gx - floorf(gx), ry = gy - floorf(gy), rz = gz - floorf(gz);  // fracf (:47)
        const int v0 = (int)((int64_t)fx * YZ + (int64_t)fy * g.Z + fz);               // < 2^31 (make_geom)
        const float objness = ok ? s_o[t] : 0.f;
        const float s0 = ok ? s_s[3 * t] : 0.f, s1 = ok ? s_s[3 * t + 1] : 0.f, s2 = ok ? s_s[3 * t + 2] : 0.f;
        const unsigned mask = __ballot_sync(0xffffffffu, ok);
        const int nv = __popc(mask);
        for (int b0 = 0; b0 < nv; b0 += 8) {
            const int k = b0 + (lane >> 2);                        // which in-bounds vote of the warp this lane works on
            const bool act = k < nv;
            const int src = act ? __fns(mask, 0, k + 1) : 0;       // the lane that holds it
            const int v = __shfl_sync(0xffffffffu, v0, src);
            const float wx1 = __shfl_sync(0xffffffffu, rx, src), wy1 = __shfl_sync(0xffffffffu, ry, src), wz1 = __shfl_sync(0xffffffffu, rz, src);
            const float ob = __shfl_sync(0xffffffffu, objness, src);
            const float c0 = __shfl_sync(0xffffffffu, s0, src), c1 = __shfl_sync(0xffffffffu, s1, src), c2 = __shfl_sync(0xffffffffu, s2, src);
            if (!act) continue;
            const float wz = d ? wz1 : 1.f - wz1;
            float *base = work + 8 * ((int64_t)v + d) + 4 * h;
#pragma unroll
            for (int ab = 0; ab < 4; ab++) {
                const int a = ab >> 1, b = ab & 1;
                // ((wx*wy)*wz)*objness, the reference's association (:52-59)
                const float w = __fmul_rn(__fmul_rn(__fmul_rn(a ? wx1 : 1.f - wx1, b ? wy1 : 1.f - wy1), wz), ob);
                float *sec = base + 8 * ((int64_t)a * YZ + (int64_t)b * g.Z);
                if (h == 0) red_add_v4(sec, w, __fmul_rn(w, cs), __fmul_rn(w, sn), __fmul_rn(w, c0));
                else red_add_v4(sec, __fmul_rn(w, c1), __fmul_rn(w, c2), 0.f, 0.f);
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

extern "C" size_t cvb200_hv_forward_work_bytes(const int32_t dims[3]) {
    if (!dims || dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) return 0;
    return (size_t)dims[0] * dims[1] * dims[2] * 8 * sizeof(float);
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
    const size_t need = cvb200_hv_forward_work_bytes(dims);
    CVB_REQUIRE(work_bytes >= need, CVB200_ESCRATCH, "hv_forward: workspace %zu < %zu bytes", work_bytes, need);
    CVB_REQUIRE(((reinterpret_cast<uintptr_t>(d_work) | reinterpret_cast<uintptr_t>(d_grid_obj) |
                  reinterpret_cast<uintptr_t>(d_grid_rot) | reinterpret_cast<uintptr_t>(d_grid_scale)) & 15) == 0,
                CVB200_EINVAL, "hv_forward: workspace and outputs must be 16-byte aligned");
    const int64_t G = (int64_t)g.X * g.Y * g.Z;
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
