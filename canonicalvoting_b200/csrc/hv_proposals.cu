// canonicalvoting_b200/csrc/hv_proposals.cu -- vote-map proposal sampler of the SUN RGB-D variant
// (sunrgbd/brnetcanon.py:104-162, `HoughVotingModule.forward`): the two device passes around torch.multinomial.
//
//   hv_project_y_kernel   hv_map.max(1) and torch.argmax(hv_map, 1) (:120,122) in ONE pass over grid_obj [X,Y,Z]
//                         (the reference reads the grid twice); HBM-bound: 4 G bytes read, 8 X Z bytes written.
//   hv_proposals_kernel   for a batch of sampled (x, z) cells (:133-152): unravel, look up the arg-max height and the
//                         voted scale, world location = cell * res + corner, distance to the nearest vote seed
//                         (torch.cdist + min, :139), rejection (< radius; keep everything when nothing passes,
//                         :142-149), ORDER-PRESERVING compaction appended to the running proposal list -- one
//                         CTA, one launch, no host synchronisation (the reference: ~25 launches, 2 syncs per trial).
#include "common.cuh"

namespace cvb200 {

constexpr int kPyZ = 32, kPyS = 16;       // a block = 32 consecutive z x 16 interleaved y-slices of one x
constexpr int kPyBatch = 8;               // loads in flight per thread (128^3: every load of the grid is issued in one wave)

__global__ void __launch_bounds__(kPyZ * kPyS)
hv_project_y_kernel(const float *__restrict__ grid, int X, int Y, int Z, float *__restrict__ out_max, int *__restrict__ out_arg) {
    __shared__ float s_v[kPyS][kPyZ];
    __shared__ int s_i[kPyS][kPyZ];
    const int zl = threadIdx.x & (kPyZ - 1), ys = threadIdx.x / kPyZ;
    const int x = blockIdx.y, z = blockIdx.x * kPyZ + zl;
    float best = -INFINITY;
    int arg = 0x7fffffff;
    if (z < Z) {
        const float *col = grid + (size_t)x * Y * Z + z;
        for (int y0 = ys; y0 < Y; y0 += kPyS * kPyBatch) {
            float v[kPyBatch];
#pragma unroll
            for (int b = 0; b < kPyBatch; b++) {
                const int y = y0 + b * kPyS;
                v[b] = y < Y ? __ldg(col + (size_t)y * Z) : 0.f;
            }
#pragma unroll
            for (int b = 0; b < kPyBatch; b++) {
                const int y = y0 + b * kPyS;
                if (y < Y && (v[b] > best || arg == 0x7fffffff)) { best = v[b]; arg = y; }   // strict >: the first maximum wins (torch.argmax)
            }
        }
    }
    s_v[ys][zl] = best;
    s_i[ys][zl] = arg;
    __syncthreads();
    if (ys == 0 && z < Z) {
#pragma unroll
        for (int s = 1; s < kPyS; s++) {
            const float v = s_v[s][zl];
            const int i = s_i[s][zl];
            if (i != 0x7fffffff && (v > best || (v == best && i < arg))) { best = v; arg = i; }
        }
        out_max[(size_t)x * Z + z] = best;
        out_arg[(size_t)x * Z + z] = arg;
    }
}

constexpr int kPrThreads = 1024, kPrSeedChunk = 1024;

// one CTA; sample i -> thread i % 1024, processed in rounds of 1024 so that the compaction keeps the sample order
__global__ void __launch_bounds__(kPrThreads)
hv_proposals_kernel(const long long *__restrict__ samples, int n, const int *__restrict__ arg_y, const float *__restrict__ grid_scale,
                    int X, int Y, int Z, float res, float cx, float cy, float cz, const float *__restrict__ seeds, int n_seeds,
                    float radius, int max_out, float *__restrict__ out_loc, float *__restrict__ out_scale, int *__restrict__ count,
                    unsigned char *__restrict__ keep_scratch) {
    __shared__ float s_seed[kPrSeedChunk * 3];
    __shared__ int s_warp[32];
    __shared__ int s_total, s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rounds = (n + kPrThreads - 1) / kPrThreads;
    // pass 1: nearest-seed test per sample; the flags go to a scratch byte per sample, their sum decides the "nothing
    // passed -> keep all" rule of the trial (:142-145)
    int kept_local = 0;
    for (int r = 0; r < rounds; r++) {
        const int i = r * kPrThreads + tid;
        float lx = 0.f, ly = 0.f, lz = 0.f;
        if (i < n) {
            const long long s = samples[i];
            const int ix = (int)(s / Z), iz = (int)(s - (long long)ix * Z);
            const int iy = arg_y[(size_t)ix * Z + iz];
            lx = __fadd_rn(__fmul_rn((float)ix, res), cx);
            ly = __fadd_rn(__fmul_rn((float)iy, res), cy);
            lz = __fadd_rn(__fmul_rn((float)iz, res), cz);
        }
        float best = INFINITY;                       // squared distance to the nearest seed
        for (int c0 = 0; c0 < n_seeds; c0 += kPrSeedChunk) {
            const int m = min(kPrSeedChunk, n_seeds - c0);
            __syncthreads();
            for (int e = tid; e < 3 * m; e += kPrThreads) s_seed[e] = __ldg(seeds + (size_t)3 * c0 + e);
            __syncthreads();
            if (i < n) {
                for (int j = 0; j < m; j++) {
                    const float dx = __fsub_rn(lx, s_seed[3 * j]), dy = __fsub_rn(ly, s_seed[3 * j + 1]), dz = __fsub_rn(lz, s_seed[3 * j + 2]);
                    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                    best = fminf(best, d2);
                }
            }
        }
        if (i < n) {
            const int k = __fsqrt_rn(best) < radius ? 1 : 0;
            keep_scratch[i] = (unsigned char)k;
            kept_local += k;
        }
    }
    // block sum of the flags
    int v = kept_local;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int w = 0; w < kPrThreads / 32; w++) t += s_warp[w];
        s_total = t;
        s_base = *count;
    }
    __syncthreads();
    const bool keep_all = s_total == 0;
    // pass 2: ordered compaction, round by round
    for (int r = 0; r < rounds; r++) {
        const int i = r * kPrThreads + tid;
        const int k = (i < n && (keep_all || keep_scratch[i])) ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, k);
        __syncthreads();                              // s_warp / s_base of the previous round have been consumed
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int before = 0;
        for (int w = 0; w < warp; w++) before += s_warp[w];
        const int pos = s_base + before + __popc(bal & ((1u << lane) - 1u));
        if (k && pos < max_out) {
            const long long s = samples[i];
            const int ix = (int)(s / Z), iz = (int)(s - (long long)ix * Z);
            const int iy = arg_y[(size_t)ix * Z + iz];
            out_loc[3 * pos] = __fadd_rn(__fmul_rn((float)ix, res), cx);
            out_loc[3 * pos + 1] = __fadd_rn(__fmul_rn((float)iy, res), cy);
            out_loc[3 * pos + 2] = __fadd_rn(__fmul_rn((float)iz, res), cz);
            const float *sc = grid_scale + (((size_t)ix * Y + iy) * Z + iz) * 3;
            out_scale[3 * pos] = __ldg(sc);
            out_scale[3 * pos + 1] = __ldg(sc + 1);
            out_scale[3 * pos + 2] = __ldg(sc + 2);
        }
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < kPrThreads / 32; w++) t += s_warp[w];
            s_base += t;
        }
    }
    __syncthreads();
    if (tid == 0) *count = s_base;                    // may exceed max_out: the caller stops when count >= num_proposal
}

}  // namespace cvb200

using namespace cvb200;

extern "C" int cvb200_hv_project_y(const float *d_grid_obj, const int32_t dims[3], float *d_max, int32_t *d_arg, void *stream) {
    CVB_REQUIRE(d_grid_obj && dims && d_max && d_arg, CVB200_EINVAL, "hv_project_y: NULL argument");
    CVB_REQUIRE(dims[0] > 0 && dims[1] > 0 && dims[2] > 0 && dims[0] <= 65535, CVB200_EINVAL, "hv_project_y: bad dims (%d, %d, %d)", dims[0],
                dims[1], dims[2]);
    dim3 grid((unsigned)ceil_div(dims[2], kPyZ), (unsigned)dims[0]);
    hv_project_y_kernel<<<grid, kPyZ * kPyS, 0, (cudaStream_t)stream>>>(d_grid_obj, dims[0], dims[1], dims[2], d_max, d_arg);
    CVB_LAUNCH_CHECK("hv_project_y_kernel");
    return 0;
}

extern "C" size_t cvb200_hv_proposals_work_bytes(int32_t n_samples) { return n_samples > 0 ? (size_t)n_samples : 1; }

extern "C" int cvb200_hv_proposals(const int64_t *d_samples, int32_t n_samples, const int32_t *d_arg, const float *d_grid_scale,
                                   const int32_t dims[3], float res, const float corner[3], const float *d_seeds, int32_t n_seeds,
                                   float radius, int32_t max_out, float *d_loc, float *d_scale, int32_t *d_count, void *d_work,
                                   size_t work_bytes, void *stream) {
    CVB_REQUIRE(d_samples && d_arg && d_grid_scale && dims && corner && d_loc && d_scale && d_count && d_work, CVB200_EINVAL,
                "hv_proposals: NULL argument");
    CVB_REQUIRE(n_samples >= 0 && n_seeds >= 0 && max_out >= 0 && (n_seeds == 0 || d_seeds), CVB200_EINVAL, "hv_proposals: bad sizes");
    CVB_REQUIRE(work_bytes >= cvb200_hv_proposals_work_bytes(n_samples), CVB200_ESCRATCH, "hv_proposals: workspace too small");
    if (n_samples == 0) return 0;
    hv_proposals_kernel<<<1, kPrThreads, 0, (cudaStream_t)stream>>>((const long long *)d_samples, n_samples, d_arg, d_grid_scale, dims[0],
                                                                   dims[1], dims[2], res, corner[0], corner[1], corner[2], d_seeds, n_seeds,
                                                                   radius, max_out, d_loc, d_scale, d_count,
                                                                   reinterpret_cast<unsigned char *>(d_work));
    CVB_LAUNCH_CHECK("hv_proposals_kernel");
    return 0;
}
