// canonicalvoting_b200/csrc/bp_loop.cu -- the candidate loop with the LCC-aware back-projection check.
//
// Replaces the Python `while True:` loop of the reference scripts (eval_joint.py:204-263; the same loop is
// inlined in train_joint.py:364-424): argmax of grid_obj -> stop below thresh_high -> zero the 5^3
// neighbourhood -> oriented box from grid_rot / grid_scale at the peak -> zero the grid voxels inside the
// box -> inverse-transform ALL points into the box frame -> accept the box if enough confident points
// agree with their predicted local canonical coordinates -> class vote, score, 8 corners.
// The reference runs ~40 micro-kernels and >= 15 blocking device->host reads per iteration.  Here the
// whole loop is ONE persistent kernel with zero host synchronisation:
//   * a per-block maximum table (blocks of 1024 voxels) makes the per-iteration argmax a scan of G/1024
//     entries; zeroing only invalidates a block when its own argmax voxel is zeroed (values only
//     decrease), and just those blocks are re-scanned;
//   * the N-point pass reduces {n_in, n_conf, sum err*p, max p, class histogram} with warp shuffles;
//   * the kernel is one thread-block CLUSTER (8 CTAs on 8 SMs): every CTA derives the same candidate from the block maxima,
//     takes its share of the box voxels, the points and the dirty blocks, publishes its partial statistics in shared memory,
//     and after a cluster barrier every CTA sums all partials through distributed shared memory (same order everywhere, so
//     all CTAs take the same accept / reject decision without a broadcast).  Two cluster barriers per iteration.  A single
//     CTA spent 56 us per iteration on the 200k points of config C5 (10.6 ms per scene, more than the U-Net).
// Arithmetic order is spelled out with non-contractable intrinsics and mirrored by
// oracle/candidate_loop.py::loop_numpy, so the integer decisions can be compared bit-exactly.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace cvb200 {

constexpr int kBpBlockVox = 1024;    // voxels per block-maximum entry
constexpr int kBpThreads = 1024;     // threads per CTA of the loop kernel
constexpr int kBpCluster = 8;        // the loop kernel is ONE thread-block cluster: the point pass, the box zeroing and the
                                     // re-scans are split over its CTAs, statistics are combined through distributed shared memory
constexpr int kBpMaxClasses = 32;

struct BpGeom {
    float cx, cy, cz, res;
    int X, Y, Z;
};

struct BpPart {              // per-CTA partial statistics of one iteration, read by the whole cluster
    int nin, nconf;
    double err;
    float maxp;
    int hist[kBpMaxClasses];
};

struct BpCand {              // per-iteration state, computed by thread 0 (eval_joint.py:205-223)
    int done;
    int arg;                 // flat index of the peak
    int c[3];
    float world[3];          // cand_world = corner + res * cand            (:206)
    float cs, sn;            // cos / sin of rot = atan2(rot_vec[1], rot_vec[0])  (:213-215)
    float sc[3];             // scale_full                                  (:216)
    float bbox[8][3];        // Rm @ diag(scale) @ bbox_raw                 (:219)
    int lo[3], hi[3];        // clamped bounding volume (inclusive)         (:220-223)
    int box_ok;
};

// first-maximum ordering of torch.argmax: larger value wins, then the smaller flat index
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

__device__ __forceinline__ void warp_argmax(float &v, int &i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (better(ov, oi, v, i)) { v = ov; i = oi; }
    }
}

// (max, first index) of every block of kBpBlockVox voxels
__global__ void __launch_bounds__(256)
bp_blockmax_kernel(const float *__restrict__ grid_obj, int64_t G, float *__restrict__ blockmax, int *__restrict__ blockarg) {
    __shared__ float s_v[8];
    __shared__ int s_i[8];
    const int64_t b0 = (int64_t)blockIdx.x * kBpBlockVox;
    float v = -INFINITY;
    int idx = 0x7fffffff;
    for (int k = threadIdx.x; k < kBpBlockVox; k += 256) {
        const int64_t g = b0 + k;
        if (g < G) {
            const float x = __ldg(grid_obj + g);
            if (better(x, (int)g, v, idx)) { v = x; idx = (int)g; }
        }
    }
    warp_argmax(v, idx);
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = v; s_i[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x < 32) {
        v = threadIdx.x < 8 ? s_v[threadIdx.x] : -INFINITY;
        idx = threadIdx.x < 8 ? s_i[threadIdx.x] : 0x7fffffff;
        warp_argmax(v, idx);
        if (threadIdx.x == 0) { blockmax[blockIdx.x] = v; blockarg[blockIdx.x] = idx; }
    }
}

// inverse transform into the box frame: q = ((d) @ Rm) / scale with Rm = [[c,0,-s],[0,1,0],[s,0,c]]  (:225,:231)
__device__ __forceinline__ bool in_unit_box(float dx, float dy, float dz, const BpCand &cd, float &qx, float &qy, float &qz) {
    qx = __fdiv_rn(__fadd_rn(__fmul_rn(dx, cd.cs), __fmul_rn(dz, cd.sn)), cd.sc[0]);
    qy = __fdiv_rn(dy, cd.sc[1]);
    qz = __fdiv_rn(__fadd_rn(__fmul_rn(dx, -cd.sn), __fmul_rn(dz, cd.cs)), cd.sc[2]);
    return -1.f < qx && qx < 1.f && -1.f < qy && qy < 1.f && -1.f < qz && qz < 1.f;
}

__global__ void __launch_bounds__(kBpThreads, 1)
bp_loop_kernel(float *__restrict__ grid_obj, const float *__restrict__ grid_rot, const float *__restrict__ grid_scale,
               BpGeom g, const float *__restrict__ points, const float *__restrict__ xyz, const float *__restrict__ prob,
               const int64_t *__restrict__ cls, int64_t n, cvb200_bp_params prm, float *__restrict__ blockmax,
               int *__restrict__ blockarg, int *__restrict__ dirty, int *__restrict__ ndirty /*[2], zero*/, int nb, float *__restrict__ out_boxes,
               float *__restrict__ out_scores, int32_t *__restrict__ out_classes, int32_t *__restrict__ out_counts,
               int32_t *__restrict__ out_trace) {
    __shared__ BpCand cd;
    __shared__ float s_v[32];
    __shared__ int s_i[32];
    __shared__ int s_nin[32], s_nconf[32];
    __shared__ double s_err[32];
    __shared__ float s_maxp[32];
    __shared__ BpPart s_part;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), nrank = (int)cluster.num_blocks();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t ctid = (int64_t)rank * kBpThreads + tid, cthreads = (int64_t)nrank * kBpThreads;
    const int cwarp = rank * (kBpThreads / 32) + wid, cwarps = nrank * (kBpThreads / 32);
    const int64_t G = (int64_t)g.X * g.Y * g.Z;
    const int YZ = g.Y * g.Z;
    int n_boxes = 0, iters = 0;

    while (true) {
        // ---- P1: argmax over the block maxima (first maximum)                          (:205)
        float v = -INFINITY;
        int idx = 0x7fffffff;
        for (int b = tid; b < nb; b += kBpThreads) {
            const float x = blockmax[b];
            const int a = blockarg[b];
            if (better(x, a, v, idx)) { v = x; idx = a; }
        }
        warp_argmax(v, idx);
        if (lane == 0) { s_v[wid] = v; s_i[wid] = idx; }
        if (tid < kBpMaxClasses) s_part.hist[tid] = 0;
        int *const nd_now = ndirty + (iters & 1);
        if (rank == 0 && tid == 0) ndirty[(iters + 1) & 1] = 0;      // the other counter: last read before the previous barrier
        __syncthreads();
        if (wid == 0) {
            v = s_v[lane];
            idx = s_i[lane];
            warp_argmax(v, idx);
            if (lane == 0) {
                // ---- P2: candidate parameters                                          (:206-223)
                cd.done = !(v >= prm.thresh_high) || iters >= prm.max_iters || n_boxes >= prm.max_boxes;   // `< thresh_high: break` (:208)
                if (!cd.done) {
                    cd.arg = idx;
                    cd.c[0] = idx / YZ;
                    cd.c[1] = (idx - cd.c[0] * YZ) / g.Z;
                    cd.c[2] = idx - cd.c[0] * YZ - cd.c[1] * g.Z;
                    cd.world[0] = __fadd_rn(g.cx, __fmul_rn(g.res, (float)cd.c[0]));
                    cd.world[1] = __fadd_rn(g.cy, __fmul_rn(g.res, (float)cd.c[1]));
                    cd.world[2] = __fadd_rn(g.cz, __fmul_rn(g.res, (float)cd.c[2]));
                    const float rot = atan2f(grid_rot[2 * (int64_t)idx + 1], grid_rot[2 * (int64_t)idx]);
                    cd.cs = cosf(rot);
                    cd.sn = sinf(rot);
                    cd.sc[0] = grid_scale[3 * (int64_t)idx];
                    cd.sc[1] = grid_scale[3 * (int64_t)idx + 1];
                    cd.sc[2] = grid_scale[3 * (int64_t)idx + 2];
                    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                    for (int k = 0; k < 8; k++) {   // bbox_raw (:203): x = ++--++--, y = ++++----, z = +--++--+
                        const float rx = (k & 2) ? -1.f : 1.f, ry = (k & 4) ? -1.f : 1.f, rz = ((k + 1) & 2) ? -1.f : 1.f;
                        const float ex = __fmul_rn(cd.sc[0], rx), ey = __fmul_rn(cd.sc[1], ry), ez = __fmul_rn(cd.sc[2], rz);
                        cd.bbox[k][0] = __fadd_rn(__fmul_rn(cd.cs, ex), __fmul_rn(-cd.sn, ez));
                        cd.bbox[k][1] = ey;
                        cd.bbox[k][2] = __fadd_rn(__fmul_rn(cd.sn, ex), __fmul_rn(cd.cs, ez));
#pragma unroll
                        for (int d = 0; d < 3; d++) { mn[d] = fminf(mn[d], cd.bbox[k][d]); mx[d] = fmaxf(mx[d], cd.bbox[k][d]); }
                    }
                    const int dims[3] = {g.X, g.Y, g.Z};
                    cd.box_ok = 1;
#pragma unroll
                    for (int d = 0; d < 3; d++) {
                        // (min|max / res).int(): truncation toward zero; clamp the float first so that the cast is defined
                        const float fmn = fminf(fmaxf(__fdiv_rn(mn[d], g.res), -2.0e9f), 2.0e9f);
                        const float fmx = fminf(fmaxf(__fdiv_rn(mx[d], g.res), -2.0e9f), 2.0e9f);
                        const long long bmin = (long long)fmn, bmax = (long long)fmx;   // NaN -> 0
                        if (!(bmax >= bmin)) cd.box_ok = 0;
                        cd.lo[d] = (int)min(max((long long)cd.c[d] + bmin, 0LL), (long long)dims[d] - 1);
                        cd.hi[d] = (int)min(max((long long)cd.c[d] + bmax, 0LL), (long long)dims[d] - 1);
                    }
                }
            }
        }
        __syncthreads();
        if (cd.done) break;

        // ---- P3: zero the elimination neighbourhood (:211) and the voxels inside the box (:225-229,:243)
        auto zero_voxel = [&](int x, int y, int z) {
            const int64_t f = ((int64_t)x * g.Y + y) * g.Z + z;
            grid_obj[f] = 0.f;
            const int b = (int)(f / kBpBlockVox);
            if (atomicCAS(blockarg + b, (int)f, -1) == (int)f) dirty[atomicAdd(nd_now, 1)] = b;
        };
        {
            const int el = prm.elimination, ehi = prm.elim_hi_inclusive ? el + 1 : el, side = el + ehi;
            if (rank == 0 && tid < side * side * side) {
                const int x = cd.c[0] - el + tid / (side * side), y = cd.c[1] - el + (tid / side) % side, z = cd.c[2] - el + tid % side;
                if (x >= 0 && y >= 0 && z >= 0 && x < g.X && y < g.Y && z < g.Z) zero_voxel(x, y, z);
            }
        }
        if (cd.box_ok) {
            const int ex = cd.hi[0] - cd.lo[0] + 1, ey = cd.hi[1] - cd.lo[1] + 1, ez = cd.hi[2] - cd.lo[2] + 1;
            const int64_t vol = (int64_t)ex * ey * ez;
            for (int64_t i = ctid; i < vol; i += cthreads) {
                const int z = cd.lo[2] + (int)(i % ez), y = cd.lo[1] + (int)((i / ez) % ey), x = cd.lo[0] + (int)(i / ((int64_t)ez * ey));
                float qx, qy, qz;
                if (in_unit_box(__fmul_rn((float)(x - cd.c[0]), g.res), __fmul_rn((float)(y - cd.c[1]), g.res),
                                __fmul_rn((float)(z - cd.c[2]), g.res), cd, qx, qy, qz))
                    zero_voxel(x, y, z);
            }
        }

        // ---- P4: all points into the box frame; LCC back-projection statistics          (:231-258)
        int nin = 0, nconf = 0;
        double err = 0.0;
        float maxp = -INFINITY;
        for (int64_t i = ctid; i < n; i += cthreads) {
            const float px = __ldg(points + 3 * i), py = __ldg(points + 3 * i + 1), pz = __ldg(points + 3 * i + 2);
            float qx, qy, qz;
            if (!in_unit_box(__fsub_rn(px, cd.world[0]), __fsub_rn(py, cd.world[1]), __fsub_rn(pz, cd.world[2]), cd, qx, qy, qz))
                continue;
            const float p = __ldg(prob + i);
            nin++;
            maxp = fmaxf(maxp, p);
            if (p > prm.prob_thresh) {   // mask = prob_pred[bbox_mask_world] > 0.3  (:245)
                nconf++;
                const float e0 = __fsub_rn(__ldg(xyz + 3 * i), qx), e1 = __fsub_rn(__ldg(xyz + 3 * i + 1), qy),
                            e2 = __fsub_rn(__ldg(xyz + 3 * i + 2), qz);
                const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(e0, e0), __fmul_rn(e1, e1)), __fmul_rn(e2, e2)));
                err += (double)__fmul_rn(nrm, p);   // ||xyz_pred - lcc|| * prob  (:250)
                const long long c = cls[i];
                if (c >= 0 && c < kBpMaxClasses) atomicAdd(&s_part.hist[c], 1);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            nin += __shfl_xor_sync(0xffffffffu, nin, o);
            nconf += __shfl_xor_sync(0xffffffffu, nconf, o);
            err += __shfl_xor_sync(0xffffffffu, err, o);
            maxp = fmaxf(maxp, __shfl_xor_sync(0xffffffffu, maxp, o));
        }
        if (lane == 0) { s_nin[wid] = nin; s_nconf[wid] = nconf; s_err[wid] = err; s_maxp[wid] = maxp; }
        __syncthreads();
        if (wid == 0) {      // this CTA's partial statistics
            nin = s_nin[lane]; nconf = s_nconf[lane]; err = s_err[lane]; maxp = s_maxp[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                nin += __shfl_xor_sync(0xffffffffu, nin, o);
                nconf += __shfl_xor_sync(0xffffffffu, nconf, o);
                err += __shfl_xor_sync(0xffffffffu, err, o);
                maxp = fmaxf(maxp, __shfl_xor_sync(0xffffffffu, maxp, o));
            }
            if (lane == 0) { s_part.nin = nin; s_part.nconf = nconf; s_part.err = err; s_part.maxp = maxp; }
        }
        cluster.sync();      // zeroed voxels, dirty list and every CTA's partials are visible to the cluster

        // ---- P5: re-scan the blocks whose argmax voxel was zeroed (one warp per block, all warps of the cluster)
        const int nd = *(volatile int *)nd_now;
        for (int j = cwarp; j < nd; j += cwarps) {
            const int b = dirty[j];
            const int64_t b0 = (int64_t)b * kBpBlockVox;
            float bv = -INFINITY;
            int bi = 0x7fffffff;
#pragma unroll 4
            for (int k = lane; k < kBpBlockVox; k += 32) {
                const int64_t f = b0 + k;
                if (f < G) {
                    const float x = grid_obj[f];
                    if (better(x, (int)f, bv, bi)) { bv = x; bi = (int)f; }
                }
            }
            warp_argmax(bv, bi);
            if (lane == 0) { blockmax[b] = bv; blockarg[b] = bi; }
        }

        // ---- P6: accept / reject (:246-253), class vote (:255-256), score (:258), corners (:259) -- every CTA sums the
        // partials of all CTAs in rank order (distributed shared memory) and reaches the same decision
        if (wid == 0) {
            nin = 0; nconf = 0; err = 0.0; maxp = -INFINITY;
            int hist = 0;
            for (int r = 0; r < nrank; r++) {
                const BpPart *pr = cluster.map_shared_rank(&s_part, r);
                nin += pr->nin; nconf += pr->nconf; err += pr->err; maxp = fmaxf(maxp, pr->maxp);
                hist += pr->hist[lane];
            }
            // sum(mask) < valid_ratio * sum(in)  is evaluated in float32 by torch's type promotion
            const bool reject = ((float)nconf < __fmul_rn(prm.valid_ratio, (float)nin)) || nin < prm.thresh_low;
            bool accept = false;
            if (!reject) accept = !(err / (double)nconf > (double)prm.err_thresh);
            if (accept) {
                // smallest class id among the most frequent ones (torch.unique is sorted, argmax takes the first)
                int cbest = lane, nbest = hist;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const int oc = __shfl_xor_sync(0xffffffffu, cbest, o), on = __shfl_xor_sync(0xffffffffu, nbest, o);
                    if (on > nbest || (on == nbest && oc < cbest)) { nbest = on; cbest = oc; }
                }
                if (rank == 0 && lane < 24) out_boxes[24 * (int64_t)n_boxes + lane] = __fadd_rn(cd.bbox[lane / 3][lane % 3], cd.world[lane % 3]);
                if (rank == 0 && lane == 0) { out_scores[n_boxes] = maxp; out_classes[n_boxes] = cbest; }
            }
            if (out_trace && rank == 0 && lane == 0 && iters < prm.max_trace) {
                out_trace[4 * iters] = cd.arg; out_trace[4 * iters + 1] = nin; out_trace[4 * iters + 2] = nconf;
                out_trace[4 * iters + 3] = accept ? 1 : 0;
            }
            if (lane == 0) s_i[0] = accept ? 1 : 0;
        }
        __syncthreads();
        n_boxes += s_i[0];
        iters++;
        cluster.sync();      // block maxima of the re-scans are visible; nobody still reads this iteration's partials
    }
    if (rank == 0 && tid == 0) { out_counts[0] = n_boxes; out_counts[1] = iters; }
}

static size_t bp_nb(const int32_t dims[3]) {
    return (size_t)(((int64_t)dims[0] * dims[1] * dims[2] + kBpBlockVox - 1) / kBpBlockVox);
}

}  // namespace cvb200

using namespace cvb200;

extern "C" void cvb200_bp_default_params(cvb200_bp_params *p) {
    if (!p) return;
    p->thresh_high = 60.f;      // eval_joint.py:18
    p->thresh_low = 10;         // :19
    p->valid_ratio = 0.2f;      // :20
    p->elimination = 2;         // :21
    p->elim_hi_inclusive = 1;   // eval_joint.py:211 (`+elimination+1`); eval_separate.py:209 omits the +1
    p->prob_thresh = 0.3f;      // :245
    p->err_thresh = 0.3f;       // :252
    p->max_boxes = 4096;
    p->max_iters = 1 << 30;
    p->max_trace = 0;
}

extern "C" size_t cvb200_bp_work_bytes(const int32_t dims[3]) {
    if (!dims || dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) return 0;
    return bp_nb(dims) * (sizeof(float) + 2 * sizeof(int)) + 256;   // block maxima, arguments, dirty list + the two dirty counters
}

extern "C" int cvb200_back_project(float *d_grid_obj, const float *d_grid_rot, const float *d_grid_scale,
                                   const int32_t dims[3], const float corner[3], float res, const float *d_points,
                                   const float *d_xyz, const float *d_prob, const int64_t *d_class, int64_t n,
                                   const cvb200_bp_params *params, float *d_boxes, float *d_scores,
                                   int32_t *d_classes, int32_t *d_counts, int32_t *d_trace, void *d_work,
                                   size_t work_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVB_REQUIRE(dims && corner && params, CVB200_EINVAL, "back_project: NULL dims/corner/params");
    CVB_REQUIRE(dims[0] > 0 && dims[1] > 0 && dims[2] > 0 && (int64_t)dims[0] * dims[1] * dims[2] < ((int64_t)1 << 31),
                CVB200_EINVAL, "back_project: bad grid dims");
    CVB_REQUIRE(d_grid_obj && d_grid_rot && d_grid_scale && d_boxes && d_scores && d_classes && d_counts && d_work,
                CVB200_EINVAL, "back_project: NULL grid/output/work pointer");
    CVB_REQUIRE(n >= 0 && (n == 0 || (d_points && d_xyz && d_prob && d_class)), CVB200_EINVAL, "back_project: NULL input");
    CVB_REQUIRE(params->thresh_high > 0.f, CVB200_EINVAL,
                "back_project: thresh_high must be > 0 (the reference loop would never terminate)");
    CVB_REQUIRE(params->elimination >= 0 && params->elimination <= 4, CVB200_EINVAL, "back_project: elimination must be in [0,4]");
    CVB_REQUIRE(params->max_boxes > 0 && (params->max_trace == 0 || d_trace), CVB200_EINVAL, "back_project: bad max_boxes / trace");
    CVB_REQUIRE(work_bytes >= cvb200_bp_work_bytes(dims), CVB200_ESCRATCH, "back_project: workspace %zu < %zu bytes",
                work_bytes, cvb200_bp_work_bytes(dims));
    const int nb = (int)bp_nb(dims);
    float *blockmax = (float *)d_work;
    int *blockarg = (int *)(blockmax + nb);
    int *dirty = blockarg + nb;
    BpGeom g;
    g.cx = corner[0]; g.cy = corner[1]; g.cz = corner[2]; g.res = res;
    g.X = dims[0]; g.Y = dims[1]; g.Z = dims[2];
    const int64_t G = (int64_t)g.X * g.Y * g.Z;
    bp_blockmax_kernel<<<nb, 256, 0, stream>>>(d_grid_obj, G, blockmax, blockarg);
    CVB_LAUNCH_CHECK("bp_blockmax_kernel");
    int *ndirty = dirty + nb;
    CVB_CUDA(cudaMemsetAsync(ndirty, 0, 2 * sizeof(int), stream));
    // 16 CTAs per cluster (non-portable size, one GPC) when the device schedules it and the scene is large; 8 otherwise
    static DeviceOnce probed;                    // per device: the attribute and the occupancy answer belong to its context
    static std::atomic<int> can16[256];
    const int dev_id = DeviceOnce::device();
    if (!probed.done()) {
        int max16 = 0;
        if (cudaFuncSetAttribute(bp_loop_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3(16);
            q.blockDim = dim3(kBpThreads);
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = 16; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            q.attrs = qa; q.numAttrs = 1;
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, bp_loop_kernel, &q) == cudaSuccess && nclusters >= 1) max16 = 1;
        }
        (void)cudaGetLastError();
        can16[dev_id].store(max16);
        probed.mark();
    }
    const int csize = (can16[dev_id].load() == 1 && n >= 30000) ? 16 : kBpCluster;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize);
    cfg.blockDim = dim3(kBpThreads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CVB_CUDA(cudaLaunchKernelEx(&cfg, bp_loop_kernel, d_grid_obj, d_grid_rot, d_grid_scale, g, d_points, d_xyz, d_prob, d_class, n, *params,
                                blockmax, blockarg, dirty, ndirty, nb, d_boxes, d_scores, d_classes, d_counts, d_trace));
    return 0;
}
