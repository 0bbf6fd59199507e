// canonicalvoting_b200/csrc/obb_nms.cu -- oriented-box IoU + per-class greedy NMS on the device.
//
// The step right after the candidate loop (eval_joint.py:265-280): for every class the boxes are sorted by score and
// greedily suppressed with `get_iou_obb` (utils/calc_map.py:6-21: xz-rectangle intersection x y-overlap) at 0.3.  The
// reference does this in Python with shapely polygons, O(K^2) calls; here ONE CTA ranks the boxes (class ascending, score
// descending, ties: larger index first -- what a stable argsort + "take the last" gives) and walks the ranking; the IoUs of
// the current pick against all remaining boxes of its class are evaluated in parallel.  Geometry in float64 like
// numpy/shapely: Sutherland-Hodgman clipping of the two convex quadrilaterals.
#include "common.cuh"

namespace cvb200 {

constexpr int kNmsThreads = 256;
constexpr int kNmsMaxBoxes = 2048;

struct Quad { double x[4], z[4]; };

__device__ __forceinline__ double quad_signed_area(const double *x, const double *z, int n) {
    double a = 0.0;
    for (int i = 0; i < n; i++) {
        const int j = i + 1 == n ? 0 : i + 1;
        a += x[i] * z[j] - x[j] * z[i];
    }
    return 0.5 * a;
}

__device__ double quad_intersection_area(const Quad &p, const Quad &q) {
    double sx[10], sz[10], tx[10], tz[10];
    int n = 4;
    for (int i = 0; i < 4; i++) { sx[i] = p.x[i]; sz[i] = p.z[i]; }
    const double sign = quad_signed_area(q.x, q.z, 4) >= 0.0 ? 1.0 : -1.0;
    for (int e = 0; e < 4 && n > 0; e++) {
        const double ax = q.x[e], az = q.z[e], bx = q.x[(e + 1) & 3], bz = q.z[(e + 1) & 3];
        int m = 0;
        for (int i = 0; i < n; i++) {
            const int j = i + 1 == n ? 0 : i + 1;
            const double dp = sign * ((bx - ax) * (sz[i] - az) - (bz - az) * (sx[i] - ax));
            const double dq = sign * ((bx - ax) * (sz[j] - az) - (bz - az) * (sx[j] - ax));
            if (dp >= 0.0) { tx[m] = sx[i]; tz[m] = sz[i]; m++; }
            if ((dp >= 0.0) != (dq >= 0.0)) {
                const double t = dp / (dp - dq);
                tx[m] = sx[i] + t * (sx[j] - sx[i]);
                tz[m] = sz[i] + t * (sz[j] - sz[i]);
                m++;
            }
        }
        n = m;
        for (int i = 0; i < n; i++) { sx[i] = tx[i]; sz[i] = tz[i]; }
    }
    if (n < 3) return 0.0;
    return fabs(quad_signed_area(sx, sz, n));
}

// boxes [K][8][3]: corners 0-3 top face, 4-7 bottom face
__device__ double obb_iou(const float *b1, const float *b2) {
    const double t1 = b1[1], o1 = b1[4 * 3 + 1], t2 = b2[1], o2 = b2[4 * 3 + 1];
    if (!(t1 > o1 && t2 > o2)) return 0.0;
    Quad p, q;
    for (int i = 0; i < 4; i++) {
        p.x[i] = b1[3 * i]; p.z[i] = b1[3 * i + 2];
        q.x[i] = b2[3 * i]; q.z[i] = b2[3 * i + 2];
    }
    const double inter = quad_intersection_area(p, q);
    const double vol = inter * fmax(0.0, fmin(t1, t2) - fmax(o1, o2));
    const double a1 = fabs(quad_signed_area(p.x, p.z, 4)), a2 = fabs(quad_signed_area(q.x, q.z, 4));
    return vol / (a1 * (t1 - o1) + a2 * (t2 - o2) - vol);
}

__global__ void __launch_bounds__(kNmsThreads)
obb_nms_kernel(const float *__restrict__ boxes, const float *__restrict__ scores, const int *__restrict__ classes, int K, int nclasses,
               double thr, int *__restrict__ pick, int *__restrict__ n_pick) {
    __shared__ int order[kNmsMaxBoxes];
    __shared__ unsigned char dead[kNmsMaxBoxes];
    __shared__ int s_count;
    const int tid = threadIdx.x;
    // rank = number of boxes that come before: (class asc, score desc, index desc); boxes of classes outside [0, nclasses) drop out
    for (int i = tid; i < K; i += kNmsThreads) {
        const int ci = classes[i];
        const float si = scores[i];
        int rank = 0;
        for (int j = 0; j < K; j++) {
            const int cj = classes[j];
            const float sj = scores[j];
            const bool valid_j = cj >= 0 && cj < nclasses;
            const bool before = cj < ci || (cj == ci && (sj > si || (sj == si && j > i)));
            rank += (valid_j && before) ? 1 : 0;
        }
        dead[i] = (ci < 0 || ci >= nclasses) ? 1 : 0;
        if (!dead[i]) order[rank] = i;
    }
    if (tid == 0) s_count = 0;
    __syncthreads();
    int n_valid = 0;
    for (int i = 0; i < K; i++) n_valid += dead[i] ? 0 : 1;     // every thread: K is small
    __syncthreads();
    for (int r = 0; r < n_valid; r++) {
        const int i = order[r];
        if (dead[i]) { __syncthreads(); continue; }             // uniform: dead[] is read after a barrier
        if (tid == 0) pick[s_count++] = i;
        const int ci = classes[i];
        for (int r2 = r + 1 + tid; r2 < n_valid; r2 += kNmsThreads) {
            const int j = order[r2];
            if (classes[j] != ci) break;                         // the ranking is class-major
            if (!dead[j] && obb_iou(boxes + (size_t)i * 24, boxes + (size_t)j * 24) > thr) dead[j] = 1;
        }
        __syncthreads();
    }
    if (tid == 0) *n_pick = s_count;
}

__global__ void obb_iou_matrix_kernel(const float *__restrict__ a, int na, const float *__restrict__ b, int nb, double *__restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)na * nb) return;
    out[t] = obb_iou(a + (size_t)(t / nb) * 24, b + (size_t)(t % nb) * 24);
}

}  // namespace cvb200

using namespace cvb200;

extern "C" int cvb200_obb_nms(const float *d_boxes, const float *d_scores, const int32_t *d_classes, int32_t k, int32_t nclasses,
                              double overlap_threshold, int32_t *d_pick, int32_t *d_n_pick, void *stream_) {
    CVB_REQUIRE(k >= 0 && k <= kNmsMaxBoxes && nclasses >= 1, CVB200_EINVAL, "obb_nms: 0 <= k <= %d boxes expected (got %d)", kNmsMaxBoxes, k);
    CVB_REQUIRE(d_n_pick && (k == 0 || (d_boxes && d_scores && d_classes && d_pick)), CVB200_EINVAL, "obb_nms: NULL argument");
    obb_nms_kernel<<<1, kNmsThreads, 0, (cudaStream_t)stream_>>>(d_boxes, d_scores, d_classes, k, nclasses, overlap_threshold, d_pick, d_n_pick);
    CVB_LAUNCH_CHECK("obb_nms_kernel");
    return 0;
}

extern "C" int cvb200_obb_iou_matrix(const float *d_a, int32_t na, const float *d_b, int32_t nb, double *d_out, void *stream_) {
    CVB_REQUIRE(na >= 0 && nb >= 0, CVB200_EINVAL, "obb_iou_matrix: bad sizes");
    if (na == 0 || nb == 0) return 0;
    CVB_REQUIRE(d_a && d_b && d_out, CVB200_EINVAL, "obb_iou_matrix: NULL argument");
    const long long total = (long long)na * nb;
    obb_iou_matrix_kernel<<<(unsigned)ceil_div(total, 128), 128, 0, (cudaStream_t)stream_>>>(d_a, na, d_b, nb, d_out);
    CVB_LAUNCH_CHECK("obb_iou_matrix_kernel");
    return 0;
}
