/*
 * oracle/hv_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement ("port") of the
 * reference Hough-voting op.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may call this; the product path never does.
 *
 * Follows /root/reference/houghvoting/src/hv_cuda_kernel.cu:
 *   hv_cuda_forward_kernel   :12-97     -> cvo_vote_forward (scatter part)
 *   hv_cuda_average_kernel   :100-119   -> cvo_average
 *   hv_cuda_forward (host)   :121-165   -> cvo_grid_dims (+ zero-init by caller)
 *   hv_cuda_backward_kernel  :168-261   -> cvo_vote_backward
 * and the six helper_math.h operations the kernel uses (helper_math.h:157-160
 * make_int3 = truncation, :1338-1349 fracf = v - floorf(v), component-wise + - /).
 *
 * Floating-point contract.  The integer voxel indices are the bit-exact part of
 * the op, so the float32 operation ORDER AND FUSION of the reference's sm_100
 * build are restated explicitly (read off the SASS of oracle/_ref/hv_cuda_ref.so;
 * nvcc contracts with -fmad=true):
 *     rot_interval = 6.2831854820251464844f / (float)R             (IEEE div)
 *     theta_i      = (float)i * rot_interval
 *     corr.x = xyz.x*scale.x ; corr.z = xyz.z*scale.z               (rounded)
 *     off.x  = fma(corr.z,  sin, -(corr.x*cos))
 *     off.z  = fma(corr.x, -sin, -(corr.z*cos))
 *     g.x = ((p.x + off.x) - corner.x) / res                       (IEEE div)
 *     g.y = (fma(xyz.y, -scale.y, p.y) - corner.y) / res           (corr.y fused)
 *     g.z = ((p.z + off.z) - corner.z) / res
 * Compile with -ffp-contract=off so that gcc adds no fusion of its own.
 * cos/sin: the reference evaluates CUDA's cosf/sinf on device; glibc's may differ
 * in the last bit, so every entry point takes an optional (cos,sin) table -- the
 * GPU tests pass the table computed on the device, the golden fixtures carry it.
 *
 * Accumulation: the reference uses atomicAdd (hv_cuda_kernel.cu:61-93), i.e. an
 * unspecified float32 summation order.  acc64 != 0 accumulates in double and
 * rounds once (the "true" sum both implementations must be close to); acc64 == 0
 * accumulates in float32 in point-major / theta-minor order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define CVO_TWO_PI_F 6.2831854820251464844f /* 2 * 3.141592654f, hv_cuda_kernel.cu:35 */

int cvo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* hv_cuda_kernel.cu:35-37 ; glibc cosf/sinf (see header note). */
void cvo_theta_table(int num_rots, float *cos_tab, float *sin_tab) {
    const float rot_interval = CVO_TWO_PI_F / (float)num_rots;
    for (int i = 0; i < num_rots; i++) {
        float theta = (float)i * rot_interval;
        cos_tab[i] = cosf(theta);
        sin_tab[i] = sinf(theta);
    }
}

/* hv_cuda_kernel.cu:129-134,151: corner = min(points,0); dims = int((max-min)/res)+1,
 * all in float32, int() truncates.  Returns 0, or -1 for N == 0. */
int cvo_grid_dims(const float *points, int64_t n, float res, float corner[3], int32_t dims[3]) {
    if (n <= 0) return -1;
    float mn[3], mx[3];
    for (int k = 0; k < 3; k++) mn[k] = mx[k] = points[k];
    for (int64_t c = 1; c < n; c++)
        for (int k = 0; k < 3; k++) {
            float v = points[3 * c + k];
            if (v < mn[k]) mn[k] = v;
            if (v > mx[k]) mx[k] = v;
        }
    for (int k = 0; k < 3; k++) {
        float diff = (mx[k] - mn[k]) / res;
        corner[k] = mn[k];
        dims[k] = (int32_t)diff + 1;
    }
    return 0;
}

/* One vote: centre hypothesis in grid units (hv_cuda_kernel.cu:37-40 as compiled).
 * Returns 0 when the vote is dropped by the bounds test (:41-44). */
static inline int cvo_center(const float *p, const float *xyz, const float *sc, float cs, float sn,
                             const float corner[3], float res, const int32_t dims[3], float g[3]) {
    const float corr_x = xyz[0] * sc[0];
    const float corr_z = xyz[2] * sc[2];
    const float off_x = fmaf(corr_z, sn, -(corr_x * cs));
    const float off_z = fmaf(corr_x, -sn, -(corr_z * cs));
    g[0] = ((p[0] + off_x) - corner[0]) / res;
    g[1] = (fmaf(xyz[1], -sc[1], p[1]) - corner[1]) / res;
    g[2] = ((p[2] + off_z) - corner[2]) / res;
    if (g[0] < 0.f || g[1] < 0.f || g[2] < 0.f) return 0;
    if (g[0] >= (float)(dims[0] - 1) || g[1] >= (float)(dims[1] - 1) || g[2] >= (float)(dims[2] - 1)) return 0;
    /* NaN: every comparison above is false, the reference goes on and indexes with
     * int(NaN); undefined there, dropped here. */
    if (!(g[0] == g[0]) || !(g[1] == g[1]) || !(g[2] == g[2])) return 0;
    return 1;
}

/* Scatter (hv_cuda_kernel.cu:25-96).  Grids must be zero-initialised by the caller
 * (:132-134).  vote_idx (optional, [N,R,3] int32) receives the floor voxel of every
 * vote, or -1,-1,-1 for dropped votes: the bit-exact integer part of the op.
 * threads <= 1: serial, deterministic.  threads > 1: OpenMP over points with atomic
 * float adds (timing baseline; acc64 is ignored). Returns the number of kept votes. */
int64_t cvo_vote_forward(const float *points, const float *xyz, const float *scale, const float *obj,
                         int64_t n, float res, int num_rots, const float *cos_tab, const float *sin_tab,
                         const float corner[3], const int32_t dims[3],
                         float *grid_obj, float *grid_rot, float *grid_scale,
                         int32_t *vote_idx, int acc64, int threads) {
    float *ct = (float *)malloc(sizeof(float) * (size_t)(num_rots > 0 ? num_rots : 1));
    float *st = (float *)malloc(sizeof(float) * (size_t)(num_rots > 0 ? num_rots : 1));
    if (cos_tab && sin_tab) {
        memcpy(ct, cos_tab, sizeof(float) * (size_t)num_rots);
        memcpy(st, sin_tab, sizeof(float) * (size_t)num_rots);
    } else {
        cvo_theta_table(num_rots, ct, st);
    }
    const int64_t Y = dims[1], Z = dims[2];
    const int64_t G = (int64_t)dims[0] * Y * Z;
    int64_t kept = 0;

    if (threads <= 1) {
        double *acc = NULL;
        if (acc64) acc = (double *)calloc((size_t)G * 6, sizeof(double));
        for (int64_t c = 0; c < n; c++) {
            const float objness = obj[c];
            for (int i = 0; i < num_rots; i++) {
                float g[3];
                int ok = cvo_center(points + 3 * c, xyz + 3 * c, scale + 3 * c, ct[i], st[i], corner, res, dims, g);
                if (vote_idx) {
                    int32_t *vi = vote_idx + ((size_t)c * num_rots + i) * 3;
                    vi[0] = ok ? (int32_t)g[0] : -1;
                    vi[1] = ok ? (int32_t)g[1] : -1;
                    vi[2] = ok ? (int32_t)g[2] : -1;
                }
                if (!ok) continue;
                kept++;
                const int fx = (int)g[0], fy = (int)g[1], fz = (int)g[2];
                const float rx = g[0] - floorf(g[0]), ry = g[1] - floorf(g[1]), rz = g[2] - floorf(g[2]);
                const float wx[2] = {1.f - rx, rx}, wy[2] = {1.f - ry, ry}, wz[2] = {1.f - rz, rz};
                const float ch[5] = {ct[i], st[i], scale[3 * c], scale[3 * c + 1], scale[3 * c + 2]};
                for (int a = 0; a < 2; a++)
                    for (int b = 0; b < 2; b++)
                        for (int d = 0; d < 2; d++) {
                            const float w = wx[a] * wy[b] * wz[d] * objness; /* :52-59 */
                            const int64_t v = ((int64_t)(fx + a) * Y + (fy + b)) * Z + (fz + d);
                            if (acc) {
                                acc[v * 6] += (double)w;
                                for (int j = 0; j < 5; j++) acc[v * 6 + 1 + j] += (double)(w * ch[j]);
                            } else {
                                grid_obj[v] += w;
                                grid_rot[v * 2] += w * ch[0];
                                grid_rot[v * 2 + 1] += w * ch[1];
                                grid_scale[v * 3] += w * ch[2];
                                grid_scale[v * 3 + 1] += w * ch[3];
                                grid_scale[v * 3 + 2] += w * ch[4];
                            }
                        }
            }
        }
        if (acc) {
            for (int64_t v = 0; v < G; v++) {
                grid_obj[v] = (float)acc[v * 6];
                grid_rot[v * 2] = (float)acc[v * 6 + 1];
                grid_rot[v * 2 + 1] = (float)acc[v * 6 + 2];
                grid_scale[v * 3] = (float)acc[v * 6 + 3];
                grid_scale[v * 3 + 1] = (float)acc[v * 6 + 4];
                grid_scale[v * 3 + 2] = (float)acc[v * 6 + 5];
            }
            free(acc);
        }
    } else {
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads) schedule(static) reduction(+ : kept)
#endif
        for (int64_t c = 0; c < n; c++) {
            const float objness = obj[c];
            for (int i = 0; i < num_rots; i++) {
                float g[3];
                if (!cvo_center(points + 3 * c, xyz + 3 * c, scale + 3 * c, ct[i], st[i], corner, res, dims, g)) continue;
                kept++;
                const int fx = (int)g[0], fy = (int)g[1], fz = (int)g[2];
                const float rx = g[0] - floorf(g[0]), ry = g[1] - floorf(g[1]), rz = g[2] - floorf(g[2]);
                const float wx[2] = {1.f - rx, rx}, wy[2] = {1.f - ry, ry}, wz[2] = {1.f - rz, rz};
                const float ch[5] = {ct[i], st[i], scale[3 * c], scale[3 * c + 1], scale[3 * c + 2]};
                for (int a = 0; a < 2; a++)
                    for (int b = 0; b < 2; b++)
                        for (int d = 0; d < 2; d++) {
                            const float w = wx[a] * wy[b] * wz[d] * objness;
                            const int64_t v = ((int64_t)(fx + a) * Y + (fy + b)) * Z + (fz + d);
                            float *po = grid_obj + v, *pr = grid_rot + v * 2, *ps = grid_scale + v * 3;
#ifdef _OPENMP
#pragma omp atomic
#endif
                            po[0] += w;
                            for (int j = 0; j < 2; j++) {
                                const float t = w * ch[j];
#ifdef _OPENMP
#pragma omp atomic
#endif
                                pr[j] += t;
                            }
                            for (int j = 0; j < 3; j++) {
                                const float t = w * ch[2 + j];
#ifdef _OPENMP
#pragma omp atomic
#endif
                                ps[j] += t;
                            }
                        }
            }
        }
    }
    free(ct);
    free(st);
    return kept;
}

/* hv_cuda_average_kernel (hv_cuda_kernel.cu:100-119): `x /= w + 1e-7` with a double
 * literal: float / (double)(w + 1e-7) evaluated in double, rounded to float. */
void cvo_average(int64_t g, const float *grid_obj, float *grid_rot, float *grid_scale, int threads) {
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads > 1 ? threads : 1) schedule(static)
#endif
    for (int64_t v = 0; v < g; v++) {
        const double d = (double)grid_obj[v] + 1e-7;
        for (int j = 0; j < 2; j++) grid_rot[v * 2 + j] = (float)((double)grid_rot[v * 2 + j] / d);
        for (int j = 0; j < 3; j++) grid_scale[v * 3 + j] = (float)((double)grid_scale[v * 3 + j] / d);
    }
}

/* hv_cuda_backward_kernel (hv_cuda_kernel.cu:183-259).  grad_grid = dL/dgrid_obj only;
 * the 1/res chain-rule factor is (faithfully) absent.  Outputs are overwritten. */
void cvo_vote_backward(const float *grad_grid, const float *points, const float *xyz, const float *scale,
                       const float *obj, int64_t n, float res, int num_rots, const float *cos_tab,
                       const float *sin_tab, const float corner[3], const int32_t dims[3],
                       float *d_xyz, float *d_scale, float *d_obj, int threads) {
    float *ct = (float *)malloc(sizeof(float) * (size_t)(num_rots > 0 ? num_rots : 1));
    float *st = (float *)malloc(sizeof(float) * (size_t)(num_rots > 0 ? num_rots : 1));
    if (cos_tab && sin_tab) {
        memcpy(ct, cos_tab, sizeof(float) * (size_t)num_rots);
        memcpy(st, sin_tab, sizeof(float) * (size_t)num_rots);
    } else {
        cvo_theta_table(num_rots, ct, st);
    }
    const int64_t Y = dims[1], Z = dims[2];
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads > 1 ? threads : 1) schedule(static)
#endif
    for (int64_t c = 0; c < n; c++) {
        const float objness = obj[c];
        float dobj = 0.f, dx[3] = {0.f, 0.f, 0.f}, ds[3] = {0.f, 0.f, 0.f};
        for (int i = 0; i < num_rots; i++) {
            float g[3];
            if (!cvo_center(points + 3 * c, xyz + 3 * c, scale + 3 * c, ct[i], st[i], corner, res, dims, g)) continue;
            const int fx = (int)g[0], fy = (int)g[1], fz = (int)g[2];
            const float rx = g[0] - floorf(g[0]), ry = g[1] - floorf(g[1]), rz = g[2] - floorf(g[2]);
            const float wx[2] = {1.f - rx, rx}, wy[2] = {1.f - ry, ry}, wz[2] = {1.f - rz, rz};
            float gg[2][2][2];
            for (int a = 0; a < 2; a++)
                for (int b = 0; b < 2; b++)
                    for (int d = 0; d < 2; d++)
                        gg[a][b][d] = grad_grid[((int64_t)(fx + a) * Y + (fy + b)) * Z + (fz + d)];
            float gx = 0.f, gy = 0.f, gz = 0.f;
            for (int a = 0; a < 2; a++)
                for (int b = 0; b < 2; b++)
                    for (int d = 0; d < 2; d++) {
                        dobj += gg[a][b][d] * wx[a] * wy[b] * wz[d];                 /* :210-217 */
                        gx += (a ? 1.f : -1.f) * gg[a][b][d] * wy[b] * wz[d];        /* :219-227 */
                        gy += (b ? 1.f : -1.f) * gg[a][b][d] * wx[a] * wz[d];        /* :228-235 */
                        gz += (d ? 1.f : -1.f) * gg[a][b][d] * wx[a] * wy[b];        /* :236-243 */
                    }
            gx *= objness; gy *= objness; gz *= objness;
            const float dcx = -ct[i] * gx - st[i] * gz;                              /* :249-250 */
            const float dcy = -gy;
            const float dcz = st[i] * gx - ct[i] * gz;
            dx[0] += dcx * scale[3 * c];     dx[1] += dcy * scale[3 * c + 1]; dx[2] += dcz * scale[3 * c + 2];
            ds[0] += dcx * xyz[3 * c];       ds[1] += dcy * xyz[3 * c + 1];   ds[2] += dcz * xyz[3 * c + 2];
        }
        d_obj[c] = dobj;
        for (int k = 0; k < 3; k++) { d_xyz[3 * c + k] = dx[k]; d_scale[3 * c + k] = ds[k]; }
    }
    free(ct);
    free(st);
}
