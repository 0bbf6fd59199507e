/*
 * include/cvb200.h -- C ABI of libcvb200.so, the B200 (sm_100a) implementation of the
 * CanonicalVoting hot path.  Plain pointers and sizes only: no torch / C++ types.
 *
 * Every entry point below is what a binding for the corresponding reference
 * interface would call; the reference interface it replaces is cited as
 * (file:line) relative to the qq456cvb/CanonicalVoting tree.
 *
 * Conventions
 *   - all `d_*` pointers are DEVICE pointers on the current CUDA device, float32 /
 *     int32, row-major contiguous (the reference enforces contiguity with
 *     CHECK_CONTIGUOUS, houghvoting/src/hv_cuda.cpp:26-28);
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream, which
 *     is what the reference launches on, hv_cuda_kernel.cu:143);
 *   - every function returns 0 on success, a positive cudaError_t on a CUDA
 *     failure, or a negative CVB200_E* code on an argument error;
 *     cvb200_last_error() returns a thread-local human-readable message;
 *   - functions are asynchronous with respect to the host unless stated otherwise.
 */
#ifndef CVB200_H_
#define CVB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVB200_ABI_VERSION 14

#define CVB200_EINVAL   (-1) /* bad argument (null pointer, negative size, ...) */
#define CVB200_ESCRATCH (-2) /* workspace too small */
#define CVB200_EEMPTY   (-3) /* empty input where the reference would fail too */

int cvb200_abi_version(void);
const char *cvb200_last_error(void);

/* ------------------------------------------------------------------ vote op ---- */

/* Grid geometry of one scene: corner = min(points, 0), dims = int((max-min)/res)+1
 * evaluated in float32 exactly like the reference host code
 * (houghvoting/src/hv_cuda_kernel.cu:129-134,151).  SYNCHRONOUS: one device
 * reduction + one 48-byte D->H copy + stream sync (the reference does 12 blocking
 * .item() calls for the same information).
 *   d_points   [n,3]
 *   d_work     >= cvb200_hv_grid_dims_work_bytes() bytes of device scratch
 *   h_corner   [3] out (host), h_maxpt [3] out (host, may be NULL), h_dims [3] out (host) */
size_t cvb200_hv_grid_dims_work_bytes(void);
int cvb200_hv_grid_dims(const float *d_points, int64_t n, float res, void *d_work,
                        float *h_corner, float *h_maxpt, int32_t *h_dims, void *stream);

/* Bytes of the device workspace of cvb200_hv_forward for a grid of dims[0..2] voxels:
 * an interleaved accumulator of 8 floats (one 32-byte sector) per voxel.
 * CONTRACT: the workspace must be all-zero when cvb200_hv_forward is entered and is
 * all-zero again when the call's work completes (the write-out pass re-zeroes it), so a caller zero-fills it once after allocation and may
 * then reuse it forever on the same stream.  16-byte alignment required. */
size_t cvb200_hv_forward_work_bytes(const int32_t dims[3]);

/* hv_cuda.forward (houghvoting/src/hv_cuda.cpp:30-45 -> hv_cuda_kernel.cu:121-165):
 * scatter every point's num_rots oriented centre votes into the grid with trilinear
 * weights (hv_cuda_forward_kernel, :12-97) and divide grid_rot / grid_scale by
 * (grid_obj + 1e-7) (hv_cuda_average_kernel, :100-119).
 *   d_points,d_xyz,d_scale [n,3]; d_obj [n]
 *   corner[3], dims[3]     host values (from cvb200_hv_grid_dims or the caller)
 *   d_grid_obj [X,Y,Z], d_grid_rot [X,Y,Z,2], d_grid_scale [X,Y,Z,3]: outputs, every
 *       element is written exactly once (no pre-zeroing needed; the reference needs 3
 *       memsets)
 *   d_work / work_bytes    see cvb200_hv_forward_work_bytes */
int cvb200_hv_forward(const float *d_points, const float *d_xyz, const float *d_scale, const float *d_obj,
                      int64_t n, float res, int32_t num_rots, const float corner[3], const int32_t dims[3],
                      float *d_grid_obj, float *d_grid_rot, float *d_grid_scale,
                      void *d_work, size_t work_bytes, void *stream);

/* hv_cuda.backward (houghvoting/src/hv_cuda.cpp:47-71 -> hv_cuda_kernel.cu:265-302,
 * kernel :168-261): gradient of sum(grad_grid * grid_obj) w.r.t. xyz / scale / obj.
 * Faithful to the reference: only dL/dgrid_obj is consumed, the 1/res factor of the
 * grid coordinate is not applied.  Outputs are fully overwritten.
 *   d_grad_grid [X,Y,Z]; d_dxyz,d_dscale [n,3]; d_dobj [n] */
int cvb200_hv_backward(const float *d_grad_grid, const float *d_points, const float *d_xyz,
                       const float *d_scale, const float *d_obj, int64_t n, float res, int32_t num_rots,
                       const float corner[3], const int32_t dims[3],
                       float *d_dxyz, float *d_dscale, float *d_dobj, void *stream);

/* Verification aid (no reference counterpart): the integer floor voxel of every vote,
 * d_vote_idx [n,num_rots,3] int32, (-1,-1,-1) for votes the bounds test drops
 * (hv_cuda_kernel.cu:41-45).  This is the bit-exact part of the op. */
int cvb200_hv_vote_indices(const float *d_points, const float *d_xyz, const float *d_scale, int64_t n,
                           float res, int32_t num_rots, const float corner[3], const int32_t dims[3],
                           int32_t *d_vote_idx, void *stream);

/* cos/sin of theta_i = i * (2*3.141592654f / num_rots) as the device evaluates them
 * (hv_cuda_kernel.cu:35-38); d_cos,d_sin [num_rots].  Lets a CPU checker share the
 * exact table. */
int cvb200_hv_theta_table(int32_t num_rots, float *d_cos, float *d_sin, void *stream);

/* Vote-map proposal sampler of the SUN RGB-D variant (sunrgbd/brnetcanon.py:104-162, HoughVotingModule.forward).
 * cvb200_hv_project_y: d_max[x*Z+z] = max_y grid_obj[x,y,z], d_arg[x*Z+z] = first y attaining it -- `hv_map.max(1)[0]`
 * and `torch.argmax(hv_map, 1)` (:120,122) in one pass over the grid. */
int cvb200_hv_project_y(const float *d_grid_obj, const int32_t dims[3], float *d_max, int32_t *d_arg, void *stream);
/* One rejection-sampling trial (:133-152).  d_samples int64 [n_samples]: flat (x*Z+z) cells drawn by the caller
 * (torch.multinomial, :133).  Per sample: y = d_arg[cell], location = (x,y,z)*res + corner (:137), scale =
 * grid_scale[x,y,z,:] (:138), distance to the nearest of d_seeds [n_seeds,3] (:139); samples closer than `radius` are
 * kept -- all of them when none is (:142-149) -- and appended IN SAMPLE ORDER to d_loc / d_scale [max_out,3] at row
 * *d_count, rows >= max_out dropped (:154-159); *d_count += kept (device int32, may pass max_out).  Asynchronous.
 *   d_work >= cvb200_hv_proposals_work_bytes(n_samples) bytes */
size_t cvb200_hv_proposals_work_bytes(int32_t n_samples);
int cvb200_hv_proposals(const int64_t *d_samples, int32_t n_samples, const int32_t *d_arg, const float *d_grid_scale,
                        const int32_t dims[3], float res, const float corner[3], const float *d_seeds, int32_t n_seeds,
                        float radius, int32_t max_out, float *d_loc, float *d_scale, int32_t *d_count, void *d_work,
                        size_t work_bytes, void *stream);

/* ------------------------------------------------------- candidate loop / back-projection ---- */

/* Thresholds of the candidate loop; cvb200_bp_default_params() fills in the reference's module
 * globals (eval_joint.py:18-21: thresh_high=60, thresh_low=10, valid_ratio=0.2, elimination=2)
 * and literals (0.3 / 0.3 at eval_joint.py:245,252). */
typedef struct cvb200_bp_params {
    float thresh_high;          /* stop when max(grid_obj) < thresh_high               (:208) */
    int32_t thresh_low;         /* reject a box with fewer points inside               (:246) */
    float valid_ratio;          /* reject if confident points < valid_ratio * inside   (:246) */
    int32_t elimination;        /* half-width of the zeroed neighbourhood              (:211) */
    int32_t elim_hi_inclusive;  /* 1: [c-e, c+e] (eval_joint.py:211), 0: [c-e, c+e) (eval_separate.py:209) */
    float prob_thresh;          /* "confident" = prob_pred > prob_thresh               (:245) */
    float err_thresh;           /* reject if mean(||xyz_pred - lcc|| * prob) > err_thresh (:250-253) */
    int32_t max_boxes;          /* capacity of the output arrays */
    int32_t max_iters;          /* safety bound on loop iterations */
    int32_t max_trace;          /* capacity (iterations) of d_trace, 0 = no trace */
} cvb200_bp_params;

void cvb200_bp_default_params(cvb200_bp_params *p);
size_t cvb200_bp_work_bytes(const int32_t dims[3]);

/* The reference's inline `while True:` candidate loop (eval_joint.py:204-263, train_joint.py:364-424)
 * as one persistent device loop with no host synchronisation: argmax(grid_obj) -> zero the
 * neighbourhood -> oriented box from grid_rot/grid_scale -> zero the voxels inside it -> LCC-aware
 * back-projection check over all points -> class vote / score / corners.
 *   d_grid_obj [X,Y,Z]      in/out: zeroed in place exactly like the script does
 *   d_grid_rot [X,Y,Z,2], d_grid_scale [X,Y,Z,3]   outputs of cvb200_hv_forward
 *   corner[3], res          grid origin (= min(points,0), eval_joint.py:201,206) and voxel size
 *   d_points,d_xyz [n,3]; d_prob [n]; d_class [n] int64 (torch.argmax dtype), class ids in [0,32)
 *   d_boxes [max_boxes,8,3], d_scores [max_boxes], d_classes [max_boxes] int32
 *   d_counts [2] int32      out: {number of boxes, loop iterations}
 *   d_trace [max_trace,4] int32 (may be NULL): per iteration {peak flat index, points inside,
 *                           confident points inside, accepted} -- verification aid
 * Asynchronous; read d_counts after synchronising the stream. */
int cvb200_back_project(float *d_grid_obj, const float *d_grid_rot, const float *d_grid_scale,
                        const int32_t dims[3], const float corner[3], float res, const float *d_points,
                        const float *d_xyz, const float *d_prob, const int64_t *d_class, int64_t n,
                        const cvb200_bp_params *params, float *d_boxes, float *d_scores, int32_t *d_classes,
                        int32_t *d_counts, int32_t *d_trace, void *d_work, size_t work_bytes, void *stream);

/* Detection post-process after the candidate loop (eval_joint.py:265-280): per-class greedy NMS of oriented boxes with
 * get_iou_obb (utils/calc_map.py:6-21: xz-rectangle intersection x y-overlap), float64 geometry.  d_boxes [k,8,3]
 * (corners 0-3 top face, 4-7 bottom face), d_scores [k], d_classes int32 [k] (classes outside [0,nclasses) are dropped).
 * d_pick receives the kept indices class by class in pick order (score descending; ties: larger index first, i.e. a
 * stable argsort + "take the last" as nms() at eval_joint.py:75-89 does), *d_n_pick their number.  k <= 2048. */
int cvb200_obb_nms(const float *d_boxes, const float *d_scores, const int32_t *d_classes, int32_t k, int32_t nclasses,
                   double overlap_threshold, int32_t *d_pick, int32_t *d_n_pick, void *stream);
/* out[i*nb + j] = get_iou_obb(a[i], b[j]) as float64 (utils/calc_map.py:6-21; used by the mAP evaluation, :78-168). */
int cvb200_obb_iou_matrix(const float *d_a, int32_t na, const float *d_b, int32_t nb, double *d_out, void *stream);

/* ------------------------------------------------- sparse-voxel U-Net: coordinates + convolution ---- */
/* These replace what the reference gets from the external MinkowskiEngine package (v0.5.3, README.md:53):
 * ME.SparseTensor's coordinate manager (train_joint.py:250, eval_joint.py:169) and the kernels behind
 * ME.MinkowskiConvolution / MinkowskiConvolutionTranspose (utils/minkunet.py:53-114, utils/resnet.py:128).
 * Coordinates are int32 rows (batch, x, y, z) as produced by ME.utils.batched_coordinates
 * (train_joint.py:82); batch in [0,65535], x/y/z in [-32768,32767].  Kernel offset k of a K^3 kernel is
 * k = ix + K*(iy + K*iz) (x fastest; ME's ordering from recollection, unpinned -- see DESIGN.md). */

/* slots of the open-addressing hash table for n rows (a power of two >= 2n) */
int64_t cvb200_sc_hash_capacity(int64_t n);

/* (re)build the table: d_keys [capacity] uint64, d_vals [capacity] int32; value = row index */
int cvb200_sc_build_table(const int32_t *d_coords, int64_t n, void *d_keys, int32_t *d_vals, int64_t capacity, void *stream);

/* stride-2 down-sampling, step 1: build the table of COARSE coordinates floor(c / new_stride) * new_stride
 * and flag the first fine child of every coarse voxel (d_flag [n] int32, 0/1).  The caller turns the
 * flags into an exclusive prefix sum (coarse rows are numbered by their first child: deterministic). */
int cvb200_sc_down_flags(const int32_t *d_coords, int64_t n, int32_t new_stride, void *d_keys, int32_t *d_vals,
                         int64_t capacity, int32_t *d_flag, void *stream);

/* step 2: coarse coordinate rows [n_coarse,4], per fine voxel its parent row and kernel offset (0..7),
 * the children table [n_coarse,8] (neighbour table of the stride-2 2^3 convolution) and the parent table
 * [n,8] (neighbour table of the transposed 2^3 convolution: -1 except column koff = parent). */
int cvb200_sc_down_finish(const int32_t *d_coords, int64_t n, int32_t new_stride, void *d_keys, int32_t *d_vals,
                          int64_t capacity, const int32_t *d_flag, const int32_t *d_excl_scan, int64_t n_coarse,
                          int32_t *d_out_coords, int32_t *d_parent, int32_t *d_koff, int32_t *d_children,
                          int32_t *d_up_table, void *stream);

/* neighbour table of a stride-1 convolution with odd kernel size: d_nbr [n_out, ksize^3] =
 * row of (coord + (i - ksize/2) * step) in the table, or -1; step = tensor stride of the level */
int cvb200_sc_kernel_map(const int32_t *d_out_coords, int64_t n_out, const void *d_keys, const int32_t *d_vals,
                         int64_t capacity, int32_t ksize, int32_t step, int32_t *d_nbr, void *stream);

/* out[o,:] = sum_k in[nbr[o,k],:] @ W[k] (+ bias);  d_w [k3,cin,cout], d_nbr [n_out,k3], d_bias [cout] or NULL.
 * fp32 CUDA-core path (exact fp32 accumulation). */
int cvb200_sc_conv_forward(const float *d_in, int32_t cin, const float *d_w, int32_t cout, const int32_t *d_nbr,
                           int64_t n_out, int32_t k3, const float *d_bias, float *d_out, void *stream);

/* Same contraction on the tcgen05 tensor cores (kind::tf32, fp32 accumulation in tensor memory):
 * d_wt is the weight PRE-TRANSPOSED to [k3, cout, cin].  Requires cin % 32 == 0, cout % 16 == 0,
 * 16 <= cout <= 256, k3 <= 32 and 16-byte aligned pointers; n_in = rows of d_in. */
int cvb200_sc_conv_forward_tc(const float *d_in, int64_t n_in, int32_t cin, const float *d_wt, int32_t cout, const int32_t *d_nbr,
                              int64_t n_out, int32_t k3, const float *d_bias, float *d_out, void *stream);

/* Weight gradient on the tensor cores (csrc/sparse_wgrad_tc.cu): dW [k3, cin, cout] = sum_o x[table[o,k]]^T (x) dout[o], the
 * contraction over rows as tcgen05 kind::tf32 MMAs with MN-major operands, fp32 accumulation in tensor memory, partial
 * tiles combined with float atomics.  Needs cin % 32 == 0, cout % 32 == 0, cout <= 256; x row stride = cin. */
int cvb200_sc_conv_wgrad_tc(const float *d_x, int32_t cin, const float *d_dout, int32_t cout, const int32_t *d_table,
                            int64_t n_rows, int32_t k3, float *d_dw, void *stream);

/* Options of the tensor-core convolution.  allow_split = 0: never cut a tile into pieces (no float atomics: bit-reproducible
 * results; slower on small levels).  use_pdl = 0: no programmatic dependent launch.  Defaults: 1, 1. */
int cvb200_sc_set_conv_options(int32_t allow_split, int32_t use_pdl);

/* Measurement aid (tools/conv_probe.py): switch off parts of the kernel to find which side bounds it -- results are
 * garbage while mask != 0.  1 = no gather copies, 2 = no zero-fill copies, 4 = no MMA, 8 = no weight TMA. */
int cvb200_sc_set_conv_debug(int32_t mask);   /* effective only in a probe build (CVB200_PROBE=1 python -m canonicalvoting_b200.build --force) */
/* Measurement aid: device buffer of 3 x 768 int64 that receives clock64 stamps (before wait, after wait, after issue) of the
 * first 256 k-blocks of CTA 0 for the MMA thread, one gather warp and the weight-TMA thread; NULL switches it off. */
int cvb200_sc_set_conv_trace(void *d_trace);

/* Host-only inspection of the work plan of one tensor-core convolution (no device work, no GPU needed): the units the
 * persistent kernel's CTAs walk.  h_plan[12] = {n_tiles, n_splits, n_whole, ks, n_units, total_kb, cblocks, stages, nc,
 * acc_stride, tmem_cols, smem_bytes}; h_units (may be NULL) receives min(n_units, max_units) rows {row0, n0, kb0, kb1, pieces,
 * split_tile}.  Used by the CPU tests to check that every (row tile, channel block, k-block) is covered exactly once. */
int cvb200_sc_conv_plan(int64_t n_out, int32_t cin, int32_t cout, int32_t k3, int32_t *h_plan, int32_t *h_units, int32_t max_units);

/* On-device voxelisation = ME.utils.sparse_quantize (utils/dataloader.py:197, sunrgbd/brnetcanon.py:218): d_xyz float32 [n,3];
 * voxel = floor(p / quantization_size) evaluated in float32 (quantization_size <= 0: floor(p)); d_voxel int32 [n,4] receives
 * (batch, x, y, z) of every point; d_rep[i] = row of the FIRST point of i's voxel, d_flag[i] = 1 iff i is that point.  The
 * caller compacts (indices = positions of the flags, inverse = exclusive_scan(flag)[rep]).  d_keys (uint64) / d_vals (int32):
 * scratch hash map of `capacity` = cvb200_sc_hash_capacity(n) entries. */
int cvb200_sc_quantize(const float *d_xyz, int64_t n, float quantization_size, int32_t batch, void *d_keys, int32_t *d_vals,
                       int64_t capacity, int32_t *d_voxel, int32_t *d_rep, int32_t *d_flag, void *stream);

/* All coordinate levels and kernel maps of a MinkUNet-shaped network in one enqueue, without the host in the loop
 * (csrc/sparse_maps.cu).  Level l has tensor stride 2^l; every table is allocated for the upper bound n inside ONE workspace
 * of layout->total_bytes bytes (offsets below are in bytes); the real sizes are written to counts[0 .. n_down] on the
 * device and copied to h_counts_pinned at the end of the enqueued work: synchronise the stream ONCE, then use
 * rows [0, counts[l]) of each table.  Tables: coords[l] int32 [.,4]; keys/vals: the level's hash map (capacity entries);
 * nbr3[l] int32 [.,27] (3^3 map of the level); stem_table int32 [n, stem_ksize^3] (level 0); arange int32 [n] (identity
 * table of 1x1x1 convolutions); per stride-2 step l -> l+1: children int32 [counts[l+1], 8], up_table int32
 * [counts[l], 8], parent / koff int32 [counts[l]].  Same numbering and offset order as cvb200_sc_down_* / _kernel_map. */
typedef struct cvb200_sc_maps_layout_t {
    int64_t total_bytes, capacity;
    int64_t counts, arange, stem_table;
    int64_t coords[5], keys[5], vals[5], nbr3[5];
    int64_t children[4], up_table[4], parent[4], koff[4];
    int64_t flag, scan, cub_temp, cub_temp_bytes;
    int64_t fill_ff, fill_ff_bytes;      /* the region cvb200_sc_build_maps starts by filling with 0xff (tables, keys, first-child slots) */
} cvb200_sc_maps_layout_t;
int cvb200_sc_maps_layout(int64_t n, int32_t stem_ksize, int32_t n_down, cvb200_sc_maps_layout_t *layout);
int cvb200_sc_build_maps(const int32_t *d_coords, int64_t n, int32_t stem_ksize, int32_t n_down, void *d_workspace,
                         const cvb200_sc_maps_layout_t *layout, int32_t *h_counts_pinned, void *stream);

/* One fused convolution of an inference program (cvb200_sc_run_program):
 *   out[:, 0:cout) (row stride ldo) = [relu]( sum_k in[table[o,k], 0:cin) (row stride ldi) @ W[k] + bias + residual )
 * `in`, `out` and `residual` may point into column slices of wider buffers (that is how ME.cat,
 * utils/minkunet.py:153-177, costs nothing).  kind CVB200_OP_CONV_TC: tcgen05 path, w = [k3,cout,cin]
 * (pre-transposed, BatchNorm folded in), cin % 32 == 0, cout % 16 == 0.  (Kind 1, a CUDA-core kernel for the 3-channel stem,
 * was retired in round 2: the stem runs through CVB200_OP_CONV_TC_GATHER4.) */
#define CVB200_OP_CONV_TC 0
#define CVB200_OP_IM2COL 2        /* out[o, k*cin + c] = in[table[o,k], c] (0 if missing / padding), ldo % cin == 0 */
#define CVB200_OP_CONV_TC_GATHER4 3 /* tcgen05 convolution of a 4-channel input (ldi = 4; the 3-channel stem padded with a zero
                                   * channel): table [n_out, k3], w = [cout][K], K = cin = 32*ceil(k3/8), w[co][4*k + c]; the
                                   * kernel gathers 8 neighbours x 4 channels per k-block, no im2col matrix */
/* Head decode of the joint model fused into the epilogue of the LAST convolution of a program (`final`, utils/minkunet.py:114 ->
 * eval_joint.py:173-193): instead of storing its [n, 64] output the kernel decodes every row while it sits in registers and
 * writes what cvb200_head_decode_points would write -- the N x 64 float32 round trip through HBM disappears.  Needs
 * nclasses == 9 (64 channels = one tile column block); outputs as in cvb200_head_decode_points, d_coords / d_points may be NULL. */
typedef struct cvb200_decode_args {
    float *xyz, *scale;        /* [n,3] each */
    int64_t *class_pred;       /* [n] */
    float *prob;               /* [n] */
    const int32_t *coords;     /* [n,4] rows (batch, x, y, z), or NULL */
    float *points;             /* [n,3] = coords[:, 1:4] * res, or NULL */
    float res;
    int32_t nclasses, log_scale;
} cvb200_decode_args;

typedef struct cvb200_sc_op {
    int32_t kind, cin, cout, k3;
    int32_t ldi, ldo, ldr, relu;
    int64_t n_out, n_in;     /* rows of the output / of the input feature matrix */
    const float *in;
    const float *w;
    const float *bias;       /* [cout] or NULL */
    const float *residual;   /* [n_out, cout] with row stride ldr, or NULL */
    const int32_t *table;    /* [n_out, k3] neighbour table */
    float *out;
    const int32_t *n_out_dev; /* NULL, or the address of the real row count in DEVICE memory (written by cvb200_sc_build_maps into the
                               * counts array of its workspace): n_out is then only an upper bound (the size the buffers have) and the
                               * kernel plans its work itself -- the launch does not depend on a size the host would have to read
                               * back, so a whole program can be captured into one CUDA graph and replayed for any scene of that
                               * bucket.  Tensor-core kinds only. */
    const cvb200_decode_args *decode; /* NULL, or: do not store `out` (may be NULL then), decode the rows instead (see above); HOST pointer */
} cvb200_sc_op;

/* Launch the ops of a program in order on `stream` (host array of ops; asynchronous). */
int cvb200_sc_run_program(const cvb200_sc_op *ops, int32_t n_ops, void *stream);

/* CUDA-graph capture of a sequence of launches of this library (and of anything else enqueued on `stream` and on streams
 * forked from / joined to it in between): begin, enqueue, end -> executable graph; launch it on any stream any number of times.
 * use_node_priority: kernels keep the priority of the stream they were captured on (cudaGraphInstantiateFlagUseNodePriority) --
 * the engine captures the coordinate-map builder on a high-priority branch.  n_nodes (may be NULL) receives the node count.
 * cvb200_graph_abort ends a capture that went wrong.  Thread-local capture mode. */
int cvb200_graph_begin(void *stream);
int cvb200_graph_end(void *stream, int32_t use_node_priority, void **exec_out, int64_t *n_nodes);
int cvb200_graph_abort(void *stream);
int cvb200_graph_launch(void *exec, void *stream);
int cvb200_graph_destroy(void *exec);

/* Head decode of the joint model (eval_joint.py:173-190): d_feats [n, >= 7*nclasses+1] (row stride ld) =
 * xyz[C][3] | scale[C][3] | logits[C+1] -> xyz_pred [n,3], scale_pred [n,3] (exp() if log_scale,
 * config/config.yaml:17), class_pred [n] int64 (argmax over the C object classes), prob_pred [n]. */
int cvb200_head_decode(const float *d_feats, int32_t ld, int64_t n, int32_t nclasses, int32_t log_scale, float *d_xyz,
                       float *d_scale, int64_t *d_class, float *d_prob, void *stream);
/* Same, plus the vote op's first argument in the same pass: d_points [n,3] = d_coords[:, 1:4] * res for int32 coordinate rows
 * (batch, x, y, z) (eval_joint.py:193: `scan_points = coords * res`). */
int cvb200_head_decode_points(const float *d_feats, int32_t ld, int64_t n, int32_t nclasses, int32_t log_scale, float *d_xyz,
                              float *d_scale, int64_t *d_class, float *d_prob, const int32_t *d_coords, float res, float *d_points,
                              void *stream);

/* dW[k] [ca,cb] = sum_r A[ia(r,k),:]^T (x) B[ib(r,k),:] over the table rows r;
 * table_on_b = 0: ia = table[r,k], ib = r;  table_on_b = 1: ia = r, ib = table[r,k].  d_dw is overwritten. */
int cvb200_sc_conv_wgrad(const float *d_a, int32_t ca, const float *d_b, int32_t cb, const int32_t *d_table,
                         int64_t n_rows, int32_t k3, int32_t table_on_b, float *d_dw, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CVB200_H_ */
