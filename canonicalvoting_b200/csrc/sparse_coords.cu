// canonicalvoting_b200/csrc/sparse_coords.cu -- coordinate manager of the sparse-voxel U-Net.
//
// Replaces what the reference gets from MinkowskiEngine's coordinate manager (external package, v0.5.3 by
// README.md:53; call sites: ME.SparseTensor at train_joint.py:250 / eval_joint.py:169 and every
// MinkowskiConvolution / MinkowskiConvolutionTranspose in utils/minkunet.py:53-114): a GPU hash map over
// (batch, x, y, z) voxel coordinates, the coordinate set of the next tensor stride, and kernel maps.
// Design (not ME's): kernel maps are stored OUTPUT-STATIONARY as dense neighbour tables
//     nbr[n_out][K^3] = row of the input voxel at  out_coord + offset_k * tensor_stride,  or -1
// which is what an implicit-GEMM convolution wants (one gather per (output tile, offset), no scatter, no
// atomics); stride-2 convolutions get a children table [n_coarse][8] and the transposed convolution a
// parent table of the same shape as a 2^3 neighbour table.  Coarse voxels are numbered by their first
// fine child, so every table is deterministic.
#include "common.cuh"
#include "sparse_hash.cuh"

namespace cvb200 {

__global__ void sc_insert_rows_kernel(const int4 *__restrict__ coords, int n, unsigned long long *keys, int *vals,
                                      unsigned int mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = __ldg(coords + i);   // (b, x, y, z)
    vals[hash_insert(keys, mask, pack_coord(c.x, c.y, c.z, c.w))] = i;
}

// coarse key of a fine voxel: floor(c / 2^shift) * 2^shift (arithmetic shift = floor for negatives)
__device__ __forceinline__ int4 coarse_of(int4 c, int shift) {
    return make_int4(c.x, (c.y >> shift) << shift, (c.z >> shift) << shift, (c.w >> shift) << shift);
}

__global__ void sc_insert_coarse_min_kernel(const int4 *__restrict__ coords, int n, int shift, unsigned long long *keys,
                                            int *vals, unsigned int mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = coarse_of(__ldg(coords + i), shift);
    atomicMin(vals + hash_insert(keys, mask, pack_coord(c.x, c.y, c.z, c.w)), i);   // vals pre-filled with INT_MAX
}

// flag[i] = 1 iff fine voxel i is the first (smallest row) child of its coarse voxel
__global__ void sc_first_child_kernel(const int4 *__restrict__ coords, int n, int shift,
                                      const unsigned long long *__restrict__ keys, const int *__restrict__ vals,
                                      unsigned int mask, int *__restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = coarse_of(__ldg(coords + i), shift);
    flag[i] = hash_lookup(keys, vals, mask, pack_coord(c.x, c.y, c.z, c.w)) == i;
}

// pass 1: first children write the coarse coordinate row and re-point the hash value at the coarse row
__global__ void sc_number_coarse_kernel(const int4 *__restrict__ coords, int n, int shift, const int *__restrict__ flag,
                                        const int *__restrict__ excl_scan, unsigned long long *keys, int *vals,
                                        unsigned int mask, int4 *__restrict__ out_coords) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    const int4 c = coarse_of(__ldg(coords + i), shift);
    const int row = excl_scan[i];
    out_coords[row] = c;
    vals[hash_insert(keys, mask, pack_coord(c.x, c.y, c.z, c.w))] = row;
}

// pass 2: parent row and kernel offset of every fine voxel; children table of every coarse voxel.
// offset index k = ix + 2*(iy + 2*iz), i* in {0,1} = (c - parent) / tensor_stride   [ME-recall: x fastest]
__global__ void sc_link_children_kernel(const int4 *__restrict__ coords, int n, int shift,
                                        const unsigned long long *__restrict__ keys, const int *__restrict__ vals,
                                        unsigned int mask, int *__restrict__ parent, int *__restrict__ koff,
                                        int *__restrict__ children, int *__restrict__ up_table) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 f = __ldg(coords + i);
    const int4 c = coarse_of(f, shift);
    const int p = hash_lookup(keys, vals, mask, pack_coord(c.x, c.y, c.z, c.w));
    const int half = shift - 1;   // fine tensor stride = 2^(shift-1)
    const int k = ((f.y - c.y) >> half) + 2 * (((f.z - c.z) >> half) + 2 * ((f.w - c.w) >> half));
    parent[i] = p;
    koff[i] = k;
    children[8 * (size_t)p + k] = i;   // pre-filled with -1
#pragma unroll
    for (int j = 0; j < 8; j++) up_table[8 * (size_t)i + j] = j == k ? p : -1;
}

// nbr[o][k] = row of (coord_o + offset_k * step) in the table, offsets centred for odd ksize:
// k = ix + K*(iy + K*iz), offset = (i - K/2) * step                                  [ME-recall: x fastest]
__global__ void sc_kernel_map_kernel(const int4 *__restrict__ out_coords, int n_out, const unsigned long long *__restrict__ keys,
                                     const int *__restrict__ vals, unsigned int mask, int ksize, int step,
                                     int *__restrict__ nbr) {
    const int k3 = ksize * ksize * ksize;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_out * k3) return;
    const int o = (int)(t / k3), k = (int)(t - (long long)o * k3);
    const int ix = k % ksize, iy = (k / ksize) % ksize, iz = k / (ksize * ksize);
    const int4 c = __ldg(out_coords + o);
    const int h = ksize / 2;
    nbr[t] = hash_lookup(keys, vals, mask, pack_coord(c.x, c.y + (ix - h) * step, c.z + (iy - h) * step, c.w + (iz - h) * step));
}

}  // namespace cvb200

using namespace cvb200;

static bool pow2(int64_t c) { return c > 0 && (c & (c - 1)) == 0; }

extern "C" int64_t cvb200_sc_hash_capacity(int64_t n) {
    int64_t c = 1024;
    while (c < 2 * n) c <<= 1;
    return c;
}

extern "C" int cvb200_sc_build_table(const int32_t *d_coords, int64_t n, void *d_keys, int32_t *d_vals, int64_t capacity,
                                     void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVB_REQUIRE(d_keys && d_vals && pow2(capacity) && capacity >= 2 * n && n >= 0 && n < (1LL << 31), CVB200_EINVAL,
                "sc_build_table: bad table (capacity %lld for %lld rows)", (long long)capacity, (long long)n);
    CVB_CUDA(cudaMemsetAsync(d_keys, 0xff, sizeof(unsigned long long) * (size_t)capacity, stream));
    if (n == 0) return 0;
    CVB_REQUIRE(d_coords, CVB200_EINVAL, "sc_build_table: NULL coords");
    sc_insert_rows_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>((const int4 *)d_coords, (int)n,
                                                                         (unsigned long long *)d_keys, d_vals,
                                                                         (unsigned int)(capacity - 1));
    CVB_LAUNCH_CHECK("sc_insert_rows_kernel");
    return 0;
}

extern "C" int cvb200_sc_down_flags(const int32_t *d_coords, int64_t n, int32_t new_stride, void *d_keys, int32_t *d_vals,
                                    int64_t capacity, int32_t *d_flag, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVB_REQUIRE(d_coords && d_keys && d_vals && d_flag && n > 0 && n < (1LL << 31), CVB200_EINVAL, "sc_down_flags: bad argument");
    CVB_REQUIRE(pow2(capacity) && capacity >= 2 * n && pow2(new_stride) && new_stride >= 2, CVB200_EINVAL,
                "sc_down_flags: capacity / stride must be powers of two");
    int shift = 0;
    while ((1 << shift) < new_stride) shift++;
    const unsigned int mask = (unsigned int)(capacity - 1);
    CVB_CUDA(cudaMemsetAsync(d_keys, 0xff, sizeof(unsigned long long) * (size_t)capacity, stream));
    CVB_CUDA(cudaMemsetAsync(d_vals, 0x7f, sizeof(int) * (size_t)capacity, stream));
    const unsigned blocks = (unsigned)ceil_div(n, 256);
    sc_insert_coarse_min_kernel<<<blocks, 256, 0, stream>>>((const int4 *)d_coords, (int)n, shift, (unsigned long long *)d_keys,
                                                            d_vals, mask);
    CVB_LAUNCH_CHECK("sc_insert_coarse_min_kernel");
    sc_first_child_kernel<<<blocks, 256, 0, stream>>>((const int4 *)d_coords, (int)n, shift, (const unsigned long long *)d_keys,
                                                      d_vals, mask, d_flag);
    CVB_LAUNCH_CHECK("sc_first_child_kernel");
    return 0;
}

extern "C" int cvb200_sc_down_finish(const int32_t *d_coords, int64_t n, int32_t new_stride, void *d_keys, int32_t *d_vals,
                                     int64_t capacity, const int32_t *d_flag, const int32_t *d_excl_scan, int64_t n_coarse,
                                     int32_t *d_out_coords, int32_t *d_parent, int32_t *d_koff, int32_t *d_children,
                                     int32_t *d_up_table, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVB_REQUIRE(d_coords && d_keys && d_vals && d_flag && d_excl_scan && d_out_coords && d_parent && d_koff && d_children &&
                    d_up_table && n > 0 && n_coarse > 0,
                CVB200_EINVAL, "sc_down_finish: bad argument");
    CVB_REQUIRE(pow2(capacity) && pow2(new_stride) && new_stride >= 2, CVB200_EINVAL, "sc_down_finish: capacity / stride");
    int shift = 0;
    while ((1 << shift) < new_stride) shift++;
    const unsigned int mask = (unsigned int)(capacity - 1);
    const unsigned blocks = (unsigned)ceil_div(n, 256);
    CVB_CUDA(cudaMemsetAsync(d_children, 0xff, sizeof(int) * 8 * (size_t)n_coarse, stream));
    sc_number_coarse_kernel<<<blocks, 256, 0, stream>>>((const int4 *)d_coords, (int)n, shift, d_flag, d_excl_scan,
                                                        (unsigned long long *)d_keys, d_vals, mask, (int4 *)d_out_coords);
    CVB_LAUNCH_CHECK("sc_number_coarse_kernel");
    sc_link_children_kernel<<<blocks, 256, 0, stream>>>((const int4 *)d_coords, (int)n, shift, (const unsigned long long *)d_keys,
                                                        d_vals, mask, d_parent, d_koff, d_children, d_up_table);
    CVB_LAUNCH_CHECK("sc_link_children_kernel");
    return 0;
}

extern "C" int cvb200_sc_kernel_map(const int32_t *d_out_coords, int64_t n_out, const void *d_keys, const int32_t *d_vals,
                                    int64_t capacity, int32_t ksize, int32_t step, int32_t *d_nbr, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVB_REQUIRE(d_out_coords && d_keys && d_vals && d_nbr && n_out > 0 && pow2(capacity), CVB200_EINVAL, "sc_kernel_map: bad argument");
    CVB_REQUIRE(ksize >= 1 && ksize <= 7 && (ksize & 1) && step >= 1, CVB200_EINVAL, "sc_kernel_map: odd kernel size in [1,7] expected");
    const int64_t total = n_out * ksize * ksize * ksize;
    sc_kernel_map_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>((const int4 *)d_out_coords, (int)n_out,
                                                                            (const unsigned long long *)d_keys, d_vals,
                                                                            (unsigned int)(capacity - 1), ksize, step, d_nbr);
    CVB_LAUNCH_CHECK("sc_kernel_map_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// On-device voxelisation: ME.utils.sparse_quantize (utils/dataloader.py:197, sunrgbd/brnetcanon.py:218) -- the step right
// before the hot path.  voxel = floor(p / quantization_size) in float32 (numpy's arithmetic for a float32 array), one
// representative point per occupied voxel: the FIRST in input order (atomicMin of the row index per voxel key), so the
// result is deterministic and equal to the host implementation in canonicalvoting_b200/sparse/utils.py.
namespace cvb200 {

__global__ void sc_quantize_insert_kernel(const float *__restrict__ xyz, int n, float qsize, int batch, int4 *__restrict__ voxel,
                                          unsigned long long *keys, int *vals, unsigned int mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = __ldg(xyz + 3 * (size_t)i), y = __ldg(xyz + 3 * (size_t)i + 1), z = __ldg(xyz + 3 * (size_t)i + 2);
    if (qsize > 0.f) { x = __fdiv_rn(x, qsize); y = __fdiv_rn(y, qsize); z = __fdiv_rn(z, qsize); }
    const int4 v = make_int4(batch, (int)floorf(x), (int)floorf(y), (int)floorf(z));
    voxel[i] = v;
    atomicMin(vals + hash_insert(keys, mask, pack_coord(v.x, v.y, v.z, v.w)), i);   // vals pre-filled with INT_MAX
}

// rep[i] = first point of i's voxel, flag[i] = 1 iff i is that point
__global__ void sc_quantize_lookup_kernel(const int4 *__restrict__ voxel, int n, const unsigned long long *__restrict__ keys,
                                          const int *__restrict__ vals, unsigned int mask, int *__restrict__ rep, int *__restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 v = voxel[i];
    const int r = hash_lookup(keys, vals, mask, pack_coord(v.x, v.y, v.z, v.w));
    rep[i] = r;
    flag[i] = r == i;
}

}  // namespace cvb200

extern "C" int cvb200_sc_quantize(const float *d_xyz, int64_t n, float quantization_size, int32_t batch, void *d_keys, int32_t *d_vals,
                                  int64_t capacity, int32_t *d_voxel, int32_t *d_rep, int32_t *d_flag, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVB_REQUIRE(n >= 0 && n < (1LL << 31) && pow2(capacity) && capacity >= 2 * n && batch >= 0 && batch < 65536, CVB200_EINVAL,
                "sc_quantize: bad sizes (n=%lld, capacity=%lld, batch=%d)", (long long)n, (long long)capacity, batch);
    CVB_REQUIRE(d_keys && d_vals, CVB200_EINVAL, "sc_quantize: NULL table");
    CVB_CUDA(cudaMemsetAsync(d_keys, 0xff, sizeof(unsigned long long) * (size_t)capacity, stream));
    CVB_CUDA(cudaMemsetAsync(d_vals, 0x7f, sizeof(int) * (size_t)capacity, stream));
    if (n == 0) return 0;
    CVB_REQUIRE(d_xyz && d_voxel && d_rep && d_flag, CVB200_EINVAL, "sc_quantize: NULL argument");
    const unsigned blocks = (unsigned)ceil_div(n, 256);
    const unsigned int mask = (unsigned int)(capacity - 1);
    sc_quantize_insert_kernel<<<blocks, 256, 0, stream>>>(d_xyz, (int)n, quantization_size, batch, (int4 *)d_voxel,
                                                         (unsigned long long *)d_keys, d_vals, mask);
    sc_quantize_lookup_kernel<<<blocks, 256, 0, stream>>>((const int4 *)d_voxel, (int)n, (const unsigned long long *)d_keys, d_vals, mask,
                                                         d_rep, d_flag);
    CVB_LAUNCH_CHECK("sc_quantize");
    return 0;
}
