// canonicalvoting_b200/csrc/sparse_maps.cu -- all coordinate levels and kernel maps of the U-Net in ONE enqueue.
//
// Round 2: 33 graph nodes per scene instead of 56 -- everything that starts as all-ones lies in one region (one memset instead of
// fifteen), the row counts and the identity table are written by the kernels that know them (no one-thread kernels).
//
// sparse_coords.cu exposes the coordinate manager step by step (the module path mirrors MinkowskiEngine's lazy
// behaviour: a level or kernel map is created when the first layer asks for it) and reads the size of every new
// level back to the host: 4 blocking reads + ~45 launches issued from Python per scene, 1.2 ms of host time for a
// 50k-voxel scan -- more than the rest of the inference step together (tools/host_profile.py).  The inference engine
// knows up front what it needs (utils/minkunet.py:50-120: a 5^3 stem map, four stride-2 levels, one 3^3 map, one
// children table and one parent table per level), so this file builds everything without the host in the loop:
//   * every level is allocated for the upper bound n (a coarse level never has more voxels than the input) inside one
//     caller-provided workspace; the real sizes live in a device array `counts[level]` that the kernels read;
//   * kernels are grid-stride loops over *counts, launched with a fixed grid, so no launch depends on a size;
//   * the counts are copied to pinned host memory at the end: ONE synchronisation per scene (by the caller), after
//     which the host slices the tables (`[:counts[l]]`) and plans the convolutions.
// Numbering (coarse voxel = rank of its first fine child), offset order and table formats are those of
// sparse_coords.cu, so both paths produce identical tables (tests/test_sparse_gpu.py).
#include <cub/device/device_scan.cuh>

#include <string.h>

#include "common.cuh"
#include "sparse_hash.cuh"

namespace cvb200 {

constexpr int kMapThreads = 256;
constexpr int kMapBlocks = 4 * kNumSMs;

// level 0: hash the input rows; the same pass writes the identity table of the 1x1x1 convolutions and the level's row count
__global__ void mp_insert_rows_kernel(const int4 *__restrict__ coords, int n, unsigned long long *keys, int *vals, unsigned int mask,
                                      int *__restrict__ arange, int *__restrict__ counts) {
    if (blockIdx.x == 0 && threadIdx.x == 0) counts[0] = n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 c = __ldg(coords + i);
        vals[hash_insert(keys, mask, pack_coord(c.x, c.y, c.z, c.w))] = i;
        arange[i] = i;
    }
}

__device__ __forceinline__ int4 mp_coarse_of(int4 c, int shift) {
    return make_int4(c.x, (c.y >> shift) << shift, (c.z >> shift) << shift, (c.w >> shift) << shift);
}

__global__ void mp_insert_coarse_min_kernel(const int4 *__restrict__ coords, const int *__restrict__ cnt, int shift,
                                            unsigned long long *keys, int *vals, unsigned int mask) {
    const int n = *cnt;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 c = mp_coarse_of(__ldg(coords + i), shift);
        // vals pre-filled with 0xffffffff (the same memset as the keys and the tables): unsigned minimum
        atomicMin(reinterpret_cast<unsigned int *>(vals) + hash_insert(keys, mask, pack_coord(c.x, c.y, c.z, c.w)), (unsigned int)i);
    }
}

// flag[i] = 1 iff fine voxel i is the first child of its coarse voxel; 0 beyond the level's size (the scan runs over n_ub)
__global__ void mp_first_child_kernel(const int4 *__restrict__ coords, const int *__restrict__ cnt, int n_ub, int shift,
                                      const unsigned long long *__restrict__ keys, const int *__restrict__ vals, unsigned int mask,
                                      int *__restrict__ flag) {
    const int n = *cnt;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_ub; i += gridDim.x * blockDim.x) {
        int f = 0;
        if (i < n) {
            const int4 c = mp_coarse_of(__ldg(coords + i), shift);
            f = hash_lookup(keys, vals, mask, pack_coord(c.x, c.y, c.z, c.w)) == i;
        }
        flag[i] = f;
    }
}

// numbers the coarse voxels (rank of the first child) and records how many there are: counts[level + 1]
__global__ void mp_number_coarse_kernel(const int4 *__restrict__ coords, const int *__restrict__ cnt, int shift,
                                        const int *__restrict__ flag, const int *__restrict__ excl, unsigned long long *keys, int *vals,
                                        unsigned int mask, int4 *__restrict__ out_coords, int n_ub, int *__restrict__ count_out) {
    const int n = *cnt;
    if (blockIdx.x == 0 && threadIdx.x == 0) *count_out = excl[n_ub - 1] + flag[n_ub - 1];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (!flag[i]) continue;
        const int4 c = mp_coarse_of(__ldg(coords + i), shift);
        const int row = excl[i];
        out_coords[row] = c;
        vals[hash_insert(keys, mask, pack_coord(c.x, c.y, c.z, c.w))] = row;
    }
}

__global__ void mp_link_children_kernel(const int4 *__restrict__ coords, const int *__restrict__ cnt, int shift,
                                        const unsigned long long *__restrict__ keys, const int *__restrict__ vals, unsigned int mask,
                                        int *__restrict__ parent, int *__restrict__ koff, int *__restrict__ children,
                                        int *__restrict__ up_table) {
    const int n = *cnt;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 f = __ldg(coords + i);
        const int4 c = mp_coarse_of(f, shift);
        const int p = hash_lookup(keys, vals, mask, pack_coord(c.x, c.y, c.z, c.w));
        const int half = shift - 1;
        const int k = ((f.y - c.y) >> half) + 2 * (((f.z - c.z) >> half) + 2 * ((f.w - c.w) >> half));
        parent[i] = p;
        koff[i] = k;
        children[8 * (size_t)p + k] = i;   // pre-filled with -1
        const int4 lo = make_int4(k == 0 ? p : -1, k == 1 ? p : -1, k == 2 ? p : -1, k == 3 ? p : -1);
        const int4 hi = make_int4(k == 4 ? p : -1, k == 5 ? p : -1, k == 6 ? p : -1, k == 7 ? p : -1);
        reinterpret_cast<int4 *>(up_table)[2 * (size_t)i] = lo;
        reinterpret_cast<int4 *>(up_table)[2 * (size_t)i + 1] = hi;
    }
}

// Stride-1 kernel map of a level onto itself.  The map is symmetric -- if j is the neighbour of o at offset k, o is the
// neighbour of j at the mirrored offset K^3-1-k -- so only the lower half of the offsets is looked up in the hash map and
// the mirrored entry is written alongside (the table is pre-filled with -1, the centre offset is the voxel itself):
// half the hash probes of the 5^3 stem map (6.25 M for a 50k-voxel scene) and of the 3^3 maps.
__global__ void mp_kernel_map_kernel(const int4 *__restrict__ coords, const int *__restrict__ cnt, const unsigned long long *__restrict__ keys,
                                     const int *__restrict__ vals, unsigned int mask, int ksize, int step, int *__restrict__ nbr) {
    const int k3 = ksize * ksize * ksize, h = ksize / 2, mid = k3 / 2;
    const long long total = (long long)(*cnt) * (mid + 1);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(t / (mid + 1)), k = (int)(t - (long long)o * (mid + 1));
        if (k == mid) {
            nbr[(size_t)o * k3 + mid] = o;
            continue;
        }
        const int ix = k % ksize, iy = (k / ksize) % ksize, iz = k / (ksize * ksize);
        const int4 c = __ldg(coords + o);
        const int j = hash_lookup(keys, vals, mask, pack_coord(c.x, c.y + (ix - h) * step, c.z + (iy - h) * step, c.w + (iz - h) * step));
        if (j >= 0) {
            nbr[(size_t)o * k3 + k] = j;
            nbr[(size_t)j * k3 + (k3 - 1 - k)] = o;
        }
    }
}

static int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }

}  // namespace cvb200

using namespace cvb200;

extern "C" int cvb200_sc_maps_layout(int64_t n, int32_t stem_ksize, int32_t n_down, cvb200_sc_maps_layout_t *L) {
    CVB_REQUIRE(L && n > 0 && n < (1LL << 31) && n_down >= 0 && n_down <= 4 && (stem_ksize == 0 || (stem_ksize & 1)) && stem_ksize <= 7,
                CVB200_EINVAL, "sc_maps_layout: bad argument (n=%lld, stem %d, levels %d)", (long long)n, stem_ksize, n_down);
    memset(L, 0, sizeof(*L));
    int64_t cap = 1024;
    while (cap < 2 * n) cap <<= 1;
    L->capacity = cap;
    int64_t off = 0;
    auto take = [&](int64_t bytes) { const int64_t o = off; off += align256(bytes); return o; };
    L->counts = take(64);
    L->arange = take(4 * n);
    // everything that starts as all-ones (-1 table entries, empty hash keys, "no first child yet") is contiguous: ONE memset
    L->fill_ff = off;
    L->stem_table = stem_ksize ? take(4 * n * stem_ksize * stem_ksize * stem_ksize) : 0;
    for (int l = 0; l <= n_down; l++) L->nbr3[l] = take(4 * n * 27);
    for (int l = 0; l <= n_down; l++) L->keys[l] = take(8 * cap);
    for (int l = 1; l <= n_down; l++) L->vals[l] = take(4 * cap);
    for (int l = 0; l < n_down; l++) L->children[l] = take(32 * n);
    L->fill_ff_bytes = off - L->fill_ff;
    L->vals[0] = take(4 * cap);
    for (int l = 1; l <= n_down; l++) L->coords[l] = take(16 * n);
    for (int l = 0; l < n_down; l++) {
        L->up_table[l] = take(32 * n);
        L->parent[l] = take(4 * n);
        L->koff[l] = take(4 * n);
    }
    L->flag = take(4 * n);
    L->scan = take(4 * n);
    size_t temp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, temp, (const int *)nullptr, (int *)nullptr, (int)n);
    L->cub_temp_bytes = (int64_t)temp;
    L->cub_temp = take((int64_t)temp + 256);
    L->total_bytes = off;
    return 0;
}

extern "C" int cvb200_sc_build_maps(const int32_t *d_coords, int64_t n, int32_t stem_ksize, int32_t n_down, void *d_ws,
                                    const cvb200_sc_maps_layout_t *L, int32_t *h_counts_pinned, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVB_REQUIRE(d_coords && d_ws && L && h_counts_pinned && n > 0 && n < (1LL << 31) && n_down >= 0 && n_down <= 4, CVB200_EINVAL,
                "sc_build_maps: bad argument");
    unsigned char *ws = (unsigned char *)d_ws;
    int *counts = (int *)(ws + L->counts);
    const unsigned int mask = (unsigned int)(L->capacity - 1);
    const int n_ub = (int)n;
    auto keys = [&](int l) { return (unsigned long long *)(ws + L->keys[l]); };
    auto vals = [&](int l) { return (int *)(ws + L->vals[l]); };
    auto coords = [&](int l) { return l ? (const int4 *)(ws + L->coords[l]) : (const int4 *)d_coords; };

    CVB_CUDA(cudaMemsetAsync(ws + L->fill_ff, 0xff, (size_t)L->fill_ff_bytes, stream));
    mp_insert_rows_kernel<<<kMapBlocks, kMapThreads, 0, stream>>>(coords(0), n_ub, keys(0), vals(0), mask, (int *)(ws + L->arange), counts);
    if (stem_ksize)
        mp_kernel_map_kernel<<<8 * kMapBlocks, kMapThreads, 0, stream>>>(coords(0), counts, keys(0), vals(0), mask, stem_ksize, 1,
                                                                        (int *)(ws + L->stem_table));
    mp_kernel_map_kernel<<<kMapBlocks, kMapThreads, 0, stream>>>(coords(0), counts, keys(0), vals(0), mask, 3, 1, (int *)(ws + L->nbr3[0]));
    for (int l = 0; l < n_down; l++) {
        const int shift = l + 1;
        int *flag = (int *)(ws + L->flag), *scan = (int *)(ws + L->scan);
        mp_insert_coarse_min_kernel<<<kMapBlocks, kMapThreads, 0, stream>>>(coords(l), counts + l, shift, keys(l + 1), vals(l + 1), mask);
        mp_first_child_kernel<<<kMapBlocks, kMapThreads, 0, stream>>>(coords(l), counts + l, n_ub, shift, keys(l + 1), vals(l + 1), mask, flag);
        size_t temp = (size_t)L->cub_temp_bytes;
        CVB_CUDA(cub::DeviceScan::ExclusiveSum(ws + L->cub_temp, temp, (const int *)flag, scan, n_ub, stream));
        mp_number_coarse_kernel<<<kMapBlocks, kMapThreads, 0, stream>>>(coords(l), counts + l, shift, flag, scan, keys(l + 1), vals(l + 1), mask,
                                                                       (int4 *)(ws + L->coords[l + 1]), n_ub, counts + l + 1);
        mp_link_children_kernel<<<kMapBlocks, kMapThreads, 0, stream>>>(coords(l), counts + l, shift, keys(l + 1), vals(l + 1), mask,
                                                                       (int *)(ws + L->parent[l]), (int *)(ws + L->koff[l]),
                                                                       (int *)(ws + L->children[l]), (int *)(ws + L->up_table[l]));
        mp_kernel_map_kernel<<<kMapBlocks, kMapThreads, 0, stream>>>(coords(l + 1), counts + l + 1, keys(l + 1), vals(l + 1), mask, 3, 1 << shift,
                                                                    (int *)(ws + L->nbr3[l + 1]));
    }
    CVB_LAUNCH_CHECK("sc_build_maps");
    CVB_CUDA(cudaMemcpyAsync(h_counts_pinned, counts, 4 * (size_t)(n_down + 1), cudaMemcpyDeviceToHost, stream));
    return 0;
}
