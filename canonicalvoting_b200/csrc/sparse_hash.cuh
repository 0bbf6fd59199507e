// canonicalvoting_b200/csrc/sparse_hash.cuh -- device hash map over packed (batch, x, y, z) voxel coordinates
// (shared by sparse_coords.cu and sparse_maps.cu).
#pragma once
#include "common.cuh"

namespace cvb200 {

constexpr unsigned long long kEmptyKey = ~0ULL;

// 16 bits per field: batch in [0, 65535], x/y/z in [-32768, 32767]
__host__ __device__ __forceinline__ unsigned long long pack_coord(int b, int x, int y, int z) {
    return ((unsigned long long)(b & 0xffff) << 48) | ((unsigned long long)((x + 32768) & 0xffff) << 32) |
           ((unsigned long long)((y + 32768) & 0xffff) << 16) | (unsigned long long)((z + 32768) & 0xffff);
}

__device__ __forceinline__ unsigned int hash_key(unsigned long long k) {   // murmur3 finaliser
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return (unsigned int)k;
}

// returns the slot of `key`, inserting it if absent (linear probing; capacity is a power of two >= 2n)
__device__ __forceinline__ unsigned int hash_insert(unsigned long long *keys, unsigned int mask, unsigned long long key) {
    unsigned int h = hash_key(key) & mask;
    while (true) {
        const unsigned long long prev = atomicCAS(keys + h, kEmptyKey, key);
        if (prev == kEmptyKey || prev == key) return h;
        h = (h + 1) & mask;
    }
}

__device__ __forceinline__ int hash_lookup(const unsigned long long *__restrict__ keys, const int *__restrict__ vals,
                                           unsigned int mask, unsigned long long key) {
    unsigned int h = hash_key(key) & mask;
    while (true) {
        const unsigned long long k = __ldg(keys + h);
        if (k == key) return __ldg(vals + h);
        if (k == kEmptyKey) return -1;
        h = (h + 1) & mask;
    }
}


}  // namespace cvb200
