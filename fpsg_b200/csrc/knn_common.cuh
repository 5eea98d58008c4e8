// Ranking keys and warp-level sorting networks shared by the kNN kernels (knn.cu, knn_tc.cu).
#pragma once
#include "common.cuh"

namespace fpsg {

typedef unsigned long long u64;

__device__ __forceinline__ unsigned f32_to_ord(float f) {
    const unsigned u = __float_as_uint(f);
    return u ^ (((int)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float ord_to_f32(unsigned o) {
    return __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
}
__device__ __forceinline__ u64 key_cmpx(u64 key, int d, bool take_max) {
    const u64 o = __shfl_xor_sync(0xffffffffu, key, d);
    const bool gt = key > o;
    return (gt == take_max) ? key : o;
}
// sorts the warp's 32 keys descending by lane
__device__ __forceinline__ u64 warp_sort_desc(u64 key, int lane) {
#pragma unroll
    for (int s = 2; s <= 32; s <<= 1) {
        const bool desc = s == 32 ? true : ((lane & s) == 0);
#pragma unroll
        for (int d = s >> 1; d > 0; d >>= 1) key = key_cmpx(key, d, ((lane & d) == 0) == desc);
    }
    return key;
}
// bitonic sequence -> descending
__device__ __forceinline__ u64 warp_merge_desc(u64 key, int lane) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) key = key_cmpx(key, d, (lane & d) == 0);
    return key;
}

__device__ __forceinline__ float max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// key = (order-preserving bits of pd) << 32 | (0xFFFFFFFF - j): larger key = nearer, lower index first among equal values
__device__ __forceinline__ u64 knn_key_of(float pd, int j) { return ((u64)f32_to_ord(pd) << 32) | (u64)(0xffffffffu - (unsigned)j); }

}  // namespace fpsg
