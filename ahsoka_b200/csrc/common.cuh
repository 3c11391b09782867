// common.cuh — shared device helpers for the sm_100a phasing kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ahs {

constexpr uint64_t KEY_NONE    = ~0ull;
constexpr int32_t  W_CLAMP     = 1 << 17;
constexpr int      MAX_ALLELES = 15;                 // mask bits 0..14, bit 15 = full-containment flag
constexpr int      MAX_PLOIDY  = 6;
constexpr int      MAX_K       = 2 * MAX_PLOIDY;     // |covMap[pos]| <= 2p (alignmentstoreadset.cpp:766)
constexpr int      PR_K        = 8;                  // clusters per DP column kept in a PosRec: 2p up to ploidy 4, p + 2 above (rule R3c)
constexpr int      DP_INF      = 1 << 29;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// first index in [0,n) with off[idx+1] > x, i.e. the CSR owner of element x
__device__ __forceinline__ int owner_of(const int64_t* __restrict__ off, int n, int64_t x) {
    int lo = 0, hi = n - 1;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (off[mid + 1] > x) hi = mid; else lo = mid + 1; }
    return lo;
}

__device__ __forceinline__ uint32_t hash64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return (uint32_t)k;
}

constexpr int CH_HAS_UNIV = 1 << 16;       // ch_flags bit: some bubble of the chain has an allele path of <= 2 nodes (SURVEY A#9)
// slot of a node id in an open-addressing table of mask + 1 = 2^k >= 4 slots: Fibonacci hashing (the top k bits of x * 2^32/phi);
// consecutive ids — the usual numbering of a chain's nodes — land far apart
__device__ __forceinline__ uint32_t hash_slot(uint32_t x, uint32_t mask) { return (x * 0x9E3779B1u) >> __clz(mask); }
__device__ __forceinline__ uint32_t hash32(uint32_t x) {      // murmur3 finaliser
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    return x;
}

__device__ __forceinline__ void atomic_or_u16(uint16_t* p, uint16_t v) {
    uintptr_t a = (uintptr_t)p;
    unsigned int* w = (unsigned int*)(a & ~(uintptr_t)3);
    atomicOr(w, (unsigned int)v << ((a & 2) ? 16 : 0));
}

__device__ __forceinline__ uint64_t make_key(uint32_t pos, uint32_t allele, uint32_t entry) {
    return ((uint64_t)pos << 40) | ((uint64_t)allele << 32) | entry;
}

// int(float(id) * 100): float*int -> float, then truncation (alignmentstoreadset.cpp:117,:234)
__device__ __forceinline__ int32_t mapq_of(float identity) { return (int32_t)__fmul_rn(identity, 100.0f); }
// (alignment.id*100) > 90, float compare (alignmentstoreadset.cpp:245)
__device__ __forceinline__ bool good_identity(float identity) { return __fmul_rn(identity, 100.0f) > 90.0f; }

__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
    for (int o = 16; o > 0; o >>= 1) { uint64_t t = __shfl_xor_sync(0xffffffffu, v, o); v = t < v ? t : v; }
    return v;
}
__device__ __forceinline__ int64_t warp_sum_i64(int64_t v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i32(int v) { return __reduce_add_sync(0xffffffffu, v); }
__device__ __forceinline__ int warp_max_i32(int v) { return __reduce_max_sync(0xffffffffu, v); }
__device__ __forceinline__ int warp_min_i32(int v) { return __reduce_min_sync(0xffffffffu, v); }

// packed code rows: `bits` per cell (2 or 4), code 0 = missing, code = allele + 1
__device__ __forceinline__ uint32_t get_code(const uint32_t* __restrict__ row, int b, int bits) {
    int bitpos = b * bits;
    return (row[bitpos >> 5] >> (bitpos & 31)) & ((1u << bits) - 1u);
}

// overlap (n) and disagreement (k) of two packed words
template <int BITS>
__device__ __forceinline__ void word_nk(uint32_t x, uint32_t y, int& n, int& k) {
    uint32_t px, py, d;
    if (BITS == 2) {
        px = (x | (x >> 1)) & 0x55555555u; py = (y | (y >> 1)) & 0x55555555u;
        d = x ^ y; d = (d | (d >> 1)) & 0x55555555u;
    } else {
        px = (x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x11111111u; py = (y | (y >> 1) | (y >> 2) | (y >> 3)) & 0x11111111u;
        d = x ^ y; d = (d | (d >> 1) | (d >> 2) | (d >> 3)) & 0x11111111u;
    }
    uint32_t both = px & py;
    n += __popc(both);
    k += __popc(d & both);
}

__device__ __forceinline__ int64_t floordiv1024(int64_t a) { return a >> 10; }   // arithmetic shift = floor for negatives

}  // namespace ahs
