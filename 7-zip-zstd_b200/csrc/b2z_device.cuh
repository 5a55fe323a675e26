// b2z_device.cuh -- small device helpers shared by the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "b2z_params.h"

#define B2Z_FULL 0xFFFFFFFFu

namespace b2z {

__device__ __forceinline__ uint32_t lane_id() { uint32_t l; asm volatile("mov.u32 %0, %%laneid;" : "=r"(l)); return l; }
__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// 64-bit funnel: bytes [s/8, s/8+8) of the 16-byte little-endian pair (a, b); s in {0,8,..,56}
__device__ __forceinline__ uint64_t funnel64(uint64_t a, uint64_t b, uint32_t s) {
    return (a >> s) | ((b << 1) << (63u - s));
}

// read-only (non-coherent, L1-cacheable) aligned 8-byte load of word `i` of the frame
__device__ __forceinline__ uint64_t ldw(const uint64_t* __restrict__ w, uint32_t i, uint32_t nWords) {
    return i < nWords ? __ldg(w + i) : 0ull;
}

// unaligned 8 bytes at byte offset `o` of a frame whose base is 8-byte aligned; bytes at or
// beyond nWords*8 read as zero (never dereferenced)
__device__ __forceinline__ uint64_t ld64u(const uint64_t* __restrict__ w, uint32_t o, uint32_t nWords) {
    uint32_t i = o >> 3, s = (o & 7u) * 8u;
    uint64_t a = ldw(w, i, nWords), b = s ? ldw(w, i + 1, nWords) : 0ull;
    return funnel64(a, b, s);
}

// common-prefix length of frame[q..] and frame[p..], capped at maxLen (q < p)
__device__ __forceinline__ uint32_t match_len(const uint64_t* __restrict__ w, uint32_t q, uint32_t p,
                                              uint32_t maxLen, uint32_t nWords) {
    uint32_t qi = q >> 3, pi = p >> 3, qs = (q & 7u) * 8u, ps = (p & 7u) * 8u;
    uint64_t qa = ldw(w, qi, nWords), pa = ldw(w, pi, nWords);
    uint32_t len = 0;
    while (len < maxLen) {
        uint64_t qb = ldw(w, qi + 1, nWords), pb = ldw(w, pi + 1, nWords);
        uint64_t x = funnel64(qa, qb, qs) ^ funnel64(pa, pb, ps);
        if (x) { len += (uint32_t)(__ffsll((long long)x) - 1) >> 3; break; }
        len += 8; qa = qb; pa = pb; qi++; pi++;
    }
    return len < maxLen ? len : maxLen;
}

__device__ __forceinline__ uint32_t highbit32(uint32_t v) { return 31u - (uint32_t)__clz((int)v); }

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t lane, uint32_t* total) {
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(B2Z_FULL, x, d); if (lane >= (uint32_t)d) x += y; }
    *total = __shfl_sync(B2Z_FULL, x, 31);
    return x - v;
}

}  // namespace b2z
