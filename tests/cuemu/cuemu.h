// cuemu.h -- TEST INFRASTRUCTURE ONLY: a minimal host emulation of the warp-synchronous CUDA subset the kernels in
// 7-zip-zstd_b200/csrc use, so that the KERNEL SOURCES can be compiled with g++ and their logic compared with the oracle on a
// machine without a GPU (tests/test_cuemu_*.py, `-m "not gpu"`).  It is not a CPU path of the product: nothing outside tests/
// includes it, libb200z.so is built by nvcc only and fails without a device.
//
// Model: one CTA at a time; every CUDA thread is a fiber (own stack, cooperative switch); a fiber runs until its next warp or
// block collective, where it waits for the other live lanes -- i.e. lanes do NOT advance in lockstep between collectives, which
// is a legal schedule under independent thread scheduling and catches code that forgets a __syncwarp().  Shared memory is
// filled with 0xCD before each CTA (uninitialised reads show up), global memory is the host's.  What it cannot show: timing,
// memory-model races between warps, anything about the hardware.
#pragma once
#if !defined(__x86_64__)
#error "cuemu: x86-64 only (hand-written context switch)"
#endif
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <algorithm>
#include <vector>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __constant__ static const
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint4 { uint32_t x, y, z, w; } __attribute__((aligned(16)));
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
struct uint2 { uint32_t x, y; } __attribute__((aligned(8)));
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r; r.x = x; r.y = y; return r; }
struct dim3 { uint32_t x, y, z; dim3(uint32_t a = 1, uint32_t b = 1, uint32_t c = 1) : x(a), y(b), z(c) {} };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
using std::max;
using std::min;

extern "C" void cuemu_switch(void** saveSp, void* newSp);
asm(".text\n.globl cuemu_switch\n.type cuemu_switch,@function\ncuemu_switch:\n"
    "pushq %rbp\npushq %rbx\npushq %r12\npushq %r13\npushq %r14\npushq %r15\n"
    "movq %rsp, (%rdi)\nmovq %rsi, %rsp\n"
    "popq %r15\npopq %r14\npopq %r13\npopq %r12\npopq %rbx\npopq %rbp\nret\n");

namespace cuemu {

struct Fiber { void* sp; char* stack; bool done; dim3 tid; };
struct Warp { uint64_t slot[32], res[2][32]; uint32_t resMask[2], arrivedMask, gen, liveMask; };
struct Block {
    std::vector<Fiber> fibers; std::vector<Warp> warps;
    uint32_t cur = 0, nThreads = 0, liveThreads = 0, barArrived = 0, barGen = 0;
    uint32_t nbArrived[16] = {0}, nbGen[16] = {0};      // named barriers (bar.sync / bar.arrive id, count)
    dim3 bidx, bdim, gdim;
    void* mainSp = nullptr;
    unsigned char* dynSmem = nullptr;
    std::function<void()> body;
    uint64_t collectives = 0;
};
inline Block*& blk() { static Block* b = nullptr; return b; }
static const size_t kStack = 256 * 1024;

inline void yield() { Block* b = blk(); cuemu_switch(&b->fibers[b->cur].sp, b->mainSp); }
inline uint32_t lane() { Block* b = blk(); return b->cur & 31u; }
inline Warp& warp() { Block* b = blk(); return b->warps[b->cur >> 5]; }

inline void warp_complete(Warp& w) {
    const uint32_t g = w.gen & 1u;
    memcpy(w.res[g], w.slot, sizeof(w.slot)); w.resMask[g] = w.arrivedMask; w.arrivedMask = 0; w.gen++;
}
// every live lane of the warp contributes v; returns once all have arrived; out = all lanes' values, *mask = who contributed
inline void collect(uint64_t v, uint64_t out[32], uint32_t* mask) {
    Warp& w = warp(); const uint32_t ln = lane(), gen = w.gen;
    blk()->collectives++;
    w.slot[ln] = v; w.arrivedMask |= 1u << ln;
    if (w.arrivedMask == w.liveMask) warp_complete(w);
    else while (w.gen == gen) yield();
    memcpy(out, w.res[gen & 1u], sizeof(w.slot)); *mask = w.resMask[gen & 1u];
}
inline void fiber_exit() {
    Block* b = blk(); Fiber& f = b->fibers[b->cur];
    Warp& w = b->warps[b->cur >> 5];
    w.liveMask &= ~(1u << (b->cur & 31u));
    if (w.arrivedMask && w.arrivedMask == w.liveMask) warp_complete(w);
    b->liveThreads--;
    if (b->barArrived && b->barArrived == b->liveThreads) { b->barArrived = 0; b->barGen++; }
    f.done = true;
    for (;;) yield();
}
inline void fiber_entry() { blk()->body(); fiber_exit(); }

inline void run_block(Block& b) {
    blk() = &b;
    b.liveThreads = b.nThreads; b.barArrived = 0; b.barGen = 0;
    memset(b.nbArrived, 0, sizeof(b.nbArrived)); memset(b.nbGen, 0, sizeof(b.nbGen));
    b.warps.assign((b.nThreads + 31) / 32, Warp());
    for (auto& w : b.warps) memset(&w, 0, sizeof(w));
    for (uint32_t t = 0; t < b.nThreads; t++) {
        Fiber& f = b.fibers[t];
        f.done = false; f.tid = dim3(t % b.bdim.x, (t / b.bdim.x) % b.bdim.y, t / (b.bdim.x * b.bdim.y));
        b.warps[t >> 5].liveMask |= 1u << (t & 31u);
        uintptr_t top = ((uintptr_t)(f.stack + kStack)) & ~(uintptr_t)15;
        void** sp = (void**)top;
        *--sp = nullptr;                       // fake return address of fiber_entry (it never returns)
        *--sp = (void*)&fiber_entry;           // popped by cuemu_switch's ret
        for (int k = 0; k < 6; k++) *--sp = nullptr;
        f.sp = sp;
    }
    uint32_t live = b.nThreads;
    while (live) {
        live = 0;
        for (uint32_t t = 0; t < b.nThreads; t++) {
            if (b.fibers[t].done) continue;
            b.cur = t;
            cuemu_switch(&b.mainSp, b.fibers[t].sp);
            if (!b.fibers[t].done) live++;
        }
    }
}

// run `call` (a lambda that calls the kernel function with its arguments) for every CTA of the grid, one after another
template <class F> inline uint64_t launch(dim3 grid, dim3 block, size_t dynSmemBytes, F call) {
    Block b; b.bdim = block; b.gdim = grid; b.nThreads = block.x * block.y * block.z;
    b.fibers.resize(b.nThreads);
    for (auto& f : b.fibers) f.stack = (char*)malloc(kStack);
#ifdef __SANITIZE_ADDRESS__
    b.dynSmem = (unsigned char*)malloc(dynSmemBytes ? dynSmemBytes : 1);       // exact size: an overrun of the dynamic shared memory hits a redzone
#else
    b.dynSmem = (unsigned char*)aligned_alloc(128, (dynSmemBytes + 255) & ~(size_t)127);
#endif
    b.body = call;
    for (uint32_t z = 0; z < grid.z; z++) for (uint32_t y = 0; y < grid.y; y++) for (uint32_t x = 0; x < grid.x; x++) {
        b.bidx = dim3(x, y, z);
        memset(b.dynSmem, 0xCD, dynSmemBytes);
        run_block(b);
    }
    for (auto& f : b.fibers) free(f.stack);
    free(b.dynSmem);
    blk() = nullptr;
    return b.collectives;
}
inline void* dyn_smem() { return blk()->dynSmem; }
// bar.sync id, n (wait = true) / bar.arrive id, n (wait = false): the n-th arrival releases the barrier's generation
inline void named_bar(uint32_t id, uint32_t n, bool wait) {
    Block* b = blk(); const uint32_t gen = b->nbGen[id];
    if (++b->nbArrived[id] == n) { b->nbArrived[id] = 0; b->nbGen[id]++; }
    else if (wait) while (b->nbGen[id] == gen) yield();
}

}  // namespace cuemu

#define threadIdx (cuemu::blk()->fibers[cuemu::blk()->cur].tid)
#define blockIdx  (cuemu::blk()->bidx)
#define blockDim  (cuemu::blk()->bdim)
#define gridDim   (cuemu::blk()->gdim)

// ---- warp collectives (mask argument: the kernels always pass the full mask; the participants are the warp's live lanes)
template <class T> inline T __shfl_sync(uint32_t, T v, int src) { uint64_t o[32]; uint32_t m; uint64_t b = 0; memcpy(&b, &v, sizeof(T)); cuemu::collect(b, o, &m); T r; memcpy(&r, &o[src & 31], sizeof(T)); return r; }
template <class T> inline T __shfl_up_sync(uint32_t, T v, unsigned d) { uint64_t o[32]; uint32_t m; uint64_t b = 0; memcpy(&b, &v, sizeof(T)); cuemu::collect(b, o, &m); const uint32_t l = cuemu::lane(); T r; memcpy(&r, &o[l >= d ? l - d : l], sizeof(T)); return r; }
template <class T> inline T __shfl_down_sync(uint32_t, T v, unsigned d) { uint64_t o[32]; uint32_t m; uint64_t b = 0; memcpy(&b, &v, sizeof(T)); cuemu::collect(b, o, &m); const uint32_t l = cuemu::lane(); T r; memcpy(&r, &o[l + d < 32 ? l + d : l], sizeof(T)); return r; }
template <class T> inline T __shfl_xor_sync(uint32_t, T v, int x) { uint64_t o[32]; uint32_t m; uint64_t b = 0; memcpy(&b, &v, sizeof(T)); cuemu::collect(b, o, &m); T r; memcpy(&r, &o[(cuemu::lane() ^ (uint32_t)x) & 31], sizeof(T)); return r; }
inline uint32_t __ballot_sync(uint32_t, int pred) { uint64_t o[32]; uint32_t m, r = 0; cuemu::collect(pred ? 1 : 0, o, &m); for (int i = 0; i < 32; i++) if ((m >> i & 1u) && o[i]) r |= 1u << i; return r; }
inline int __any_sync(uint32_t k, int pred) { return __ballot_sync(k, pred) != 0; }
inline int __all_sync(uint32_t, int pred) { uint64_t o[32]; uint32_t m; cuemu::collect(pred ? 1 : 0, o, &m); for (int i = 0; i < 32; i++) if ((m >> i & 1u) && !o[i]) return 0; return 1; }
template <class T> inline uint32_t __match_any_sync(uint32_t, T v) { uint64_t o[32]; uint32_t m, r = 0; uint64_t b = 0; memcpy(&b, &v, sizeof(T)); cuemu::collect(b, o, &m); for (int i = 0; i < 32; i++) if ((m >> i & 1u) && o[i] == b) r |= 1u << i; return r; }
inline void __syncwarp(uint32_t = 0xFFFFFFFFu) { uint64_t o[32]; uint32_t m; cuemu::collect(0, o, &m); }
inline void __syncthreads() {
    cuemu::Block* b = cuemu::blk(); const uint32_t gen = b->barGen;
    if (++b->barArrived == b->liveThreads) { b->barArrived = 0; b->barGen++; } else while (b->barGen == gen) cuemu::yield();
}

// ---- scalar intrinsics
inline int __popc(uint32_t v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (sh & 31u)); }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }
template <class T, class U> inline void __stcg(T* p, U v) { *p = (T)v; }
template <class T, class U> inline void __stcs(T* p, U v) { *p = (T)v; }
inline void __nanosleep(unsigned) { cuemu::yield(); }
inline void __threadfence() {}
inline uint32_t atomicOr(uint32_t* p, uint32_t v) { const uint32_t o = *p; *p |= v; return o; }
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { const uint32_t o = *p; *p += v; return o; }
inline uint32_t atomicMax(uint32_t* p, uint32_t v) { const uint32_t o = *p; if (v > o) *p = v; return o; }
inline uint32_t atomicMin(uint32_t* p, uint32_t v) { const uint32_t o = *p; if (v < o) *p = v; return o; }
inline uint32_t atomicExch(uint32_t* p, uint32_t v) { const uint32_t o = *p; *p = v; return o; }

inline uint32_t cuemu_lane_id() { return cuemu::lane(); }
