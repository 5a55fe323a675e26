// zstd_enc_match.cu -- stage M of the block-parallel Zstandard encoder (sm_100a).
//
// One WARP owns one independent frame (2^frameLog input bytes) and walks it in steps of 32
// positions, lane i owning position base+i.  Per step:
//   A. hash 8 bytes at every position (long: 8-byte hash, short: 5-byte hash), read both
//      table buckets, resolve same-step collisions with __match_any_sync so that every
//      position sees exactly the nearest previous occurrence of its hash key, then store
//      the new entries (highest lane of each key group wins);
//   B. verify the (at most two) candidates against the input (capped at B2Z_CAP bytes),
//      keep the longer; walk the greedy/lazy path through the step's 32 positions; join
//      capped pieces, resolve repcodes, and append literals / final sequences to the
//      block's arrays.
// Stage A of step s+1 is issued before stage B of step s so that the table loads of the next
// step are in flight while the current step's candidates are compared.
//
// Replaces (reference, /root/reference/C/zstd/): zstd_double_fast.c:103-330
// (ZSTD_compressBlock_doubleFast_noDict_generic), the window/table upkeep of
// zstd_compress.c:4591 (ZSTD_compress_frameChunk) and the job slicing of
// zstdmt_compress.c:1184-1246.  The sequential statement of exactly this algorithm is
// oracle/zstd_enc_oracle.c:find_sequences_frame; outputs must be identical.
//
// Memory: tables live in global memory (2^hashLogL + 2^hashLogS u32 per resident warp),
// accessed with L2-only loads/stores (ld.global.cg / st.global.cg); the input is read through
// the read-only path as aligned 8-byte words.
#include "b2z_device.cuh"
#include "b2z_kernels.h"

namespace b2z {

struct StepA {               // per-lane results of stage A
    uint64_t v;              // 8 input bytes at p (zero padded)
    uint32_t candL, candS;   // candidate position + 1 (0 = none), frame-relative
};

// Row-hash match finder, one step.  The table is 2^rowLog rows of 64 bytes holding the 15 most recent entries of the
// row, NEWEST FIRST (word 15 unused): a lookup is one 64-byte load and static register reads, an insert rewrites the
// row shifted by one (the line is dirty anyway).  A position's row is chosen by the 5-byte hash, entries carry a tag
// from the 8-byte hash, so ONE line per position answers both questions of the reference's double-fast finder
// (zstd_double_fast.c:103-330): "newest entry with my tag" (8-byte class, long candidate) and "newest entry"
// (5-byte class, short candidate).  Lanes of one step that fall into the same row behave as if they had inserted in
// lane order (oracle: position-by-position loop; its circular layout holds the same logical content): a lane sees the
// entries of its lower lanes first, then the 15 - rank newest stored ones; the highest lane of the group writes the row.
__device__ __forceinline__ StepA stage_a(const uint64_t* __restrict__ w, uint32_t nWords, uint32_t n, uint32_t base,
                                         uint32_t lane, uint32_t* __restrict__ TR, uint32_t rowLog, uint32_t tagBits) {
    StepA r; r.candL = 0; r.candS = 0;
    const uint32_t p = base + lane, tagMask = (1u << tagBits) - 1u;
    r.v = ld64u(w, p, nWords);
    const bool hashable = p + 8u <= n;
    const uint64_t hl = r.v * B2Z_PRIME8, hs = (r.v << 24) * B2Z_PRIME5;
    const uint32_t rowIdx = (uint32_t)(hs >> (64u - rowLog)), t8 = (uint32_t)(hl >> (64u - tagBits)) & tagMask;
    const uint32_t mine = ((p + 1u) << tagBits) | t8;
    uint4* row4 = reinterpret_cast<uint4*>(TR + (size_t)rowIdx * 16u);
    uint32_t e[16];
    {
        uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0, q2 = q0, q3 = q0;
        if (hashable) { q0 = __ldcg(row4); q1 = __ldcg(row4 + 1); q2 = __ldcg(row4 + 2); q3 = __ldcg(row4 + 3); }
        e[0] = q0.x; e[1] = q0.y; e[2] = q0.z; e[3] = q0.w; e[4] = q1.x; e[5] = q1.y; e[6] = q1.z; e[7] = q1.w;
        e[8] = q2.x; e[9] = q2.y; e[10] = q2.z; e[11] = q2.w; e[12] = q3.x; e[13] = q3.y; e[14] = q3.z; e[15] = 0;
    }
    const uint32_t g = __match_any_sync(B2Z_FULL, hashable ? rowIdx : (0x80000000u | lane));
    const uint32_t lower = g & lanemask_lt();
    if (!__any_sync(B2Z_FULL, lower != 0u)) {
        // ---- common case: every row of this step is hit once
        if (hashable) {
            uint32_t cl = 0;
#pragma unroll
            for (int k = (int)B2Z_ROW_WAYS - 1; k >= 0; k--) if (e[k] && (e[k] & tagMask) == t8) cl = e[k];   // newest match wins
            const uint32_t cs = e[0];
            r.candL = cl >> tagBits; r.candS = (cs != cl) ? (cs >> tagBits) : 0u;
            __stcg(row4, make_uint4(mine, e[0], e[1], e[2])); __stcg(row4 + 1, make_uint4(e[3], e[4], e[5], e[6]));
            __stcg(row4 + 2, make_uint4(e[7], e[8], e[9], e[10])); __stcg(row4 + 3, make_uint4(e[11], e[12], e[13], 0u));
        }
    } else {
        // ---- some row is hit by several lanes: replay their inserts in lane order (ascending = oldest first)
        const uint32_t rank = (uint32_t)__popc(lower), gsize = (uint32_t)__popc(g);
        uint32_t cl = 0, cs = 0;
        for (uint32_t j = 0; j < 32u; j++) {
            const uint32_t ej = __shfl_sync(B2Z_FULL, mine, j);
            if ((lower >> j) & 1u) {                                  // lane j inserted before me: push its entry in front
#pragma unroll
                for (int k = (int)B2Z_ROW_WAYS - 1; k > 0; k--) e[k] = e[k - 1];
                e[0] = ej;
            }
        }
        if (hashable) {
#pragma unroll
            for (int k = (int)B2Z_ROW_WAYS - 1; k >= 0; k--) if (e[k] && (e[k] & tagMask) == t8) cl = e[k];
            cs = e[0];
            r.candL = cl >> tagBits; r.candS = (cs != cl) ? (cs >> tagBits) : 0u;
            if (rank == gsize - 1u) {                                 // the newest lane of the group writes the row
                __stcg(row4, make_uint4(mine, e[0], e[1], e[2])); __stcg(row4 + 1, make_uint4(e[3], e[4], e[5], e[6]));
                __stcg(row4 + 2, make_uint4(e[7], e[8], e[9], e[10])); __stcg(row4 + 3, make_uint4(e[11], e[12], e[13], 0u));
            }
        }
    }
    __syncwarp();
    return r;
}

// Warp-uniform sequence emitter (oracle: emitter_t / emit_flush / emit_match).
struct Emitter {
    uint32_t rep0, rep1, rep2;
    uint32_t pendPos, pendLen, pendOff, pendLast, pendValid;
    uint32_t prevEnd, n;
    uint64_t* out;
    __device__ __forceinline__ void reset(uint64_t* o) {
        rep0 = rep1 = rep2 = 0; pendPos = pendLen = pendOff = pendLast = pendValid = 0; prevEnd = 0; n = 0; out = o;
    }
    __device__ __forceinline__ void flush(uint32_t lane) {
        if (!pendValid) return;
        const uint32_t ll = pendPos - prevEnd, off = pendOff;
        uint32_t code = 0, offBase;
        if (ll) { if (off == rep0) code = 1; else if (off == rep1) code = 2; else if (off == rep2) code = 3; }
        else { if (off == rep1) code = 1; else if (off == rep2) code = 2; else if (rep0 > 1u && off == rep0 - 1u) code = 3; }
        if (code == 0) { offBase = off + 3u; rep2 = rep1; rep1 = rep0; rep0 = off; }
        else {
            offBase = code;
            const uint32_t idx = code - 1u + (ll == 0u);
            if (idx != 0) {
                const uint32_t cur = idx == 3 ? rep0 - 1u : (idx == 1 ? rep1 : rep2);
                if (idx != 1) rep2 = rep1;
                rep1 = rep0; rep0 = cur;
            }
        }
        if (lane == 0) __stcs(reinterpret_cast<unsigned long long*>(out + n), (unsigned long long)B2Z_PACK_SEQ(offBase, ll, pendLen));   // streaming: keep L2 for the rows
        n++; prevEnd = pendPos + pendLen; pendValid = 0;
    }
};

__global__ void __launch_bounds__(B2Z_MATCH_THREADS)
zstd_enc_match_kernel(const uint8_t* __restrict__ src, uint64_t srcSize, EncGeom g, uint32_t* __restrict__ tables,
                      uint64_t* __restrict__ seqs, uint32_t* __restrict__ nseq,
                      uint8_t* __restrict__ lits, uint32_t* __restrict__ nlit,
                      const volatile uint32_t* ready, uint32_t readyShift) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warpSlot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nWarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t rowLog = g.rowLog, tagBits = 32u - (g.frameLog + 1u);
    const uint64_t F = 1ull << g.frameLog;
    const uint32_t W = g.windowLog >= 32 ? 0xFFFFFFFFu : (1u << g.windowLog);
    const uint32_t blocksPerFrame = (uint32_t)(F >> 17);
    const uint64_t nFrames = (srcSize + F - 1) >> g.frameLog;
    const uint32_t tableWords = 16u << rowLog;
    uint32_t* TR = tables + (size_t)warpSlot * tableWords;

    for (uint64_t f = warpSlot; f < nFrames; f += nWarps) {
        const uint64_t f0 = f << g.frameLog;
        const uint32_t n = enc_frame_bytes(g, srcSize, f);
        const uint64_t* __restrict__ w = reinterpret_cast<const uint64_t*>(src + f0);
        const uint32_t nWords = (n + 7u) >> 3;
        // host-pointer path: the input is still being uploaded chunk by chunk while this kernel runs; a frame starts
        // once the flag of the chunk that holds its last byte has been set (by a stream-ordered copy after the chunk)
        if (ready) { uint32_t spins = 0; while (ready[(f0 + n - 1u) >> readyShift] == 0u && ++spins < (1u << 21)) __nanosleep(2000); __syncwarp(); }   // bounded: never hangs the GPU
        // clear this warp's tables (16-byte stores)
        {
            uint4* t4 = reinterpret_cast<uint4*>(TR);
            const uint4 z = make_uint4(0, 0, 0, 0);
            for (uint32_t i = lane; i < tableWords / 4u; i += 32u) __stcg(t4 + i, z);
            __syncwarp();
        }
        const size_t blk0 = (size_t)f * blocksPerFrame;
        uint32_t entry = 0, nlitB = 0;
        Emitter em; em.reset(seqs + blk0 * B2Z_MAXSEQ);
        StepA cur = stage_a(w, nWords, n, 0, lane, TR, rowLog, tagBits);
        for (uint32_t base = 0; base < n; base += 32u) {
            const uint32_t blk = base >> 17, blkStart = blk << 17;
            const uint32_t blkEnd = (blkStart + B2Z_BLOCK < n) ? blkStart + B2Z_BLOCK : n;
            const uint32_t p = base + lane;
            // ---- stage A of the next step (table traffic in flight during stage B)
            StepA nxt; nxt.v = 0; nxt.candL = 0; nxt.candS = 0;
            if (base + 32u < n) nxt = stage_a(w, nWords, n, base + 32u, lane, TR, rowLog, tagBits);
            // ---- stage B
            const uint32_t cnt = (n - base) < 32u ? (n - base) : 32u;       // valid lanes
            if (entry < base + cnt) {
                uint32_t len = 0, off = 0;
                if (p + 8u <= n) {
                    uint32_t maxLen = blkEnd - p; if (maxLen > B2Z_CAP) maxLen = B2Z_CAP;
                    uint32_t lenL = 0, offL = 0, lenS = 0, offS = 0;
                    if (cur.candL) { const uint32_t q = cur.candL - 1u; if (p - q <= W) { offL = p - q; lenL = match_len_pv(w, q, p, cur.v, maxLen, nWords); } }
                    if (cur.candS) { const uint32_t q = cur.candS - 1u; if (p - q <= W) { offS = p - q; lenS = match_len_pv(w, q, p, cur.v, maxLen, nWords); } }
                    len = lenL; off = offL;
                    if (lenS > lenL || (lenS == lenL && lenS && offS < offL)) { len = lenS; off = offS; }
                    if (!b2z_accept(len, off)) { len = 0; off = 0; }
                }
                // path through the step: greedy with one-position lazy deferral
                const uint32_t lenNext = __shfl_down_sync(B2Z_FULL, len, 1);
                const bool defer = len && (lane + 1u < cnt) && lenNext >= len + B2Z_LAZY_GAIN;
                const uint32_t takeMask = __ballot_sync(B2Z_FULL, len && !defer);
                uint32_t c = (entry > base ? entry - base : 0u), litMask = 0;
                while (c < cnt) {
                    const uint32_t m = takeMask >> c;
                    const uint32_t j = m ? (uint32_t)(__ffs((int)m) - 1) : (cnt - c);   // literals before next match
                    const uint32_t jj = (c + j > cnt) ? (cnt - c) : j;
                    if (jj) litMask |= ((jj >= 32u ? 0xFFFFFFFFu : ((1u << jj) - 1u)) << c);
                    c += jj;
                    if (c >= cnt) break;
                    // take the match at lane c
                    const uint32_t mLen = __shfl_sync(B2Z_FULL, len, c), mOff = __shfl_sync(B2Z_FULL, off, c);
                    const uint32_t pos = base + c - blkStart;
                    bool merged = false;
                    if (em.pendValid && pos == em.pendPos + em.pendLen && em.pendLast == B2Z_CAP) {
                        bool same = mOff == em.pendOff;
                        if (!same) {                                            // does the pending offset cover this piece too?
                            const uint32_t k = lane * 8u;
                            bool eq = true;
                            if (k < mLen) {
                                const uint64_t a = ld64u(w, base + c + k, nWords), b = ld64u(w, base + c + k - em.pendOff, nWords);
                                uint64_t x = a ^ b;
                                if (mLen - k < 8u) x &= (1ull << ((mLen - k) * 8u)) - 1ull;
                                eq = x == 0;
                            }
                            same = __all_sync(B2Z_FULL, eq);
                        }
                        if (same) { em.pendLen += mLen; em.pendLast = mLen; merged = true; }
                    }
                    if (!merged) {
                        em.flush(lane);
                        em.pendPos = pos; em.pendLen = mLen; em.pendOff = mOff; em.pendLast = mLen; em.pendValid = 1;
                    }
                    c += mLen;
                }
                entry = base + c;
                if (litMask >> lane & 1u) __stcs(lits + f0 + blkStart + nlitB + __popc(litMask & lanemask_lt()), (uint8_t)cur.v);
                nlitB += __popc(litMask);
            }
            if (base + cnt == blkEnd) {                                         // block finished
                em.flush(lane);
                if (lane == 0) { nseq[blk0 + blk] = em.n; nlit[blk0 + blk] = nlitB; }
                nlitB = 0;
                em.reset(seqs + (blk0 + blk + 1) * B2Z_MAXSEQ);
            }
            cur = nxt;
        }
        __syncwarp();
    }
}

#ifndef B2Z_CUEMU
void launch_zstd_enc_match(const uint8_t* src, uint64_t srcSize, const EncGeom& g, uint32_t* tables, uint32_t nWarps,
                           uint64_t* seqs, uint32_t* nseq, uint8_t* lits, uint32_t* nlit, const uint32_t* ready, uint32_t readyShift,
                           cudaStream_t st) {
    if (srcSize == 0) return;
    const uint32_t warpsPerCta = B2Z_MATCH_THREADS / 32;
    const uint32_t grid = (nWarps + warpsPerCta - 1) / warpsPerCta;
    zstd_enc_match_kernel<<<grid, B2Z_MATCH_THREADS, 0, st>>>(src, srcSize, g, tables, seqs, nseq, lits, nlit, ready, readyShift);
}
#endif

}  // namespace b2z
