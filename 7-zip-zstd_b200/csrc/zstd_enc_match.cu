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

__device__ __forceinline__ StepA stage_a(const uint64_t* __restrict__ w, uint32_t nWords, uint32_t n, uint32_t base,
                                         uint32_t lane, uint32_t* __restrict__ TL, uint32_t* __restrict__ TS,
                                         uint32_t HL, uint32_t HS, uint32_t tagBits) {
    StepA r; r.candL = 0; r.candS = 0;
    const uint32_t p = base + lane, tagMask = (1u << tagBits) - 1u;
    r.v = ld64u(w, p, nWords);
    const bool hashable = p + 8u <= n;
    const uint64_t hl = r.v * B2Z_PRIME8, hs = (r.v << 24) * B2Z_PRIME5;
    const uint32_t keyL = (uint32_t)(hl >> (64u - HL - tagBits)), keyS = (uint32_t)(hs >> (64u - HS - tagBits));
    uint32_t* sl = TL + (keyL >> tagBits);
    uint32_t* ss = TS + (keyS >> tagBits);
    uint32_t eL = 0, eS = 0;
    if (hashable) { eL = __ldcg(sl); eS = __ldcg(ss); }
    // same-step resolution: the latest lower lane that hits the same BUCKET owns the slot (it would
    // have overwritten it in sequential order); it is a candidate only if its tag matches too
    const uint32_t uniq = 0x80000000u | lane;                // buckets are < 2^22
    const uint32_t mL = __match_any_sync(B2Z_FULL, hashable ? (keyL >> tagBits) : uniq);
    const uint32_t mS = __match_any_sync(B2Z_FULL, hashable ? (keyS >> tagBits) : uniq);
    const uint32_t lt = lanemask_lt();
    const uint32_t lowL = mL & lt, lowS = mS & lt;
    const uint32_t jL = lowL ? highbit32(lowL) : 0u, jS = lowS ? highbit32(lowS) : 0u;
    const uint32_t keyLj = __shfl_sync(B2Z_FULL, keyL, jL), keySj = __shfl_sync(B2Z_FULL, keyS, jS);
    if (hashable) {
        if (lowL) { if (keyLj == keyL) r.candL = base + jL + 1u; }
        else if (eL && (eL & tagMask) == (keyL & tagMask)) r.candL = eL >> tagBits;
        if (lowS) { if (keySj == keyS) r.candS = base + jS + 1u; }
        else if (eS && (eS & tagMask) == (keyS & tagMask)) r.candS = eS >> tagBits;
        // insert: the highest lane of each bucket group holds the latest position
        if ((mL >> lane) == 1u) __stcg(sl, ((p + 1u) << tagBits) | (keyL & tagMask));
        if ((mS >> lane) == 1u) __stcg(ss, ((p + 1u) << tagBits) | (keyS & tagMask));
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
        if (lane == 0) out[n] = B2Z_PACK_SEQ(offBase, ll, pendLen);
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
    const uint32_t HL = g.hashLogL, HS = g.hashLogS, tagBits = 32u - (g.frameLog + 1u);
    const uint64_t F = 1ull << g.frameLog;
    const uint32_t W = g.windowLog >= 32 ? 0xFFFFFFFFu : (1u << g.windowLog);
    const uint32_t blocksPerFrame = (uint32_t)(F >> 17);
    const uint64_t nFrames = (srcSize + F - 1) >> g.frameLog;
    const uint32_t tableWords = (1u << HL) + (1u << HS);
    uint32_t* TL = tables + (size_t)warpSlot * tableWords;
    uint32_t* TS = TL + (1u << HL);

    for (uint64_t f = warpSlot; f < nFrames; f += nWarps) {
        const uint64_t f0 = f << g.frameLog;
        const uint32_t n = (uint32_t)((srcSize - f0) < F ? (srcSize - f0) : F);
        const uint64_t* __restrict__ w = reinterpret_cast<const uint64_t*>(src + f0);
        const uint32_t nWords = (n + 7u) >> 3;
        // host-pointer path: the input is still being uploaded chunk by chunk while this kernel runs; a frame starts
        // once the flag of the chunk that holds its last byte has been set (by a stream-ordered copy after the chunk)
        if (ready) { uint32_t spins = 0; while (ready[(f0 + n - 1u) >> readyShift] == 0u && ++spins < (1u << 21)) __nanosleep(2000); __syncwarp(); }   // bounded: never hangs the GPU
        // clear this warp's tables (16-byte stores)
        {
            uint4* t4 = reinterpret_cast<uint4*>(TL);
            const uint4 z = make_uint4(0, 0, 0, 0);
            for (uint32_t i = lane; i < tableWords / 4u; i += 32u) __stcg(t4 + i, z);
            __syncwarp();
        }
        const size_t blk0 = (size_t)f * blocksPerFrame;
        uint32_t entry = 0, nlitB = 0;
        Emitter em; em.reset(seqs + blk0 * B2Z_MAXSEQ);
        StepA cur = stage_a(w, nWords, n, 0, lane, TL, TS, HL, HS, tagBits);
        for (uint32_t base = 0; base < n; base += 32u) {
            const uint32_t blk = base >> 17, blkStart = blk << 17;
            const uint32_t blkEnd = (blkStart + B2Z_BLOCK < n) ? blkStart + B2Z_BLOCK : n;
            const uint32_t p = base + lane;
            // ---- stage A of the next step (table traffic in flight during stage B)
            StepA nxt; nxt.v = 0; nxt.candL = 0; nxt.candS = 0;
            if (base + 32u < n) nxt = stage_a(w, nWords, n, base + 32u, lane, TL, TS, HL, HS, tagBits);
            // ---- stage B
            const uint32_t cnt = (n - base) < 32u ? (n - base) : 32u;       // valid lanes
            if (entry < base + cnt) {
                uint32_t len = 0, off = 0;
                if (p + 8u <= n) {
                    uint32_t maxLen = blkEnd - p; if (maxLen > B2Z_CAP) maxLen = B2Z_CAP;
                    uint32_t lenL = 0, offL = 0, lenS = 0, offS = 0;
                    if (cur.candL) { const uint32_t q = cur.candL - 1u; if (p - q <= W) { offL = p - q; lenL = match_len(w, q, p, maxLen, nWords); } }
                    if (cur.candS) { const uint32_t q = cur.candS - 1u; if (p - q <= W) { offS = p - q; lenS = match_len(w, q, p, maxLen, nWords); } }
                    len = lenL; off = offL;
                    if (lenS > lenL || (lenS == lenL && offS < offL)) { len = lenS; off = offS; }
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
                if (litMask >> lane & 1u) lits[f0 + blkStart + nlitB + __popc(litMask & lanemask_lt())] = (uint8_t)cur.v;
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

void launch_zstd_enc_match(const uint8_t* src, uint64_t srcSize, const EncGeom& g, uint32_t* tables, uint32_t nWarps,
                           uint64_t* seqs, uint32_t* nseq, uint8_t* lits, uint32_t* nlit, const uint32_t* ready, uint32_t readyShift,
                           cudaStream_t st) {
    if (srcSize == 0) return;
    const uint32_t warpsPerCta = B2Z_MATCH_THREADS / 32;
    const uint32_t grid = (nWarps + warpsPerCta - 1) / warpsPerCta;
    zstd_enc_match_kernel<<<grid, B2Z_MATCH_THREADS, 0, st>>>(src, srcSize, g, tables, seqs, nseq, lits, nlit, ready, readyShift);
}

}  // namespace b2z
