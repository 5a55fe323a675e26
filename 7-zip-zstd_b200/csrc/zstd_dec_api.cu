// zstd_dec_api.cu -- decoder half of the C ABI (include/b200z.h): drives stages D0..D3.
//
// Replaces the streaming loop of CPP/7zip/Compress/ZstdDecoder.cpp:108-173 (ZSTD_decompressStream
// over 128 KiB reads, frame after frame): here one call takes the whole Code() input, every frame
// of it (zstd or skippable) is located by the prepass and decoded in parallel.
#include "b2z_ctx.h"
#include <vector>
#include <thread>
#include <mutex>

using namespace b2z;

static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

static int dec_status_to_rc(b200z_ctx* ctx, uint32_t st) {
    if (st & B2Z_DERR_TABLE_FULL) return fail(ctx, B200Z_E_UNSUPPORTED, "more frames/blocks than the decoder tables hold%s");
    if (st & B2Z_DERR_UNSUPPORTED) return fail(ctx, B200Z_E_UNSUPPORTED, "dictionary or window > 1 GiB frames are not supported%s");
    if (st & B2Z_DERR_CORRUPT) return fail(ctx, B200Z_E_CORRUPT, "corrupt zstd data%s");          // a declared size that the blocks contradict is corruption, whatever else was noticed
    if (st & B2Z_DERR_DSTSIZE) return fail(ctx, B200Z_E_DSTSIZE, "destination too small%s");
    if (st & B2Z_DERR_CHECKSUM) return fail(ctx, B200Z_E_CHECKSUM, "content checksum mismatch%s");
    return 0;
}

extern "C" {

// Walk frame headers (and block headers) on the host: cheap, sequential, no entropy decoding.
// Follows ZSTD_getFrameHeader / ZSTD_findFrameCompressedSize (C/zstd/zstd_decompress.c:482-600,702).
int b200z_zstd_frame_info(const void* srcv, size_t srcSize, uint64_t* contentSize, uint32_t* nFrames) {
    const uint8_t* ip = (const uint8_t*)srcv; const uint8_t* iend = ip + srcSize;
    uint64_t total = 0; uint32_t frames = 0; int unknown = 0;
    while (ip < iend) {
        if (iend - ip < 4) return B200Z_E_CORRUPT;
        const uint32_t magic = rd32(ip);
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {
            if (iend - ip < 8) return B200Z_E_CORRUPT;
            const uint32_t sz = rd32(ip + 4);
            if ((size_t)(iend - ip) < 8 + (size_t)sz) return B200Z_E_CORRUPT;
            ip += 8 + sz; continue;
        }
        if (magic != 0xFD2FB528u) return B200Z_E_CORRUPT;
        if (iend - ip < 6) return B200Z_E_CORRUPT;
        const uint32_t fhd = ip[4]; ip += 5;
        const uint32_t fcsFlag = fhd >> 6, single = (fhd >> 5) & 1, checksum = (fhd >> 2) & 1, didFlag = fhd & 3;
        if (fhd & 8) return B200Z_E_CORRUPT;
        if (!single) ip += 1;
        static const int didBytes[4] = { 0, 1, 2, 4 };
        ip += didBytes[didFlag];
        const int fcsBytes = fcsFlag == 0 ? (int)single : (fcsFlag == 1 ? 2 : (fcsFlag == 2 ? 4 : 8));
        if (iend - ip < fcsBytes) return B200Z_E_CORRUPT;
        if (fcsBytes) { uint64_t fcs = 0; for (int i = 0; i < fcsBytes; i++) fcs |= (uint64_t)ip[i] << (8 * i); if (fcsBytes == 2) fcs += 256; total += fcs; }
        else unknown = 1;
        ip += fcsBytes;
        for (;;) {
            if (iend - ip < 3) return B200Z_E_CORRUPT;
            const uint32_t bh = ip[0] | (ip[1] << 8) | (ip[2] << 16); ip += 3;
            const uint32_t last = bh & 1, type = (bh >> 1) & 3, bsize = bh >> 3;
            if (type == 3) return B200Z_E_CORRUPT;
            const size_t adv = type == 1 ? 1 : bsize;
            if ((size_t)(iend - ip) < adv) return B200Z_E_CORRUPT;
            if (!fcsBytes && type != 2) total += bsize;          // lower bound only; flagged unknown below
            ip += adv;
            if (last) break;
        }
        if (checksum) { if (iend - ip < 4) return B200Z_E_CORRUPT; ip += 4; }
        frames++;
    }
    if (nFrames) *nFrames = frames;
    if (contentSize) *contentSize = total;
    return unknown ? B200Z_E_UNSUPPORTED : B200Z_OK;
}


// The complete frames at the start of a buffer that may end inside a frame (streaming callers read the packed stream piece by
// piece).  *usedBytes = end of the last complete frame taken (skippable frames go with the frame that follows; trailing ones with
// the frame before), *contentBound = bytes those frames decode to -- exact where a frame declares its size, else the sum over its
// blocks of what a block can regenerate (raw / RLE: its size field, compressed: 128 KiB).  Stops before a frame that would take
// the sum past maxContent unless it is the first one.  B200Z_E_CORRUPT: the bytes at a frame start are no frame.
int b200z_zstd_frame_prefix(const void* srcv, size_t srcSize, uint64_t maxContent, size_t* usedBytes, uint64_t* contentBound, uint32_t* nFrames) {
    const uint8_t* base = (const uint8_t*)srcv; const uint8_t* ip = base; const uint8_t* iend = ip + srcSize;
    uint64_t total = 0; uint32_t frames = 0; size_t used = 0;
    while (ip < iend) {
        if (iend - ip < 4) break;
        const uint32_t magic = rd32(ip);
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {
            if (iend - ip < 8) break;
            const uint32_t sz = rd32(ip + 4);
            if ((size_t)(iend - ip) < 8 + (size_t)sz) break;
            ip += 8 + sz;
            if (frames) used = (size_t)(ip - base);              // after a frame: belongs to what was taken; before the first: to the frame to come
            continue;
        }
        if (magic != 0xFD2FB528u) { if (usedBytes) *usedBytes = used; return B200Z_E_CORRUPT; }
        if (iend - ip < 6) break;
        const uint8_t* p = ip + 5; const uint32_t fhd = ip[4];
        const uint32_t fcsFlag = fhd >> 6, single = (fhd >> 5) & 1, checksum = (fhd >> 2) & 1, didFlag = fhd & 3;
        if (fhd & 8) { if (usedBytes) *usedBytes = used; return B200Z_E_CORRUPT; }
        if (!single) p += 1;
        p += didFlag == 3 ? 4 : didFlag;
        const int fcsBytes = fcsFlag == 0 ? (int)single : (fcsFlag == 1 ? 2 : (fcsFlag == 2 ? 4 : 8));
        if (iend - p < fcsBytes) break;
        uint64_t fcs = 0; for (int i = 0; i < fcsBytes; i++) fcs |= (uint64_t)p[i] << (8 * i); if (fcsBytes == 2) fcs += 256;
        p += fcsBytes;
        uint64_t bound = 0; bool complete = false;
        for (;;) {
            if (iend - p < 3) break;
            const uint32_t bh = p[0] | (p[1] << 8) | (p[2] << 16); p += 3;
            const uint32_t last = bh & 1, type = (bh >> 1) & 3, bsize = bh >> 3;
            if (type == 3) { if (usedBytes) *usedBytes = used; return B200Z_E_CORRUPT; }
            const size_t adv = type == 1 ? 1 : bsize;
            if ((size_t)(iend - p) < adv) break;
            bound += type == 2 ? 131072u : bsize;
            p += adv;
            if (last) { complete = true; break; }
        }
        if (!complete) break;
        if (checksum) { if (iend - p < 4) break; p += 4; }
        const uint64_t content = fcsBytes ? fcs : bound;
        if (frames && total + content > maxContent) break;
        total += content; frames++; ip = p; used = (size_t)(ip - base);
    }
    if (usedBytes) *usedBytes = used;
    if (contentBound) *contentBound = total;
    if (nFrames) *nFrames = frames;
    return B200Z_OK;
}

}  // extern "C"

// hostDst != null: the output is also downloaded (after the execute kernel: every frame-warp is latency-bound and
// they all finish together, so splitting the download by frame groups only serialises the groups -- measured)
static int dec_impl(b200z_ctx* ctx, const void* d_src, size_t srcSize, void* d_dst, size_t dstCap, size_t* dstSize, void* hostDst) {
    if (!ctx || !dstSize || (!d_src && srcSize) || (!d_dst && dstCap)) return B200Z_E_PARAM;
    if ((uintptr_t)d_src & 7u) return fail(ctx, B200Z_E_PARAM, "device source must be 8-byte aligned%s");
    *dstSize = 0;
    if (!srcSize) return 0;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    // table capacities: frames >= 9 bytes each, blocks >= 3 bytes each; bounded to keep the tables small
    uint64_t frameCap = srcSize / 9 + 2; if (frameCap > (1u << 22)) frameCap = 1u << 22;
    uint64_t blockCap = srcSize / 3 + 2; { const uint64_t lim = srcSize / 128 + (1u << 20); if (blockCap > lim) blockCap = lim; }     // raw / RLE blocks cost a table entry only
    if (blockCap > 0x7FFFFFFFull) blockCap = 0x7FFFFFFFull;
    Arena& aFrames = ctx->decScratch[0]; Arena& aBlocks = ctx->decScratch[1]; Arena& aCounts = ctx->decScratch[2];
    Arena& aLits = ctx->decScratch[3]; Arena& aSeqs = ctx->decScratch[4];
    if (aFrames.reserve(frameCap * sizeof(DecFrame)) || aBlocks.reserve(blockCap * sizeof(DecBlock)) || aCounts.reserve(64))
        return fail(ctx, B200Z_E_MEMORY, "decoder table allocation failed%s");
    DecFrame* frames = (DecFrame*)aFrames.p; DecBlock* blocks = (DecBlock*)aBlocks.p;
    DecCounts* counts = (DecCounts*)aCounts.p; uint64_t* total = (uint64_t*)((uint8_t*)aCounts.p + 40);
    CU(cudaEventRecord(ctx->ev[0], st));
    DecCounts hc;
    static_assert(sizeof(DecCounts) == 40, "DecCounts is fetched as five words");
    // stage D0, first with mcmilk's size hints trusted (one hop per frame); a stream that then fails to index -- a skippable frame that only
    // looks like a hint -- is walked again block header by block header, as the reference does for every stream
    for (int pass = 0; pass < 2; pass++) {
        CU(cudaMemsetAsync(aCounts.p, 0, 64, st));
        launch_zstd_dec_find_frames((const uint8_t*)d_src, srcSize, frames, (uint32_t)frameCap, counts, pass == 0, st);
        CU(cudaGetLastError());
        { const int frc = b2z_fetch_small(ctx, &hc, counts, sizeof(hc), st); if (frc) return frc; }
        const uint32_t hinted = hc.nUnits;
        ctx->stat[B200Z_S_KERNEL_LAUNCHES] += 1;
        if (!hc.status) {
            launch_zstd_dec_index_blocks((const uint8_t*)d_src, srcSize, frames, hc.nFrames, blocks, (uint32_t)blockCap, counts, st);
            CU(cudaGetLastError());
            { const int frc = b2z_fetch_small(ctx, &hc, counts, sizeof(hc), st); if (frc) return frc; }
            ctx->stat[B200Z_S_KERNEL_LAUNCHES] += 3;
        }
        if (!hc.status || !hinted || (hc.status & ~B2Z_DERR_CORRUPT)) break;       // fine, or not the hints' fault
    }
    CU(cudaEventRecord(ctx->ev[3], st));
    if (hc.status) return dec_status_to_rc(ctx, hc.status);
    if (aLits.reserve((size_t)hc.nSlots * 131072ull + 64) || aSeqs.reserve((size_t)hc.nSlots * B2Z_DEC_MAXSEQ * 8ull + 64) ||
        ctx->decScratch[5].reserve(zstd_dec_entropy_scratch_bytes(hc.nBlocks) + zstd_dec_unit_state_bytes(hc.nFrames, hc.nBlocks) + 64))
        return fail(ctx, B200Z_E_MEMORY, "decoder scratch allocation failed (input too large for one pass)%s");
    launch_zstd_dec_entropy((const uint8_t*)d_src, srcSize, blocks, hc.nBlocks, (uint8_t*)aLits.p, (uint64_t*)aSeqs.p, ctx->decScratch[5].p, st, st, ctx->ev[4], ctx->ev[5]);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev[1], st));
    // stage J (zstd_dec.cu): frames whose units would form one chain -- what the reference's encoder writes -- are resolved by pointer
    // jumping.  Only a stream with a frame of enough blocks pays for the extra look at the counters.
    const uint32_t jumpMode = (uint32_t)ctx->decJump;
    const bool maybeJump = jumpMode == 2u || (jumpMode == 1u && (hc.maxFrameBlocks + B2Z_DEC_UNIT_BLOCKS - 1u) / B2Z_DEC_UNIT_BLOCKS >= B2Z_DEC_JUMP_MIN_UNITS);
    launch_zstd_dec_layout(frames, hc.nFrames, blocks, dstCap, counts, total, maybeJump ? jumpMode : 0u, st);
    CU(cudaGetLastError());
    if (maybeJump) {
        struct { DecCounts c; uint64_t total; } hj;
        { const int frc = b2z_fetch_small(ctx, &hj, counts, 48, st); if (frc) return frc; }
        if (hj.c.status) return dec_status_to_rc(ctx, hj.c.status);
        if (hj.c.nJump) {
            Arena& aPtr = ctx->decScratch[8];
            const uint32_t segLog = ctx->decJumpSegLog;
            if (aPtr.reserve(zstd_dec_jump_scratch_bytes(hj.total, segLog))) return fail(ctx, B200Z_E_MEMORY, "decoder scratch allocation failed (stage J pointers)%s");
            launch_zstd_dec_jump((const uint8_t*)d_src, frames, hc.nFrames, blocks, hc.nBlocks, (const uint8_t*)aLits.p, (const uint64_t*)aSeqs.p,
                                 (uint8_t*)d_dst, hj.total, segLog, counts, aPtr.p, st);
            CU(cudaGetLastError());
            ctx->stat[B200Z_S_KERNEL_LAUNCHES] += (2 + B2Z_DEC_JUMP_ROUNDS) * ((hj.total + (1ull << segLog) - 1) >> segLog);
            ctx->stat[B200Z_S_DEC_JUMP_FRAMES] += hj.c.nJump;
        }
    }
    uint32_t* unitState = (uint32_t*)((uint8_t*)ctx->decScratch[5].p + ((zstd_dec_entropy_scratch_bytes(hc.nBlocks) + 15u) & ~(size_t)15u));   // behind D1's scratch
    launch_zstd_dec_exec((const uint8_t*)d_src, frames, hc.nFrames, blocks, hc.nBlocks, (const uint8_t*)aLits.p, (const uint64_t*)aSeqs.p,
                         (uint8_t*)d_dst, counts, unitState, st);
    CU(cudaGetLastError());
    launch_zstd_dec_verify((const uint8_t*)d_src, frames, hc.nFrames, (const uint8_t*)d_dst, counts, st);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev[2], st));
    struct { DecCounts c; uint64_t total; } hr;                                // counts at +0, the total at +40 of the same 64-byte scratch
    { const int frc = b2z_fetch_small(ctx, &hr, counts, 48, st); if (frc) return frc; }
    ctx->stat[B200Z_S_KERNEL_LAUNCHES] += 7;
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]); ctx->stat[B200Z_S_DEC_PREPASS_MS] += ms;
    cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[1]); ctx->stat[B200Z_S_DEC_ENTROPY_MS] += ms;
    cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]); ctx->stat[B200Z_S_DEC_EXEC_MS] += ms;
    if (hr.c.status) return dec_status_to_rc(ctx, hr.c.status);
    if (hostDst && hr.total) { CU(cudaMemcpyAsync(hostDst, d_dst, hr.total, cudaMemcpyDeviceToHost, st)); CU(cudaStreamSynchronize(st)); ctx->stat[B200Z_S_D2H_BYTES] += (double)hr.total; }
    *dstSize = (size_t)hr.total;
    return 0;
}

extern "C" {

int b200z_zstd_decompress_device(b200z_ctx* ctx, const void* d_src, size_t srcSize, void* d_dst, size_t dstCap, size_t* dstSize) {
    return dec_impl(ctx, d_src, srcSize, d_dst, dstCap, dstSize, nullptr);
}

// Host-side split of a compressed stream into batches of whole frames (only when every frame declares its
// content size): returns false if the stream cannot be split that way (the caller then decodes in one shot).
struct HostBatch { size_t srcOff, srcEnd; uint64_t dstOff, dstSize; };
static bool split_frames(const uint8_t* src, size_t srcSize, uint64_t targetOut, std::vector<HostBatch>& out) {
    size_t ip = 0; HostBatch cur{0, 0, 0, 0}; uint64_t dstPos = 0;
    while (ip < srcSize) {
        if (srcSize - ip < 4) return false;
        const uint32_t magic = rd32(src + ip);
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {
            if (srcSize - ip < 8) return false;
            const size_t sz = rd32(src + ip + 4);
            if (srcSize - ip < 8 + sz) return false;
            ip += 8 + sz; continue;                                  // hints/skippable data stay attached to the following frame
        }
        if (magic != 0xFD2FB528u || srcSize - ip < 6) return false;
        size_t p = ip + 5; const uint32_t fhd = src[ip + 4];
        const uint32_t fcsFlag = fhd >> 6, single = (fhd >> 5) & 1, checksum = (fhd >> 2) & 1, didFlag = fhd & 3;
        if (!single) p += 1;
        p += didFlag == 3 ? 4 : didFlag;
        const int fcsBytes = fcsFlag == 0 ? (int)single : (fcsFlag == 1 ? 2 : (fcsFlag == 2 ? 4 : 8));
        if (!fcsBytes || srcSize < p + (size_t)fcsBytes) return false;
        uint64_t fcs = 0; for (int i = 0; i < fcsBytes; i++) fcs |= (uint64_t)src[p + i] << (8 * i); if (fcsBytes == 2) fcs += 256;
        p += fcsBytes;
        for (;;) {
            if (srcSize - p < 3) return false;
            const uint32_t bh = src[p] | (src[p + 1] << 8) | (src[p + 2] << 16); p += 3;
            const uint32_t last = bh & 1, type = (bh >> 1) & 3, bsize = bh >> 3;
            if (type == 3) return false;
            const size_t adv = type == 1 ? 1 : bsize;
            if (srcSize - p < adv) return false;
            p += adv;
            if (last) break;
        }
        if (checksum) { if (srcSize - p < 4) return false; p += 4; }
        ip = p;
        cur.srcEnd = ip; cur.dstSize += fcs; dstPos += fcs;
        if (cur.dstSize >= targetOut) { out.push_back(cur); cur = HostBatch{ ip, ip, dstPos, 0 }; }
    }
    if (cur.srcEnd > cur.srcOff || cur.dstSize) { cur.srcEnd = srcSize; out.push_back(cur); }
    else if (!out.empty()) out.back().srcEnd = srcSize;
    return true;
}

// one device's share of a host-pointer decompress: batches first, first + stride, ... through H2D (stream2) | decode kernels (stream)
// | D2H (stream3), double-buffered.  Every batch knows where its output goes (the frames declare their sizes), so devices never wait
// for each other.
struct DecJob {
    const uint8_t* src = nullptr; uint8_t* dst = nullptr; const std::vector<HostBatch>* batches = nullptr; size_t inStride = 0, outStride = 0;
    std::mutex m; int rc = 0; b200z_ctx* errCtx = nullptr;
    void fail_with(int code, b200z_ctx* c) { std::lock_guard<std::mutex> g(m); if (!rc) { rc = code; errCtx = c; } }
    bool failed() { std::lock_guard<std::mutex> g(m); return rc != 0; }
};
static void dec_worker(b200z_ctx* ctx, DecJob* job, size_t first, size_t stride) {
    auto run = [&]() -> int {
        CU(cudaSetDevice(ctx->device));
        const std::vector<HostBatch>& B = *job->batches;
        if (ctx->dIn.reserve(2 * job->inStride) || ctx->dOut.reserve(2 * job->outStride)) return fail(ctx, B200Z_E_MEMORY, "device staging allocation failed%s");
        uint8_t* dIn[2] = { (uint8_t*)ctx->dIn.p, (uint8_t*)ctx->dIn.p + job->inStride };
        uint8_t* dOut[2] = { (uint8_t*)ctx->dOut.p, (uint8_t*)ctx->dOut.p + job->outStride };
        // B200Z_TRACE=1: per-batch timeline on stderr (ms since the call began): upload done | kernels begin..end | download done
        const bool trace = getenv("B200Z_TRACE") != nullptr;
        std::vector<cudaEvent_t> tev;
        auto mark = [&](cudaStream_t s) { if (!trace) return; cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); tev.push_back(e); };
        mark(ctx->stream2);
        CU(cudaMemcpyAsync(dIn[0], job->src + B[first].srcOff, B[first].srcEnd - B[first].srcOff, cudaMemcpyHostToDevice, ctx->stream2));
        CU(cudaEventRecord(ctx->pe[0], ctx->stream2));
        size_t k = 0;
        for (size_t i = first; i < B.size(); i += stride, k++) {
            const int b = (int)(k & 1);
            if (job->failed()) break;
            if (i + stride < B.size()) {
                const HostBatch& nx = B[i + stride];
                CU(cudaMemcpyAsync(dIn[b ^ 1], job->src + nx.srcOff, nx.srcEnd - nx.srcOff, cudaMemcpyHostToDevice, ctx->stream2));
                CU(cudaEventRecord(ctx->pe[b ^ 1], ctx->stream2));
            }
            CU(cudaStreamWaitEvent(ctx->stream, ctx->pe[b], 0));
            if (k >= 2) CU(cudaStreamWaitEvent(ctx->stream, ctx->pe[2 + b], 0));
            mark(ctx->stream);
            size_t out = 0;
            int rc = dec_impl(ctx, dIn[b], B[i].srcEnd - B[i].srcOff, dOut[b], (size_t)B[i].dstSize, &out, nullptr);
            if (rc) return rc;
            mark(ctx->stream);
            if (out != B[i].dstSize) return fail(ctx, B200Z_E_CORRUPT, "frame content size mismatch%s");
            if (out) CU(cudaMemcpyAsync(job->dst + B[i].dstOff, dOut[b], out, cudaMemcpyDeviceToHost, ctx->stream3));
            CU(cudaEventRecord(ctx->pe[2 + b], ctx->stream3));
            mark(ctx->stream3);
            ctx->stat[B200Z_S_H2D_BYTES] += (double)(B[i].srcEnd - B[i].srcOff); ctx->stat[B200Z_S_D2H_BYTES] += (double)out;
        }
        CU(cudaStreamSynchronize(ctx->stream3));
        if (trace) {
            for (size_t j = 1; j + 2 < tev.size() + 0 && j + 2 <= tev.size() - 0; j += 3) {
                float a = 0, b2 = 0, c = 0;
                cudaEventElapsedTime(&a, tev[0], tev[j]); cudaEventElapsedTime(&b2, tev[0], tev[j + 1]); cudaEventElapsedTime(&c, tev[0], tev[j + 2]);
                fprintf(stderr, "[b200z dec dev %d] batch %zu: kernels %.1f..%.1f  download done %.1f\n", ctx->device, (j - 1) / 3, a, b2, c);
            }
            for (cudaEvent_t e : tev) cudaEventDestroy(e);
        }
        return 0;
    };
    const int rc = run();
    if (rc) job->fail_with(rc, ctx);
}

// Host-pointer decompress: batches of whole frames flow through H2D (stream2) | decode kernels (stream) | D2H (stream3).
int b200z_zstd_decompress_host(b200z_ctx* ctx, const void* src, size_t srcSize, void* dst, size_t dstCap, size_t* dstSize) {
    if (!ctx || !dstSize || (!src && srcSize) || (!dst && dstCap)) return B200Z_E_PARAM;
    *dstSize = 0;
    if (!srcSize) return 0;
    CU(cudaSetDevice(ctx->device));
    std::vector<HostBatch> batches;
    const size_t nDev = 1 + ctx->peers.size();
    uint64_t target = 1ull << ctx->hostBatchLog;                     // decoded bytes per batch
    if (nDev > 1) { const uint64_t per = (uint64_t)srcSize * 3 / (2 * nDev) + 1; if (per < target) target = per; }   // about two batches per device (packed size x 3 ~ output)
    if ((nDev == 1 && srcSize <= (target >> 2)) || !split_frames((const uint8_t*)src, srcSize, target, batches) || (batches.size() < 2 && nDev == 1) || batches.empty()) {
        // one shot
        if (ctx->dIn.reserve(srcSize + 64) || ctx->dOut.reserve(dstCap + 64)) return fail(ctx, B200Z_E_MEMORY, "device staging allocation failed%s");
        CU(cudaMemcpyAsync(ctx->dIn.p, src, srcSize, cudaMemcpyHostToDevice, ctx->stream));
        ctx->stat[B200Z_S_H2D_BYTES] += (double)srcSize;
        size_t out = 0;
        int rc = dec_impl(ctx, ctx->dIn.p, srcSize, ctx->dOut.p, dstCap, &out, dst);
        if (rc) return rc;
        *dstSize = out;
        return 0;
    }
    size_t maxIn = 0; uint64_t maxOut = 0, total = 0;
    for (const HostBatch& b : batches) { if (b.srcEnd - b.srcOff > maxIn) maxIn = b.srcEnd - b.srcOff; if (b.dstSize > maxOut) maxOut = b.dstSize; total += b.dstSize; }
    if (total > dstCap) return fail(ctx, B200Z_E_DSTSIZE, "destination too small%s");
    DecJob job; job.src = (const uint8_t*)src; job.dst = (uint8_t*)dst; job.batches = &batches;
    job.inStride = (maxIn + 64 + 255) & ~(size_t)255; job.outStride = ((size_t)maxOut + 64 + 255) & ~(size_t)255;
    const size_t nWorkers = nDev < batches.size() ? nDev : batches.size();
    std::vector<std::thread> threads;
    for (size_t d = 1; d < nWorkers; d++) threads.emplace_back(dec_worker, ctx->peers[d - 1], &job, d, nWorkers);
    dec_worker(ctx, &job, 0, nWorkers);
    for (std::thread& t : threads) t.join();
    if (job.rc) { if (job.errCtx && job.errCtx != ctx) snprintf(ctx->err, sizeof(ctx->err), "device %d: %.200s", job.errCtx->device, job.errCtx->err); return job.rc; }
    *dstSize = (size_t)total;
    return 0;
}

}  // extern "C"
