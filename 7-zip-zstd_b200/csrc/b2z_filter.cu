// b2z_filter.cu -- the pre/post filters that sit in front of the main coder in a 7z folder or an xz filter chain, on the GPU
// (SURVEY.md 8(f) item 3): Delta and the stateless branch converters ARM64, ARM, PPC, SPARC.  In place on a device buffer.
//
//   bra_kernel     one thread per 4-byte instruction: the rule of b2z_filter_ops.h applied to the word with its address -- the
//                  converters C/Bra.c:75-252 run as sequential loops are pure per-instruction functions.  16 B per thread, coalesced.
//   delta          encode: out[i] = in[i] - in[i - d], every byte independent (delta_enc_kernel reads the ORIGINAL neighbour: the
//                  launch goes through a scratch copy).  Decode: per residue class i mod d a running sum, done in three steps:
//                  column sums of tiles of `rows` x d bytes, an exclusive scan of those sums across tiles, then each tile adds its
//                  carry while it accumulates (C/Delta.c:20-169 is the sequential statement).
//   not here       x86 BCJ / BCJ2 / ARMT / RISCV / IA64: their scan carries state from byte to byte (C/Bra86.c:50-170) -- left to the host.
// Oracle statement: oracle/filter_oracle.c; both are checked against the reference's functions (oracle/_ref/libref_xz.so).
#include "b2z_device.cuh"
#include "b2z_filter_ops.h"
#ifndef B2Z_CUEMU
#include "b2z_ctx.h"
#endif

namespace b2z {

__global__ void __launch_bounds__(256)
bra_kernel(uint32_t* __restrict__ words, uint64_t nWords, uint32_t kind, int enc, uint32_t startOffset) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nWords; i += stride) {
        const uint32_t raw = words[i], ia = startOffset + (uint32_t)(i << 2);
        uint32_t out;
        if (kind == B200Z_F_ARM64) out = b2z_conv_arm64(raw, ia, enc);
        else if (kind == B200Z_F_ARM) out = b2z_conv_arm(raw, ia, enc);
        else if (kind == B200Z_F_PPC) out = b2z_bswap32(b2z_conv_ppc(b2z_bswap32(raw), ia, enc));
        else out = b2z_bswap32(b2z_conv_sparc(b2z_bswap32(raw), ia, enc));
        if (out != raw) words[i] = out;
    }
}

__global__ void __launch_bounds__(256)
delta_enc_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint64_t n, uint32_t dist) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (uint8_t)(in[i] - (i >= dist ? in[i - dist] : 0));
}

// tile t = bytes [t * rows * dist, (t + 1) * rows * dist): thread c < dist owns column c (one residue class inside the tile)
__global__ void __launch_bounds__(256)
delta_colsum_kernel(const uint8_t* __restrict__ data, uint64_t n, uint32_t dist, uint32_t rows, uint8_t* __restrict__ sums /* [tiles][dist] */) {
    const uint32_t c = threadIdx.x;
    if (c >= dist) return;
    const uint64_t t0 = (uint64_t)blockIdx.x * rows * dist;
    uint32_t s = 0;
    for (uint32_t r = 0; r < rows; r++) { const uint64_t i = t0 + (uint64_t)r * dist + c; if (i >= n) break; s += data[i]; }
    sums[(uint64_t)blockIdx.x * dist + c] = (uint8_t)s;
}
__global__ void __launch_bounds__(256)
delta_scan_kernel(uint8_t* __restrict__ sums, uint32_t tiles, uint32_t dist) {           // exclusive scan down each column, one CTA
    const uint32_t c = threadIdx.x;
    if (c >= dist) return;
    uint32_t run = 0;
    for (uint32_t t = 0; t < tiles; t++) { const uint32_t v = sums[(uint64_t)t * dist + c]; sums[(uint64_t)t * dist + c] = (uint8_t)run; run += v; }
}
__global__ void __launch_bounds__(256)
delta_dec_kernel(uint8_t* __restrict__ data, uint64_t n, uint32_t dist, uint32_t rows, const uint8_t* __restrict__ carry) {
    const uint32_t c = threadIdx.x;
    if (c >= dist) return;
    const uint64_t t0 = (uint64_t)blockIdx.x * rows * dist;
    uint32_t s = carry[(uint64_t)blockIdx.x * dist + c];
    for (uint32_t r = 0; r < rows; r++) { const uint64_t i = t0 + (uint64_t)r * dist + c; if (i >= n) break; s += data[i]; data[i] = (uint8_t)s; }
}

}  // namespace b2z

#ifndef B2Z_CUEMU
extern "C" {

// In place on a device buffer.  methodId: 7-Zip's filter ids (b2z_filter_ops.h); prop: delta distance (1..256) or the start offset
// ("pc") of the branch converters.  Branch converters leave a tail of n % 4 bytes untouched, like the reference (C/Bra.h:78-86).
int b200z_filter_device(b200z_ctx* ctx, uint32_t methodId, int encode, void* d_data, size_t n, uint32_t prop) {
    if (!ctx || (!d_data && n)) return B200Z_E_PARAM;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (methodId == B200Z_F_DELTA) {
        if (prop < 1 || prop > 256) return fail(ctx, B200Z_E_PARAM, "delta distance must be 1..256%s");
        if (!n) return 0;
        if (encode) {
            if (ctx->batchStage.reserve(n)) return fail(ctx, B200Z_E_MEMORY, "device scratch allocation failed%s");
            CU(cudaMemcpyAsync(ctx->batchStage.p, d_data, n, cudaMemcpyDeviceToDevice, st));
            b2z::delta_enc_kernel<<<(unsigned)((n + 255) / 256 < 65535 ? (n + 255) / 256 : 65535), 256, 0, st>>>((const uint8_t*)ctx->batchStage.p, (uint8_t*)d_data, n, prop);
            ctx->stat[B200Z_S_KERNEL_LAUNCHES] += 1;
        } else {
            const uint32_t rows = (65536u / prop) ? (65536u / prop) : 1u;
            const uint64_t tileBytes = (uint64_t)rows * prop;
            const uint32_t tiles = (uint32_t)((n + tileBytes - 1) / tileBytes);
            if (ctx->batchOff.reserve((size_t)tiles * prop + 64)) return fail(ctx, B200Z_E_MEMORY, "device scratch allocation failed%s");
            b2z::delta_colsum_kernel<<<tiles, 256, 0, st>>>((const uint8_t*)d_data, n, prop, rows, (uint8_t*)ctx->batchOff.p);
            b2z::delta_scan_kernel<<<1, 256, 0, st>>>((uint8_t*)ctx->batchOff.p, tiles, prop);
            b2z::delta_dec_kernel<<<tiles, 256, 0, st>>>((uint8_t*)d_data, n, prop, rows, (const uint8_t*)ctx->batchOff.p);
            ctx->stat[B200Z_S_KERNEL_LAUNCHES] += 3;
        }
    } else if (methodId == B200Z_F_ARM64 || methodId == B200Z_F_ARM || methodId == B200Z_F_PPC || methodId == B200Z_F_SPARC) {
        if ((uintptr_t)d_data & 3u) return fail(ctx, B200Z_E_PARAM, "branch converters need a 4-byte aligned buffer%s");
        if (prop & 3u) return fail(ctx, B200Z_E_UNSUPPORTED, "start offset must be a multiple of the instruction size%s");   // BranchMisc.cpp:57,99: E_INVALIDARG / E_NOTIMPL
        const uint64_t nWords = n >> 2;
        if (!nWords) return 0;
        b2z::bra_kernel<<<(unsigned)((nWords + 255) / 256 < 148u * 64u ? (nWords + 255) / 256 : 148u * 64u), 256, 0, st>>>((uint32_t*)d_data, nWords, methodId, encode, prop);
        ctx->stat[B200Z_S_KERNEL_LAUNCHES] += 1;
    } else return fail(ctx, B200Z_E_UNSUPPORTED, "filter not built on the GPU (x86 BCJ / BCJ2 / ARMT / RISCV / IA64 scan with carried state)%s");
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    return 0;
}

int b200z_filter_host(b200z_ctx* ctx, uint32_t methodId, int encode, void* data, size_t n, uint32_t prop) {
    if (!ctx || (!data && n)) return B200Z_E_PARAM;
    CU(cudaSetDevice(ctx->device));
    if (ctx->dIn.reserve(n + 64)) return fail(ctx, B200Z_E_MEMORY, "device staging allocation failed%s");
    if (n) CU(cudaMemcpyAsync(ctx->dIn.p, data, n, cudaMemcpyHostToDevice, ctx->stream));
    int rc = b200z_filter_device(ctx, methodId, encode, ctx->dIn.p, n, prop);
    if (rc) return rc;
    if (n) { CU(cudaMemcpyAsync(data, ctx->dIn.p, n, cudaMemcpyDeviceToHost, ctx->stream)); CU(cudaStreamSynchronize(ctx->stream)); }
    ctx->stat[B200Z_S_H2D_BYTES] += (double)n; ctx->stat[B200Z_S_D2H_BYTES] += (double)n;
    return 0;
}

}  // extern "C"
#endif
